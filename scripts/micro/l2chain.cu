// Microbenchmark: can two coset passes be CHAINED through the L2?
//
// A multi-pass coset plan (K3i: 64 random strings = 8 rank-8 passes) re-streams the whole batch once per pass:
// pass 1 moves in + out, every later pass in + old out + out (23 GiB for 8 passes at 20 qubits x 64 complex128).
// Two consecutive passes A, B with spans S_A, S_B only couple rows inside one coset of S_A + S_B (rank 16: 65 536
// rows x 256 bytes = 16 MiB per column tile), so a persistent grid can run A and B chunk by chunk
//     A(c0) A(c1) B(c0) A(c2) B(c1) ...
// and B's reads of `in` and of the old output hit the L2 (126 MB) if it keeps ~3 chunks x 32 MiB.  This program
// measures what that buys with plain coalesced loads / stores and the same dependency structure (a B tile reads
// one 256-byte row segment of each of the 256 A tiles of its chunk): `unchained` = P passes one after another,
// `chained` = P/2 chained pairs with per-chunk completion counters.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2chain l2chain.cu && ./l2chain [log2 chunk rows] [passes]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            printf("%s: %s\n", #x, cudaGetErrorString(e_));                                                            \
            exit(1);                                                                                                   \
        }                                                                                                              \
    } while (0)

constexpr int kThreads = 512; // 16 warps: a warp instruction moves two 256-byte row segments
constexpr int kRowBytes = 256;
constexpr int kTileRows = 256;

// one tile: 256 rows of 256 bytes; rows = first + k * stride.  out = in * 1.0000001 (+ old out)
__device__ __forceinline__ void do_tile(double2 const *__restrict__ in, double2 *__restrict__ out, uint64_t first,
                                        uint64_t stride, int rmw)
{
    uint32_t const warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t const half = lane >> 4, jv = lane & 15u;
    double2 v[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        uint64_t const row = first + (2u * (warp + 16u * i) + half) * stride;
        v[i] = __ldcg(&in[row * 16 + jv]);
        if (rmw)
            o[i] = __ldcg(&out[row * 16 + jv]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        uint64_t const row = first + (2u * (warp + 16u * i) + half) * stride;
        double2 r{v[i].x * 1.0000001, v[i].y * 1.0000001};
        if (rmw)
        {
            r.x += o[i].x;
            r.y += o[i].y;
        }
        out[row * 16 + jv] = r;
    }
}

// read-modify-write by reduction: out += f(in) with fire-and-forget RED.ADD.F64 (the addition happens in the L2, the old
// value never travels to the SM): 2 units of SM <-> L2 traffic per pass instead of 3
__device__ __forceinline__ void do_tile_red(double2 const *__restrict__ in, double2 *__restrict__ out, uint64_t first,
                                            uint64_t stride)
{
    uint32_t const warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t const half = lane >> 4, jv = lane & 15u;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        uint64_t const row = first + (2u * (warp + 16u * i) + half) * stride;
        v[i] = __ldcg(&in[row * 16 + jv]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        uint64_t const row = first + (2u * (warp + 16u * i) + half) * stride;
        double *dst = reinterpret_cast<double *>(&out[row * 16 + jv]);
        atomicAdd(dst, v[i].x * 1.0000001);
        atomicAdd(dst + 1, v[i].y * 1.0000001);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
    pass_red_kernel(double2 const *in, double2 *out, uint64_t nChunks, uint32_t chunkTiles, int strided)
{
    uint64_t const total = nChunks * chunkTiles;
    for (uint64_t q = blockIdx.x; q < total; q += gridDim.x)
    {
        uint64_t const c = q / chunkTiles, j = q % chunkTiles;
        uint64_t const base = c * chunkTiles * kTileRows;
        if (strided)
            do_tile_red(in, out, base + j, chunkTiles);
        else
            do_tile_red(in, out, base + j * kTileRows, 1);
    }
}

// plain pass: tiles of contiguous (stride 1) or strided rows inside each chunk, round-robin over a persistent grid
__global__ void __launch_bounds__(kThreads, 1)
    pass_kernel(double2 const *in, double2 *out, uint64_t nChunks, uint32_t chunkTiles, int strided, int rmw)
{
    uint64_t const total = nChunks * chunkTiles;
    for (uint64_t q = blockIdx.x; q < total; q += gridDim.x)
    {
        uint64_t const c = q / chunkTiles, j = q % chunkTiles;
        uint64_t const base = c * chunkTiles * kTileRows;
        if (strided)
            do_tile(in, out, base + j, chunkTiles, rmw);
        else
            do_tile(in, out, base + j * kTileRows, 1, rmw);
    }
}

// chained pair: A tiles (contiguous rows) and B tiles (strided rows: one row of every A tile of the chunk), in the
// order A0 | A1 B0 | A2 B1 | ... ; B(c) waits for counter[c] == chunkTiles * 16 (every warp of every A tile)
__global__ void __launch_bounds__(kThreads, 1)
    chain_kernel(double2 const *in, double2 *out, uint64_t nChunks, uint32_t chunkTiles, int rmwA, unsigned *counter,
                 int deferred, int dist)
{
    uint64_t const blocks = 2 * nChunks;
    uint32_t const lane = threadIdx.x & 31u;
    long long pending = -1;
    for (uint64_t q = blockIdx.x; q < blocks * chunkTiles; q += gridDim.x)
    {
        uint64_t const pb = q / chunkTiles, j = q % chunkTiles;
        // block sequence with pipeline distance `dist`: A_0 .. A_{dist-1}, then (A_k, B_{k-dist}) ..., then the last B's
        bool isB;
        uint64_t c;
        uint64_t const d = static_cast<uint64_t>(dist);
        if (pb < d)
        {
            isB = false;
            c = pb;
        }
        else if (pb >= 2 * nChunks - d)
        {
            isB = true;
            c = nChunks - (2 * nChunks - pb);
        }
        else
        {
            uint64_t const k = (pb - d) / 2;
            isB = ((pb - d) & 1u) != 0;
            c = isB ? k : k + d;
        }
        uint64_t const base = c * chunkTiles * kTileRows;
        if (!isB)
        {
            if (deferred && pending >= 0)
            {
                __threadfence();
                if (lane == 0)
                    atomicAdd(&counter[pending], 1u);
            }
            do_tile(in, out, base + j * kTileRows, 1, rmwA);
            if (deferred)
                pending = static_cast<long long>(c);
            else
            {
                __threadfence();
                if (lane == 0)
                    atomicAdd(&counter[c], 1u);
            }
        }
        else
        {
            if (pending >= 0)
            {
                __threadfence();
                if (lane == 0)
                    atomicAdd(&counter[pending], 1u);
                pending = -1;
            }
            if (lane == 0)
            {
                unsigned const target = chunkTiles * (kThreads / 32);
                unsigned v;
                do
                {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&counter[c]) : "memory");
                } while (v < target);
            }
            __syncwarp();
            do_tile(in, out, base + j, chunkTiles, 1);
        }
    }
    if (pending >= 0)
    {
        __threadfence();
        if (lane == 0)
            atomicAdd(&counter[pending], 1u);
    }
}

int main(int argc, char **argv)
{
    int const logChunkRows = argc > 1 ? atoi(argv[1]) : 16; // rows of one chunk (16 -> 16 MiB in + 16 MiB out)
    int const passes = argc > 2 ? atoi(argv[2]) : 8;
    uint64_t const rows = 1ull << 22; // 2^20 rows x 4 column tiles of 256 bytes = 1 GiB
    uint64_t const bytes = rows * kRowBytes;
    uint32_t const chunkTiles = (1u << logChunkRows) / kTileRows;
    uint64_t const nChunks = rows >> logChunkRows;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int const grid = prop.multiProcessorCount;
    double2 *in, *out;
    unsigned *counter;
    CK(cudaMalloc(&in, bytes));
    CK(cudaMalloc(&out, bytes));
    CK(cudaMalloc(&counter, nChunks * sizeof(unsigned) * 8));
    CK(cudaMemset(in, 0, bytes));
    CK(cudaMemset(out, 0, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms;
    printf("chunk = 2^%d rows (%.0f MiB in + %.0f MiB out), %llu chunks, %d passes, grid %d\n", logChunkRows,
           (double)(1ull << logChunkRows) * kRowBytes / 1048576.0, (double)(1ull << logChunkRows) * kRowBytes / 1048576.0,
           (unsigned long long)nChunks, passes, grid);
    for (int rep = 0; rep < 2; ++rep)
    {
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 5; ++it)
            for (int p = 0; p < passes; ++p)
                pass_kernel<<<grid, kThreads>>>(in, out, nChunks, chunkTiles, p & 1, p > 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double const units = 2.0 + 3.0 * (passes - 1);
        if (rep)
            printf("unchained: %.3f ms per %d passes (%.2f TB/s on %.0f GiB of HBM traffic)\n", ms / 5, passes,
                   units * bytes / (ms / 5) / 1e9, units);
    }
    for (int rep = 0; rep < 2; ++rep)
    {
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 5; ++it)
            for (int p = 0; p < passes; ++p)
            {
                if (p == 0)
                    pass_kernel<<<grid, kThreads>>>(in, out, nChunks, chunkTiles, 0, 0);
                else
                    pass_red_kernel<<<grid, kThreads>>>(in, out, nChunks, chunkTiles, p & 1);
            }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep)
            printf("unchained, read-modify-write passes by RED.ADD.F64: %.3f ms per %d passes\n", ms / 5, passes);
    }
    for (int deferred = 0; deferred < 2; ++deferred)
        for (int dist = 1; dist <= 2; ++dist)
            for (int rep = 0; rep < 2; ++rep)
            {
                CK(cudaEventRecord(e0));
                for (int it = 0; it < 5; ++it)
                {
                    CK(cudaMemsetAsync(counter, 0, nChunks * sizeof(unsigned) * 8));
                    for (int p = 0; p < passes / 2; ++p)
                        chain_kernel<<<grid, kThreads>>>(in, out, nChunks, chunkTiles, p > 0, counter + p * nChunks,
                                                         deferred, dist);
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                double const units = 2.0 + 3.0 * (passes / 2 - 1);
                if (rep)
                    printf("chained (deferred release %d, distance %d): %.3f ms per %d passes (%.2f TB/s on the %.0f GiB an "
                           "ideal L2 leaves)\n",
                           deferred, dist, ms / 5, passes, units * bytes / (ms / 5) / 1e9, units);
            }
    CK(cudaGetLastError());
    return 0;
}
