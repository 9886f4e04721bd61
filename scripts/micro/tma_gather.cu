// Microbenchmark: how fast can one B200 move the rows of a GF(2) coset tile (2^R scattered rows x W bytes) between HBM
// and shared memory?  Design data for the TMA-fed coset kernel (K3e).
//
//   mode 0  LDGSTS in, LDS + STG out (what coset_kernel does today), 2-3 CTAs / SM
//   mode 1  TMA tile::gather4 in -> TMA tile::scatter4 out, persistent CTA, S-stage ring, one issuing warp
//   mode 2  TMA gather4 in -> consumer warps LDS + STG out
//   mode 4  TMA gather4 in -> consumer warps: G gathers (LDS.128 + complex FMA) per vector -> STS -> scatter4 out
//   mode 5  as 4 but LDGSTS in / LDS+STG out by dedicated copy warps (no TMA), persistent
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_gather tma_gather.cu
// run:   tma_gather <mode> <W bytes> <lanes issuing> <G> <ctas per SM> <stages>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                            \
            exit(2);                                                                                                   \
        }                                                                                                              \
    } while (0)

static constexpr int kQ = 20;
static constexpr int kRowBytes = 1024; // 64 complex128 columns
static constexpr int kR = 8;
static constexpr int kRows = 1 << kR;

struct Params
{
    uint32_t basis[kR];
    uint32_t nonpivot;
    uint32_t n_tiles;   // cosets * column tiles
    uint32_t n_ct;      // column tiles per coset
    uint32_t W;         // bytes per row segment
    uint32_t G;         // gathers per vector (mode 4/5)
    uint32_t lanes;     // issuing lanes
    int *err;
};

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *b, uint32_t parity, int *err)
{
    long long t0 = clock64();
    while (!mbar_try(b, parity))
    {
        if (clock64() - t0 > 4000000000ll)
        {
            atomicExch(err, 1);
            return false;
        }
    }
    return true;
}
__device__ __forceinline__ void tma_gather4(void *dst, CUtensorMap const *tm, int c0, int r0, int r1, int r2, int r3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
                 "%5, %6}], [%7];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_scatter4(CUtensorMap const *tm, int c0, int r0, int r1, int r2, int r3, void const *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(tm),
                 "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t deposit(uint32_t src, uint32_t mask)
{
    uint32_t res = 0;
    for (uint32_t bb = 1; mask; bb <<= 1)
    {
        uint32_t low = mask & (~mask + 1);
        if (src & bb)
            res |= low;
        mask &= mask - 1;
    }
    return res;
}
__device__ __forceinline__ uint32_t comb_of(uint32_t const *basis, uint32_t l)
{
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kR; ++k)
        if ((l >> k) & 1u)
            c ^= basis[k];
    return c;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, void const *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}

// ------------------------------------------------------------------------------------------------- mode 0
__global__ void __launch_bounds__(256) k_ldgsts_copy(Params p, uint4 const *__restrict__ in, uint4 *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t comb[kRows];
    uint4 *tile = reinterpret_cast<uint4 *>(smem);
    uint32_t const tid = threadIdx.x;
    comb[tid] = comb_of(p.basis, tid);
    __syncthreads();
    uint32_t const twc = p.W / 16, rowvecs = kRowBytes / 16;
    uint32_t const coset = blockIdx.x / p.n_ct, ct = blockIdx.x % p.n_ct;
    uint32_t const base = deposit(coset, p.nonpivot);
    uint32_t const nvec = kRows * twc;
    for (uint32_t v = tid; v < nvec; v += 256)
    {
        uint32_t l = v / twc, j = v % twc;
        cp_async16(&tile[v], &in[static_cast<uint64_t>(base ^ comb[l]) * rowvecs + ct * twc + j]);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    for (uint32_t v = tid; v < nvec; v += 256)
    {
        uint32_t l = v / twc, j = v % twc;
        out[static_cast<uint64_t>(base ^ comb[l]) * rowvecs + ct * twc + j] = tile[v];
    }
}

// ------------------------------------------------------------------------------------------------- mode 1
template <int S> __global__ void __launch_bounds__(32) k_tma_copy(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, Params p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[S];
    __shared__ uint32_t comb[kRows];
    uint32_t const lane = threadIdx.x;
    for (uint32_t l = lane; l < kRows; l += 32)
        comb[l] = comb_of(p.basis, l);
    if (lane == 0)
    {
        for (int s = 0; s < S; ++s)
            mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t const tile_bytes = kRows * p.W;
    uint32_t const n_ops = kRows / 4;
    uint32_t const wdbl = p.W / 8; // doubles per row segment
    uint32_t n_mine = 0;
    for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
        ++n_mine;

    auto load = [&](uint32_t k) {
        uint32_t t = blockIdx.x + k * gridDim.x;
        uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
        uint32_t base = deposit(coset, p.nonpivot);
        uint32_t s = k % S;
        if (lane == 0)
            mbar_expect_tx(&full[s], tile_bytes);
        __syncwarp();
        if (lane < p.lanes)
            for (uint32_t op = lane; op < n_ops; op += p.lanes)
                tma_gather4(smem + s * tile_bytes + op * 4 * p.W, &tmIn, ct * wdbl, base ^ comb[4 * op], base ^ comb[4 * op + 1],
                            base ^ comb[4 * op + 2], base ^ comb[4 * op + 3], &full[s]);
    };
    for (uint32_t k = 0; k + 1 < S && k < n_mine; ++k)
        load(k);
    for (uint32_t it = 0; it < n_mine; ++it)
    {
        if (it + S - 1 < n_mine)
        {
            if (it >= 1)
            {
                bulk_wait_read0();
                __syncwarp();
            }
            load(it + S - 1);
        }
        uint32_t s = it % S;
        if (!mbar_wait(&full[s], (it / S) & 1u, p.err))
            return;
        uint32_t t = blockIdx.x + it * gridDim.x;
        uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
        uint32_t base = deposit(coset, p.nonpivot);
        if (lane < p.lanes)
            for (uint32_t op = lane; op < n_ops; op += p.lanes)
                tma_scatter4(&tmOut, ct * wdbl, base ^ comb[4 * op], base ^ comb[4 * op + 1], base ^ comb[4 * op + 2],
                             base ^ comb[4 * op + 3], smem + s * tile_bytes + op * 4 * p.W);
        bulk_commit();
    }
    bulk_wait0();
}

// ------------------------------------------------------------------------------------------------- modes 2 and 4
// warp 0 = TMA producer (and scatter issuer in mode 4), warps 1..8 = consumers.
// MODE 2: consumers copy tile -> global with LDS + STG.   MODE 4: consumers do G gathers per vector + STS to the
// out buffer; producer scatters the out buffer.
template <int S, int MODE> __global__ void __launch_bounds__(288) k_tma_pipe(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, Params p, uint4 *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[S], empty[S], out_full, out_free;
    __shared__ uint32_t comb[kRows];
    uint32_t const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t l = tid; l < kRows; l += blockDim.x)
        comb[l] = comb_of(p.basis, l);
    if (tid == 0)
    {
        for (int s = 0; s < S; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        mbar_init(&out_full, 8);
        mbar_init(&out_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t const tile_bytes = kRows * p.W;
    unsigned char *obuf = smem + S * tile_bytes; // MODE 4 only
    uint32_t const n_ops = kRows / 4;
    uint32_t const wdbl = p.W / 8, twc = p.W / 16, rowvecs = kRowBytes / 16;
    uint32_t n_mine = 0;
    for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
        ++n_mine;

    if (warp == 0)
    {
        auto load = [&](uint32_t k) {
            uint32_t t = blockIdx.x + k * gridDim.x;
            uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
            uint32_t base = deposit(coset, p.nonpivot);
            uint32_t s = k % S;
            if (lane == 0)
                mbar_expect_tx(&full[s], tile_bytes);
            __syncwarp();
            if (lane < p.lanes)
                for (uint32_t op = lane; op < n_ops; op += p.lanes)
                    tma_gather4(smem + s * tile_bytes + op * 4 * p.W, &tmIn, ct * wdbl, base ^ comb[4 * op], base ^ comb[4 * op + 1],
                                base ^ comb[4 * op + 2], base ^ comb[4 * op + 3], &full[s]);
        };
        for (uint32_t k = 0; k < S && k < n_mine; ++k)
            load(k);
        for (uint32_t it = 0; it < n_mine; ++it)
        {
            // refill the stage tile it used, once the consumers released it (they release it before they write the
            // out buffer)
            if (it + S < n_mine)
            {
                if (!mbar_wait(&empty[it % S], (it / S) & 1u, p.err))
                    return;
                load(it + S);
            }
            if (MODE == 4)
            {
                // scatter the out buffer of tile it when the consumers filled it
                if (!mbar_wait(&out_full, it & 1u, p.err))
                    return;
                uint32_t t = blockIdx.x + it * gridDim.x;
                uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
                uint32_t base = deposit(coset, p.nonpivot);
                if (lane < p.lanes)
                    for (uint32_t op = lane; op < n_ops; op += p.lanes)
                        tma_scatter4(&tmOut, ct * wdbl, base ^ comb[4 * op], base ^ comb[4 * op + 1], base ^ comb[4 * op + 2],
                                     base ^ comb[4 * op + 3], obuf + op * 4 * p.W);
                bulk_commit();
                bulk_wait_read0();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&out_free);
            }
        }
        bulk_wait0();
        return;
    }

    // ---------------- consumers: 256 threads
    uint32_t const ctid = tid - 32;
    for (uint32_t it = 0; it < n_mine; ++it)
    {
        uint32_t s = it % S;
        if (!mbar_wait(&full[s], (it / S) & 1u, p.err))
            return;
        uint4 const *tile = reinterpret_cast<uint4 const *>(smem + s * tile_bytes);
        uint32_t t = blockIdx.x + it * gridDim.x;
        uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
        uint32_t base = deposit(coset, p.nonpivot);
        if (MODE == 2)
        {
            uint32_t const nvec = kRows * twc;
            for (uint32_t v = ctid; v < nvec; v += 256)
            {
                uint32_t l = v / twc, j = v % twc;
                out[static_cast<uint64_t>(base ^ comb[l]) * rowvecs + ct * twc + j] = tile[v];
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]);
        }
        else
        {
            // thread = row ctid, 16 vectors (W = 256) with the column rotation key = row & 15 (conflict-free dense tile)
            double2 acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                acc[j] = make_double2(0.0, 0.0);
            uint32_t const key16 = (ctid & 15u) << 4;
            double2 const *tb = reinterpret_cast<double2 const *>(tile);
            for (uint32_t g = 0; g < p.G; ++g)
            {
                uint32_t const xl = (g < 8) ? (1u << g) : (g * 37u) & 255u;
                double2 const d = make_double2(1.0 + g * 0.25, (ctid & 1) ? -0.5 : 0.5);
                uint32_t const rb = ((ctid ^ xl) << 8);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                {
                    double2 v = *reinterpret_cast<double2 const *>(reinterpret_cast<unsigned char const *>(tb) + (rb | ((j << 4) ^ key16)));
                    acc[j].x = fma(d.x, v.x, acc[j].x);
                    acc[j].x = fma(-d.y, v.y, acc[j].x);
                    acc[j].y = fma(d.x, v.y, acc[j].y);
                    acc[j].y = fma(d.y, v.x, acc[j].y);
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]);
            // out buffer free?  (freed by the producer after the scatter of tile it-1 finished reading)
            if (it >= 1)
                if (!mbar_wait(&out_free, (it - 1) & 1u, p.err))
                    return;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                *reinterpret_cast<double2 *>(obuf + ((ctid << 8) | ((j << 4) ^ key16))) = acc[j];
            fence_async();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&out_full);
        }
    }
}

// ------------------------------------------------------------------------------------------------- mode 5
// no TMA: 2 CTAs/SM style is mode 0; here the persistent warp-specialised variant with LDGSTS producers
// (4 copy warps: load next tile with cp.async, store previous out buffer with LDS+STG), 8 consumer warps.
template <int S> __global__ void __launch_bounds__(384) k_ldgsts_pipe(Params p, uint4 const *__restrict__ in, uint4 *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[S], empty[S], out_full, out_free;
    __shared__ uint32_t comb[kRows];
    uint32_t const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t l = tid; l < kRows; l += blockDim.x)
        comb[l] = comb_of(p.basis, l);
    if (tid == 0)
    {
        for (int s = 0; s < S; ++s)
        {
            mbar_init(&full[s], 128);
            mbar_init(&empty[s], 8);
        }
        mbar_init(&out_full, 8);
        mbar_init(&out_free, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t const tile_bytes = kRows * p.W;
    unsigned char *obuf = smem + S * tile_bytes;
    uint32_t const twc = p.W / 16, rowvecs = kRowBytes / 16;
    uint32_t const nvec = kRows * twc;
    uint32_t n_mine = 0;
    for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
        ++n_mine;

    if (warp < 4)
    {
        auto load = [&](uint32_t k) {
            uint32_t t = blockIdx.x + k * gridDim.x;
            uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
            uint32_t base = deposit(coset, p.nonpivot);
            uint4 *tile = reinterpret_cast<uint4 *>(smem + (k % S) * tile_bytes);
            for (uint32_t v = tid; v < nvec; v += 128)
            {
                uint32_t l = v / twc, j = v % twc;
                cp_async16(&tile[v], &in[static_cast<uint64_t>(base ^ comb[l]) * rowvecs + ct * twc + j]);
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[k % S])) : "memory");
        };
        for (uint32_t k = 0; k < S && k < n_mine; ++k)
            load(k);
        for (uint32_t it = 0; it < n_mine; ++it)
        {
            if (!mbar_wait(&out_full, it & 1u, p.err))
                return;
            uint32_t t = blockIdx.x + it * gridDim.x;
            uint32_t coset = t / p.n_ct, ct = t % p.n_ct;
            uint32_t base = deposit(coset, p.nonpivot);
            uint4 const *ob = reinterpret_cast<uint4 const *>(obuf);
            for (uint32_t v = tid; v < nvec; v += 128)
            {
                uint32_t l = v / twc, j = v % twc;
                out[static_cast<uint64_t>(base ^ comb[l]) * rowvecs + ct * twc + j] = ob[v];
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&out_free);
            if (it + S < n_mine)
            {
                if (!mbar_wait(&empty[it % S], (it / S) & 1u, p.err))
                    return;
                load(it + S);
            }
        }
        return;
    }
    uint32_t const ctid = tid - 128;
    for (uint32_t it = 0; it < n_mine; ++it)
    {
        uint32_t s = it % S;
        if (!mbar_wait(&full[s], (it / S) & 1u, p.err))
            return;
        double2 acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            acc[j] = make_double2(0.0, 0.0);
        uint32_t const key16 = (ctid & 15u) << 4;
        unsigned char const *tb = smem + s * tile_bytes;
        for (uint32_t g = 0; g < p.G; ++g)
        {
            uint32_t const xl = (g < 8) ? (1u << g) : (g * 37u) & 255u;
            double2 const d = make_double2(1.0 + g * 0.25, (ctid & 1) ? -0.5 : 0.5);
            uint32_t const rb = ((ctid ^ xl) << 8);
#pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                double2 v = *reinterpret_cast<double2 const *>(tb + (rb | ((j << 4) ^ key16)));
                acc[j].x = fma(d.x, v.x, acc[j].x);
                acc[j].x = fma(-d.y, v.y, acc[j].x);
                acc[j].y = fma(d.x, v.y, acc[j].y);
                acc[j].y = fma(d.y, v.x, acc[j].y);
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&empty[s]);
        if (it >= 1)
            if (!mbar_wait(&out_free, (it - 1) & 1u, p.err))
                return;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            *reinterpret_cast<double2 *>(obuf + ((ctid << 8) | ((j << 4) ^ key16))) = acc[j];
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&out_full);
    }
}

__global__ void k_fill(double *p, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = static_cast<double>((i * 2654435761ull) & 0xffffff) * (1.0 / 16777216.0);
}
__global__ void k_cmp(double const *a, double const *b, size_t n, unsigned long long *bad)
{
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        c += a[i] != b[i];
    if (c)
        atomicAdd(bad, c);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, cuuint64_t const *, cuuint64_t const *,
                             cuuint32_t const *, cuuint32_t const *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void *ptr, uint32_t W, uint32_t box_rows)
{
    CUtensorMap tm;
    cuuint64_t dims[2] = {kRowBytes / 8, 1ull << kQ};
    cuuint64_t strides[1] = {kRowBytes};
    cuuint32_t box[2] = {W / 8, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        printf("cuTensorMapEncodeTiled failed: %d (W=%u box_rows=%u)\n", (int)r, W, box_rows);
        exit(3);
    }
    return tm;
}

int main(int argc, char **argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 1;
    uint32_t W = argc > 2 ? atoi(argv[2]) : 256;
    uint32_t lanes = argc > 3 ? atoi(argv[3]) : 32;
    uint32_t G = argc > 4 ? atoi(argv[4]) : 8;
    int cps = argc > 5 ? atoi(argv[5]) : 1;
    int stages = argc > 6 ? atoi(argv[6]) : 2;
    int box_rows = argc > 7 ? atoi(argv[7]) : 1;

    size_t const n_dbl = (size_t(1) << kQ) * (kRowBytes / 8);
    double *in, *out;
    CK(cudaMalloc(&in, n_dbl * 8));
    CK(cudaMalloc(&out, n_dbl * 8));
    k_fill<<<1184, 256>>>(in, n_dbl);
    CK(cudaMemset(out, 0, n_dbl * 8));
    int *err;
    unsigned long long *bad;
    CK(cudaMalloc(&err, 4));
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemset(err, 0, 4));
    CK(cudaMemset(bad, 0, 8));

    // 8 random masks -> reduced echelon basis over GF(2)^20
    Params p{};
    {
        uint32_t m[kR];
        uint64_t st = 0x9e3779b97f4a7c15ull;
        auto rnd = [&]() {
            st ^= st << 13;
            st ^= st >> 7;
            st ^= st << 17;
            return static_cast<uint32_t>(st >> 20) & ((1u << kQ) - 1);
        };
        for (;;)
        {
            for (int k = 0; k < kR; ++k)
                m[k] = rnd();
            uint32_t b[kR];
            memcpy(b, m, sizeof b);
            uint32_t piv = 0;
            int rank = 0;
            for (int bit = kQ - 1; bit >= 0 && rank < kR; --bit)
            {
                int sel = -1;
                for (int k = rank; k < kR; ++k)
                    if ((b[k] >> bit) & 1)
                    {
                        sel = k;
                        break;
                    }
                if (sel < 0)
                    continue;
                std::swap(b[rank], b[sel]);
                for (int k = 0; k < kR; ++k)
                    if (k != rank && ((b[k] >> bit) & 1))
                        b[k] ^= b[rank];
                piv |= 1u << bit;
                ++rank;
            }
            if (rank == kR)
            {
                memcpy(p.basis, b, sizeof b);
                p.nonpivot = ~piv & ((1u << kQ) - 1);
                break;
            }
        }
    }
    p.W = W;
    p.n_ct = kRowBytes / W;
    p.n_tiles = (1u << (kQ - kR)) * p.n_ct;
    p.G = G;
    p.lanes = lanes;
    p.err = err;

    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&enc), cudaEnableDefault, &qres));
    if (!enc)
    {
        printf("no cuTensorMapEncodeTiled\n");
        return 3;
    }
    CUtensorMap tmIn = make_map(enc, in, W, box_rows), tmOut = make_map(enc, out, W, box_rows);

    uint32_t const tile_bytes = kRows * W;
    int const grid = 148 * cps;
    auto launch = [&]() {
        switch (mode)
        {
        case 0:
            k_ldgsts_copy<<<p.n_tiles, 256, tile_bytes>>>(p, reinterpret_cast<uint4 const *>(in), reinterpret_cast<uint4 *>(out));
            break;
        case 1:
            if (stages == 2)
                k_tma_copy<2><<<grid, 32, 2 * tile_bytes>>>(tmIn, tmOut, p);
            else
                k_tma_copy<3><<<grid, 32, 3 * tile_bytes>>>(tmIn, tmOut, p);
            break;
        case 2:
            if (stages == 2)
                k_tma_pipe<2, 2><<<grid, 288, 2 * tile_bytes>>>(tmIn, tmOut, p, reinterpret_cast<uint4 *>(out));
            else
                k_tma_pipe<3, 2><<<grid, 288, 3 * tile_bytes>>>(tmIn, tmOut, p, reinterpret_cast<uint4 *>(out));
            break;
        case 4:
            k_tma_pipe<2, 4><<<grid, 288, 3 * tile_bytes>>>(tmIn, tmOut, p, reinterpret_cast<uint4 *>(out));
            break;
        case 5:
            k_ldgsts_pipe<2><<<grid, 384, 3 * tile_bytes>>>(p, reinterpret_cast<uint4 const *>(in), reinterpret_cast<uint4 *>(out));
            break;
        default:
            printf("bad mode\n");
            exit(1);
        }
    };
    CK(cudaFuncSetAttribute(k_ldgsts_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_tma_copy<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_tma_copy<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_tma_pipe<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_tma_pipe<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_tma_pipe<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_ldgsts_pipe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));

    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int herr = 0;
    CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
    unsigned long long hbad = 0;
    if (mode <= 2)
    {
        k_cmp<<<1184, 256>>>(in, out, n_dbl, bad);
        CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
    }
    launch();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int const iters = 5;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i)
        launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("mode %d W %4u lanes %2u G %2u ctas/SM %d stages %d box_rows %d : %.3f ms  %.0f GB/s  timeout=%d mismatches=%llu\n", mode, W, lanes, G,
           cps, stages, box_rows, ms, 2.0 * n_dbl * 8 / (ms * 1e-3) / 1e9, herr, hbad);
    return 0;
}
