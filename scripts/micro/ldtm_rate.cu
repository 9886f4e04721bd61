// Microbenchmark: throughput of tcgen05.ld (SASS LDTM) used as a GATHER engine -- every load reads 32 lanes x NCOL
// 32-bit columns at a warp-uniform, data-dependent column address.  Question it answers for K3h (coset_tmem.cuh):
// can tensor memory feed one 16-byte element per lane (x4) fast enough to replace the LDS.128 gathers that bound the
// shared-memory coset kernels (128 B/clk/SM)?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/ldtm_rate scripts/micro/ldtm_rate.cu
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
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                            \
            exit(2);                                                                                                   \
        }                                                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(void const *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int NCOL> __device__ __forceinline__ void ldtm(uint32_t (&v)[NCOL], uint32_t taddr);
template <> __device__ __forceinline__ void ldtm<4>(uint32_t (&v)[4], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void ldtm<8>(uint32_t (&v)[8], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void ldtm<16>(uint32_t (&v)[16], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                 "%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// UNR loads in flight per tcgen05.wait::ld; DFMA > 0 adds that many dependent-free DFMAs per loaded 16-byte element
template <int NCOL, int UNR, int DFMA>
__global__ void __launch_bounds__(512, 1) k_ldtm(uint32_t iters, uint32_t active_warps, unsigned long long *cycles,
                                                 unsigned long long *errors, double *sink)
{
    extern __shared__ unsigned char dyn[];
    __shared__ uint32_t slot;
    uint32_t const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tbase = slot;
    uint32_t const q = warp & 3u, k = warp >> 2; // lane quarter of this warp, index among the warps sharing it
    uint32_t const lane_g = q * 32 + lane;
    uint32_t const tq = tbase + ((q * 32u) << 16);
    // fill: warp k of a quarter writes columns 128k .. 128k+127 with (lane << 16 | column)
    for (uint32_t c0 = 128 * k; c0 < 128 * k + 128; c0 += 4)
    {
        uint32_t const a = (lane_g << 16) | c0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tq + c0), "r"(a), "r"(a + 1),
                     "r"(a + 2), "r"(a + 3));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    unsigned long long err = 0;
    double acc[4] = {0, 0, 0, 0};
    long long t0 = 0, t1 = 0;
    if (warp < active_warps)
    {
        uint32_t state = 0x9e3779b9u * (warp + 1) + blockIdx.x;
        t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it)
        {
            uint32_t v[UNR][NCOL];
            uint32_t col[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
            {
                state = state * 1664525u + 1013904223u;
                col[u] = (state >> 8) & (512u - NCOL) & ~(uint32_t)(NCOL - 1);
                ldtm<NCOL>(v[u], tq + col[u]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < UNR; ++u)
            {
#pragma unroll
                for (int j = 0; j < NCOL; ++j)
                    err += (v[u][j] != ((lane_g << 16) | (col[u] + j)));
                if (DFMA)
                {
#pragma unroll
                    for (int e = 0; e < NCOL / 4; ++e)
                    {
                        double const re = __hiloint2double(v[u][4 * e + 1], v[u][4 * e]);
                        double const im = __hiloint2double(v[u][4 * e + 3], v[u][4 * e + 2]);
#pragma unroll
                        for (int d = 0; d < DFMA; ++d)
                            acc[d & 3] = fma(d & 1 ? im : re, 1.0000001, acc[d & 3]);
                    }
                }
            }
        }
        t1 = clock64();
    }
    if (lane == 0 && warp < active_warps)
    {
        atomicMax(&cycles[blockIdx.x], static_cast<unsigned long long>(t1 - t0));
        atomicAdd(errors, err);
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] == 1.2345)
        sink[0] = acc[0];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int NCOL, int UNR, int DFMA> void run(uint32_t warps)
{
    int const grid = 148;
    unsigned long long *cyc, *err;
    double *sink;
    CK(cudaMalloc(&cyc, grid * 8));
    CK(cudaMalloc(&err, 8));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(cyc, 0, grid * 8));
    CK(cudaMemset(err, 0, 8));
    uint32_t const iters = 4096;
    size_t const smem = 120 * 1024;
    CK(cudaFuncSetAttribute(k_ldtm<NCOL, UNR, DFMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ldtm<NCOL, UNR, DFMA><<<grid, 512, smem>>>(iters, warps, cyc, err, sink);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned long long h[148], e;
    CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&e, err, 8, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i)
        mean += h[i];
    mean /= grid;
    double const bytes = double(iters) * UNR * NCOL * 128.0 * warps; // per SM
    printf("x%-2d in-flight %d  dfma/elem %d  warps %2u: %8.0f cycles  %6.1f B/clk/SM  (%5.1f B/clk per quarter)  %5.2f clk per LDTM per quarter  errors %llu\n",
           NCOL, UNR, DFMA, warps, mean, bytes / mean, bytes / mean / (warps < 4 ? warps : 4),
           mean / (double(iters) * UNR * warps / (warps < 4 ? warps : 4)), e);
    cudaFree(cyc);
    cudaFree(err);
    cudaFree(sink);
}

int main()
{
    for (uint32_t w : {1u, 4u, 8u, 16u})
    {
        run<4, 8, 0>(w);
        run<8, 4, 0>(w);
        run<16, 2, 0>(w);
    }
    run<4, 1, 0>(16);
    run<4, 2, 0>(16);
    run<4, 4, 0>(16);
    // with the arithmetic of the coset kernels: 4 DFMA per gathered complex128
    run<4, 8, 4>(16);
    run<4, 8, 4>(8);
    run<16, 2, 4>(16);
    return 0;
}
