// Microbenchmark: FP64 mma.sync.m8n8k4 issue rate vs DFMA on one GPU (build: nvcc -gencode arch=compute_100a,code=sm_100a)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_kernel(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
    for (int i = 0; i < 8; ++i)
        c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i)
        s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[16];
    for (int i = 0; i < 16; ++i)
        c[i] = i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            c[i] = fma(a, c[i], b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i)
        s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int const iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2)
    {
        dmma_kernel<<<148, warps * 32>>>(out, 10);
        cudaEventRecord(e0);
        dmma_kernel<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * warps * iters * 8 * 256; // 8x8x4 FMA per mma
        printf("DMMA m8n8k4 warps/SM=%2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", warps, ms,
               2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        dfma_kernel<<<148, warps * 32>>>(out, 10);
        cudaEventRecord(e0);
        dfma_kernel<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fma = 148.0 * warps * 32 * iters * 16;
        printf("DFMA           warps/SM=%2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, ms, 2 * fma / ms / 1e9,
               fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
