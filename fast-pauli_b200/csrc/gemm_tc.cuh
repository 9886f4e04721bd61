// K5 tensor-core engine: 3xTF32 tcgen05 GEMM for the SummedPauliOp coefficient contraction (complex64 plans).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so callers take the FP32 SIMT engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fpk
{
inline bool gemm_tc_supported(uint32_t, uint64_t, uint32_t, uint32_t)
{
    return false;
}
inline int gemm_tc_3xtf32(cudaStream_t, float const *, float const *, float *, uint32_t, uint64_t, uint32_t, uint32_t,
                          uint32_t)
{
    return -1;
}
} // namespace fpk
