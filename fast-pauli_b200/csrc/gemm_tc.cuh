// K5 tensor-core engine: 3xTF32 GEMM on the 5th-generation tensor cores (tcgen05 / UMMA, accumulator in TMEM).
//
//   C[M x N] = A[M x Kd] * B[Kd x N]        (all row-major fp32; split-K planes C + z*M*N)
//
// This is SummedPauliOp's dense coefficient contraction (reference: W(j,t) = sum_k coeffs(j,k) data(k,t),
// __summed_pauli_op.hpp:413-432, and out(k,t) = sum_j coeffs(j,k) E(j,t), :579-591) for complex64 plans, with the
// complex coefficient matrix stored as planar real rows ([Re; Im]) so one real GEMM produces both parts.
//
// Precision: a single TF32 pass (10-bit mantissa) misses the 1e-5 parity bar, so every operand is split into
// hi = tf32(v) and lo = tf32(v - hi) and three MMAs are issued per k-step: hi*hi + hi*lo + lo*hi (FP32 accumulate in
// TMEM); the dropped lo*lo term is O(2^-22) relative.
//
// Structure per CTA (128 threads, one 128 x 128 output tile, one split-K slice):
//   * all threads stage a 128 x 32 slab of A and the matching 32 x 128 slab of B: fp32 global loads, hi/lo split in
//     registers (the split is why TMA is not used: the tensor-core operands do not exist in memory), stores into the
//     UMMA canonical K-major no-swizzle layout (8-row x 16-byte core matrices; B is transposed on the way in);
//   * fence.proxy.async + __syncthreads, then ONE thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 and commits
//     to an mbarrier; the slab buffer is reused when the barrier flips;
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld (32x32b.x32), transposes through shared memory
//     and writes full 128-byte rows.
// SASS evidence: UTCHMMA-class instructions (tcgen05.mma), LDTM (tcgen05.ld).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fpk
{
namespace tc
{
constexpr int BM = 128, BN = 128, BK = 32; // output tile, k-slab (floats)
constexpr int NT = 128;                    // threads per CTA
constexpr uint32_t kTmemCols = 128;        // fp32 accumulator columns (= BN)
constexpr uint32_t kChunk = 16;            // bytes of one core-matrix row
constexpr uint32_t kSBO = 128;             // 8 rows x 16 B: next 8-row group
constexpr uint32_t kLBO = (BM / 8) * 128;  // next 16-byte k-chunk (one plane of all rows) = 2048 B
constexpr uint32_t kTileBytes = BM * BK * 4; // 16 KiB per operand tile
// smem: A_hi, A_lo, B_hi, B_lo tiles + epilogue staging (4 warps x 32 x 33 floats) + barrier + tmem address
constexpr uint32_t kStageFloats = 32 * 33;
constexpr size_t kSmemBytes = 4 * kTileBytes + 4 * kStageFloats * 4 + 64;

__device__ __forceinline__ uint32_t smem_u32(void const *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float to_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(kLBO >> 4) << 16) |
           (static_cast<uint64_t>(kSBO >> 4) << 32) | (1ull << 46);
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                            (static_cast<uint32_t>(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate)
                 : "memory");
}

// byte offset of (row r, 16-byte k-chunk c) inside an operand tile in the canonical layout
__device__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t c)
{
    return c * kLBO + (r >> 3) * kSBO + (r & 7u) * kChunk;
}

__global__ void __launch_bounds__(NT)
    gemm_3xtf32_kernel(float const *__restrict__ A, float const *__restrict__ B, float *__restrict__ C, uint32_t M,
                       uint64_t N, uint32_t Kd, uint32_t kchunk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *a_hi = smem, *a_lo = smem + kTileBytes, *b_hi = smem + 2 * kTileBytes, *b_lo = smem + 3 * kTileBytes;
    float *stage = reinterpret_cast<float *>(smem + 4 * kTileBytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 * kTileBytes + 4 * kStageFloats * 4);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

    uint32_t const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t const n0 = static_cast<uint64_t>(blockIdx.x) * BN;
    uint32_t const m0 = blockIdx.y * BM;
    uint32_t const kbeg = blockIdx.z * kchunk;
    uint32_t const kend = min(Kd, kbeg + kchunk);

    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tmem_base = *tmem_slot;

    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (uint32_t k0 = kbeg; k0 < kend; k0 += BK)
    {
        // ---- stage: thread handles row r = tid of each operand, 8 sixteen-byte k-chunks
        uint32_t const r = tid;
        bool const a_row_ok = (m0 + r) < M;
        bool const b_col_ok = (n0 + r) < N;
#pragma unroll
        for (uint32_t c = 0; c < BK / 4; ++c)
        {
            uint32_t const k = k0 + 4 * c;
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_row_ok && k + 3 < kend)
                av = *reinterpret_cast<float4 const *>(A + static_cast<uint64_t>(m0 + r) * Kd + k);
            else if (a_row_ok)
            {
                float const *ap = A + static_cast<uint64_t>(m0 + r) * Kd;
                av.x = k + 0 < kend ? ap[k + 0] : 0.f;
                av.y = k + 1 < kend ? ap[k + 1] : 0.f;
                av.z = k + 2 < kend ? ap[k + 2] : 0.f;
                av.w = k + 3 < kend ? ap[k + 3] : 0.f;
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_col_ok)
            {
                float const *bp = B + n0 + r; // B^T tile: row = output column, k runs over B's rows (coalesced over r)
                bv.x = k + 0 < kend ? bp[static_cast<uint64_t>(k + 0) * N] : 0.f;
                bv.y = k + 1 < kend ? bp[static_cast<uint64_t>(k + 1) * N] : 0.f;
                bv.z = k + 2 < kend ? bp[static_cast<uint64_t>(k + 2) * N] : 0.f;
                bv.w = k + 3 < kend ? bp[static_cast<uint64_t>(k + 3) * N] : 0.f;
            }
            float4 ah = make_float4(to_tf32(av.x), to_tf32(av.y), to_tf32(av.z), to_tf32(av.w));
            float4 al = make_float4(to_tf32(av.x - ah.x), to_tf32(av.y - ah.y), to_tf32(av.z - ah.z),
                                    to_tf32(av.w - ah.w));
            float4 bh = make_float4(to_tf32(bv.x), to_tf32(bv.y), to_tf32(bv.z), to_tf32(bv.w));
            float4 bl = make_float4(to_tf32(bv.x - bh.x), to_tf32(bv.y - bh.y), to_tf32(bv.z - bh.z),
                                    to_tf32(bv.w - bh.w));
            uint32_t const off = tile_off(r, c);
            *reinterpret_cast<float4 *>(a_hi + off) = ah;
            *reinterpret_cast<float4 *>(a_lo + off) = al;
            *reinterpret_cast<float4 *>(b_hi + off) = bh;
            *reinterpret_cast<float4 *>(b_lo + off) = bl;
        }
        // generic-proxy writes -> visible to the tensor core's async proxy, then hand over to the issuing thread
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0)
        {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t const sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
#pragma unroll
            for (uint32_t kk = 0; kk < BK / 8; ++kk)
            {
                uint32_t const koff = 2 * kk * kLBO; // two 16-byte chunks per K = 8 step
                mma_tf32(tmem_base, make_desc(sa_hi + koff), make_desc(sb_hi + koff), accumulate);
                mma_tf32(tmem_base, make_desc(sa_hi + koff), make_desc(sb_lo + koff), 1u);
                mma_tf32(tmem_base, make_desc(sa_lo + koff), make_desc(sb_hi + koff), 1u);
                accumulate = 1u;
            }
            // arrives on the barrier once every MMA issued so far has finished reading shared memory / writing TMEM
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(bar))
                         : "memory");
        }
        // everyone waits for the slab to be consumed before overwriting it (and, after the last slab, before the epilogue)
        {
            uint32_t done = 0;
            while (!done)
            {
                asm volatile("{\n\t"
                             ".reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t"
                             "}\n"
                             : "=r"(done)
                             : "r"(smem_u32(bar)), "r"(phase)
                             : "memory");
            }
            phase ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w owns TMEM lanes (= tile rows) 32w .. 32w+31
    float *C_plane = C + static_cast<uint64_t>(blockIdx.z) * M * N;
    float *st = stage + warp * kStageFloats;
    bool const any_k = kend > kbeg;
#pragma unroll 1
    for (uint32_t cb = 0; cb < BN / 32; ++cb)
    {
        uint32_t v[32];
        uint32_t const taddr = tmem_base + ((warp * 32u) << 16) + cb * 32u;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                       "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                       "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                       "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // lane holds row (32 warp + lane), columns cb*32 .. +31 -> transpose through smem for 128-byte row stores
#pragma unroll
        for (int j = 0; j < 32; ++j)
            st[lane * 33 + j] = any_k ? __uint_as_float(v[j]) : 0.f;
        __syncwarp();
        for (uint32_t rr = 0; rr < 32; ++rr)
        {
            uint32_t const row = m0 + warp * 32 + rr;
            uint64_t const col = n0 + cb * 32 + lane;
            if (row < M && col < N)
                C_plane[static_cast<uint64_t>(row) * N + col] = st[rr * 33 + lane];
        }
        __syncwarp();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}
} // namespace tc

inline bool gemm_tc_supported(uint32_t M, uint64_t N, uint32_t Kd, uint32_t splitK)
{
    // float4 loads of A rows need 16-byte aligned rows; everything else is guarded in the kernel
    return M > 0 && N > 0 && Kd > 0 && (Kd % 4 == 0) && splitK >= 1 && splitK <= 65535 && (M + tc::BM - 1) / tc::BM <= 65535;
}

// returns 0 on success (launch enqueued), non-zero when the caller should fall back to the SIMT engine
inline int gemm_tc_3xtf32(cudaStream_t stream, float const *A, float const *B, float *C, uint32_t M, uint64_t N,
                          uint32_t Kd, uint32_t splitK, uint32_t kchunk)
{
    // the attribute is per device: one bit per device of this process
    static uint64_t configured = 0, failed = 0;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    uint64_t const bit = 1ull << (dev & 63);
    if (!(configured & bit))
    {
        if (cudaFuncSetAttribute(tc::gemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(tc::kSmemBytes)) != cudaSuccess)
        {
            (void)cudaGetLastError();
            failed |= bit;
        }
        configured |= bit;
    }
    bool const ok = !(failed & bit);
    if (!ok || (reinterpret_cast<uintptr_t>(A) % 16) != 0 || (splitK > 1 && kchunk % tc::BK != 0))
        return 1;
    dim3 grid(static_cast<unsigned>((N + tc::BN - 1) / tc::BN), (M + tc::BM - 1) / tc::BM, splitK);
    tc::gemm_3xtf32_kernel<<<grid, tc::NT, tc::kSmemBytes, stream>>>(A, B, C, M, N, Kd, kchunk);
    return 0;
}
} // namespace fpk
