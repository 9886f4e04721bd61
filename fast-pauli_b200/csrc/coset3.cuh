// K3i: TMA-fed coset kernel with DIRECT stores, for passes whose x-masks carry ONE string each (i.i.d. random strings:
// 64 random 20-qubit strings have 64 different x-masks, so every pass of their coset plan is of this kind).
//
// Same tiling and tile ring as K3f (coset2.cuh): a persistent CTA per SM, a producer warp that streams coset tiles
// (256 scattered rows x 256 bytes) into a ring of three 64 KiB shared-memory buffers with TMA tile::gather4.  What
// is different is the consumer side.  K3e / K3f give a thread one ROW of the tile (so that the row factors
// D_g(l) = sum_{s in g} +-c_s live in its registers) and therefore have to transpose the results through shared
// memory before they can be stored as coalesced row segments: per 64 KiB tile 512 STS + 512 LDS wavefronts, two
// barriers, and -- on accumulating passes -- one more staged copy of the old output rows, all on the load/store
// unit that the gathers already saturate.  With one string per x-mask the row factor is +-c_g: the coefficient sits
// in the constant bank and the sign is ONE BIT per (row, mask).  So here a warp owns two rows x 16 vectors per step
// (lanes run along the batch axis):
//
//   gather  LDS.128 of row (l ^ xl_g): the two half-warps read two adjacent 256-byte rows -- conflict-free without
//           any rotation or padding
//   sign    bit g of a per-thread word (par(l & zl_g), formed once per kernel) xor a per-coset word (par(base & z_g))
//           flips the coefficient: one shift + two LOP3 per complex FMA
//   store   STG.128 straight from the accumulator: a warp writes two full 256-byte row segments; on accumulating
//           passes the old values arrive by LDG.128 in the same mapping (pulled into L2 one tile ahead by
//           prefetch.global.L2) -- no staging buffer, no barrier among the consumer warps at all
//
// The consumer warps synchronise only with the producer (mbarrier full / empty per buffer), so they drift apart and
// the stores of one warp overlap the gathers of the others.  Results are bit-identical to K3b / K3e on the same
// plan (same operation order per output element).
// Reference semantics: PauliOp::apply (PO:399-468): out(i,t) (+)= sum_s h_s m_s(i) psi(i ^ x_s, t).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "coset2.cuh"

namespace fpk
{

#ifdef FP_DIR_PROFILE
__device__ unsigned long long g_dir_prof[8];
#define DIR_T(i)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        long long now_ = clock64();                                                                                    \
        prof_[i] += static_cast<unsigned long long>(now_ - t_prev_);                                                   \
        t_prev_ = now_;                                                                                                \
    } while (0)
#else
#define DIR_T(i)
#endif

constexpr int kDirMaxMasks = 32;           // x-masks (= strings) of one pass
constexpr int kDirConsumerWarps = 16;
#ifndef FP_DIR_NPROD
#define FP_DIR_NPROD 1
#endif
constexpr int kDirProducerWarps = FP_DIR_NPROD; // live warps of the producer warpgroup (1, 2 or 4): each issues 64 / n gather4 per tile
constexpr int kDirThreads = kDirConsumerWarps * 32 + 128; // + the producer warpgroup (one live warp, see K3f)

// One pass as kernel parameters (constant bank): per x-mask the coefficient (times (-i)^nY), the full z-mask (sign of
// the coset base), the local x and z coordinates in the pass' basis.
template <typename T> struct DirStrings
{
    Cx<T> c[kDirMaxMasks];
    uint64_t z[kDirMaxMasks];
    uint32_t xl[kDirMaxMasks];
    uint32_t zl[kDirMaxMasks];
    // single-state form (one 2^n vector viewed as 2^(n-4) rows x 16 "columns" = the 4 lowest index bits): the string
    // also permutes the columns (j -> j ^ xlo) and signs them ((-1)^popc(j & zlo)); both 0 for ordinary batches
    uint8_t xlo[kDirMaxMasks];
    uint8_t zlo[kDirMaxMasks];
    uint32_t n;
};

// base of the NEXT coset: deposit_bits(c + 1, mask) from deposit_bits(c, mask) by a masked increment (the carry runs
// through the pivot bits, which are set for the addition and cleared afterwards) -- 3 instructions instead of a loop over
// the mask's bits; matters when a coset has a single column tile (single states: one tile per coset)
__device__ __forceinline__ uint32_t next_coset_base(uint32_t base, uint32_t nonpivot_mask)
{
    return ((base | ~nonpivot_mask) + 1u) & nonpivot_mask;
}

// flips the sign of v when bit 31 of t is set
__device__ __forceinline__ double dir_flip(double v, uint32_t t)
{
    return __hiloint2double(__double2hiint(v) ^ static_cast<int>(t & 0x80000000u), __double2loint(v));
}
__device__ __forceinline__ float dir_flip(float v, uint32_t t)
{
    return __int_as_float(__float_as_int(v) ^ static_cast<int>(t & 0x80000000u));
}

template <typename T, int EPV, int NCH, int IB = 4> // NCH = ceil(n / 8): the mask loop is unrolled; IB row pairs at a time
__global__ void __launch_bounds__(kDirThreads, 1)
    coset_dir_tma_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, uint64_t nTiles,
                         CVec<T, EPV> *__restrict__ out, int beta, const __grid_constant__ DirStrings<T> strs,
                         const __grid_constant__ CUtensorMap tm_in)
{
    using Vec = CVec<T, EPV>;
    constexpr int TWC = 16, R = 8;
    constexpr uint32_t ROW_SHIFT = 8;
    constexpr int ITERS = 128 / kDirConsumerWarps; // row pairs per warp and tile

    extern __shared__ __align__(1024) unsigned char smem_dt[];
    __shared__ uint64_t s_full[kFewTmaBufs], s_empty[kFewTmaBufs];
    __shared__ uint32_t s_comb[256]; // XOR offsets of the 256 local rows

    uint32_t const tid = threadIdx.x;
    if (tid < 256)
        s_comb[tid] = static_cast<uint32_t>(comb_of<R>(pass.basis, tid));
    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < kFewTmaBufs; ++b)
        {
            few_mbar_init(&s_full[b], 1);
            few_mbar_init(&s_empty[b], kDirConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's contiguous range of (coset, column tile) work items, column tiles fastest
    uint64_t const t0 = nTiles * blockIdx.x / gridDim.x, t1 = nTiles * (blockIdx.x + 1) / gridDim.x;

    if (tid >= kDirConsumerWarps * 32)
    {
        // ------------------------------------------------ producer warpgroup: hands its registers to the consumers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid >= kDirConsumerWarps * 32 + 32 * kDirProducerWarps)
            return;
        uint32_t const lane = tid & 31u, pw = (tid >> 5) - kDirConsumerWarps;
        uint64_t p_coset = 0;
        uint32_t p_ct = 0, p_base = 0;
#ifdef FP_DIR_PROFILE
        long long t_prev_ = clock64();
        unsigned long long prof_[8] = {};
#endif
        uint32_t buf = 0, round = 0; // ring position: tile i of this CTA sits in buffer i % 3, round = i / 3
        for (uint64_t t = t0; t < t1; ++t)
        {
            DIR_T(5); // producer: issue
            if (round)
                few_mbar_wait(&s_empty[buf], (round - 1) & 1u);
            DIR_T(4); // producer: waiting for an empty buffer
            if (t == t0 || ++p_ct == nColTiles)
            {
                p_coset = t == t0 ? t0 / nColTiles : p_coset + 1;
                p_ct = t == t0 ? static_cast<uint32_t>(t0 - p_coset * nColTiles) : 0u;
                p_base = t == t0 ? static_cast<uint32_t>(deposit_bits(p_coset, pass.nonpivot_mask))
                                 : next_coset_base(p_base, static_cast<uint32_t>(pass.nonpivot_mask));
            }
            uint32_t const ct = p_ct, base = p_base;
            if (lane == 0 && pw == 0)
                few_mbar_expect_tx(&s_full[buf], static_cast<uint32_t>(kFewTmaTile));
            __syncwarp();
            int const c0 = static_cast<int>(ct) * TWC * static_cast<int>(16 / sizeof(T));
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                uint32_t const op = lane + 32 * h; // rows 4*op .. 4*op+3
                if (kDirProducerWarps == 2 && static_cast<uint32_t>(h) != pw)
                    continue;
                if (kDirProducerWarps == 4 && (static_cast<uint32_t>(h) != (pw >> 1) || ((lane >> 4) != (pw & 1u))))
                    continue;
                few_tma_gather4(smem_dt + buf * kFewTmaTile + (static_cast<size_t>(op) << (ROW_SHIFT + 2)), &tm_in, c0,
                                base ^ s_comb[4 * op], base ^ s_comb[4 * op + 1], base ^ s_comb[4 * op + 2],
                                base ^ s_comb[4 * op + 3], &s_full[buf]);
            }
            if (++buf == kFewTmaBufs)
            {
                buf = 0;
                ++round;
            }
        }
#ifdef FP_DIR_PROFILE
        if (lane == 0)
        {
            atomicAdd(&g_dir_prof[4], prof_[4]);
            atomicAdd(&g_dir_prof[5], prof_[5]);
        }
#endif
        return;
    }

    // ---------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    uint32_t const warp = tid >> 5, lane = tid & 31u;
    uint32_t const half = lane >> 4, jv = lane & 15u;
    uint32_t const ng = strs.n;
    // local row of this thread in step i: l_i = 2 * (warp + 16 i) + half; its row-local sign bits for every mask
    uint32_t sl[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i)
    {
        uint32_t const l = 2u * (warp + kDirConsumerWarps * i) + half;
        uint32_t w = 0;
        for (uint32_t g = 0; g < ng; ++g)
            w |= (__popc((l & strs.zl[g]) ^ ((jv & strs.zlo[g]) << 8)) & 1u) << g; // + the column's own sign (single state)
        sl[i] = w;
    }
    uint32_t const col_off = jv << 4;

    uint64_t coset = 0;
    uint32_t ct = 0, base = 0, pbm = 0;
#ifdef FP_DIR_PROFILE
    long long t_prev_ = clock64();
    unsigned long long prof_[8] = {};
#endif
    uint32_t buf = 0, round = 0;
    for (uint64_t t = t0; t < t1; ++t)
    {
        if (t == t0 || ++ct == nColTiles)
        {
            coset = t == t0 ? t0 / nColTiles : coset + 1;
            ct = t == t0 ? static_cast<uint32_t>(t0 - coset * nColTiles) : 0u;
            base = t == t0 ? static_cast<uint32_t>(deposit_bits(coset, pass.nonpivot_mask)) // launched for <= 30 qubits
                           : next_coset_base(base, static_cast<uint32_t>(pass.nonpivot_mask));
            pbm = 0;
            for (uint32_t g = 0; g < ng; ++g)
                pbm |= (__popc(base & static_cast<uint32_t>(strs.z[g])) & 1u) << g;
        }
        uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jv;

        Vec old[ITERS];
        if (beta)
        {
            // the next tile's old output rows -> L2 (a whole tile of gathers away from their use) ...
            if (t + 1 < t1)
            {
                bool const same = ct + 1 < nColTiles;
                uint32_t const ct_n = same ? ct + 1 : 0u;
                uint32_t const base_n = same ? base : next_coset_base(base, static_cast<uint32_t>(pass.nonpivot_mask));
                uint64_t const vcol_n = static_cast<uint64_t>(ct_n) * TWC + jv;
#pragma unroll
                for (int i = 0; i < ITERS; ++i)
                {
                    uint32_t const l = 2u * (warp + kDirConsumerWarps * i) + half;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(&out[static_cast<uint64_t>(base_n ^ s_comb[l]) * rowvecs + vcol_n]));
                }
            }
            // ... and this tile's into registers, in the mapping they are added in
#pragma unroll
            for (int i = 0; i < ITERS; ++i)
            {
                uint32_t const l = 2u * (warp + kDirConsumerWarps * i) + half;
                old[i] = out[static_cast<uint64_t>(base ^ s_comb[l]) * rowvecs + vcol];
            }
        }
        DIR_T(0); // per-tile setup, old-row loads issued
        few_mbar_wait(&s_full[buf], round & 1u);
        DIR_T(1); // waiting for the tile

        // IB row pairs at a time: IB independent accumulator chains per thread (one chain alone is bound by the DFMA
        // latency: 16 dependent DFMA per output element at 8 masks), gathers of two masks in flight
#pragma unroll
        for (int i0 = 0; i0 < ITERS; i0 += IB)
        {
            uint32_t own[IB], sgn[IB];
            Cx<T> acc[IB][EPV];
#pragma unroll
            for (int b = 0; b < IB; ++b)
            {
                uint32_t const l = 2u * (warp + kDirConsumerWarps * (i0 + b)) + half;
                own[b] = buf * static_cast<uint32_t>(kFewTmaTile) + ((l << ROW_SHIFT) | col_off); // offset in the ring
                sgn[b] = sl[i0 + b] ^ pbm;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[b][e] = Cx<T>{0, 0};
            }
#pragma unroll
            for (int g2 = 0; g2 < NCH * 4; ++g2)
            {
                Vec v[2][IB];
#pragma unroll
                for (int k = 0; k < 2; ++k)
                {
                    int const g = g2 * 2 + k;
                    if (static_cast<uint32_t>(g) < ng)
                    {
                        // < 64 KiB: the XOR stays inside the buffer (xlo: column permutation of the single-state form)
                        uint32_t const xo = (strs.xl[g] << ROW_SHIFT) | (static_cast<uint32_t>(strs.xlo[g]) << 4);
#pragma unroll
                        for (int b = 0; b < IB; ++b)
                            v[k][b] = *reinterpret_cast<Vec const *>(smem_dt + (own[b] ^ xo));
                    }
                }
#pragma unroll
                for (int k = 0; k < 2; ++k)
                {
                    int const g = g2 * 2 + k;
                    if (static_cast<uint32_t>(g) < ng)
                    {
                        // +-c_g psi = c_g (+-psi): the coefficient stays a constant-bank operand of the DFMAs, the sign
                        // bit of the row goes into the gathered value (one shift + one LOP3 per real component)
                        Cx<T> const c = strs.c[g];
#pragma unroll
                        for (int b = 0; b < IB; ++b)
                        {
                            uint32_t const t = sgn[b] << (31 - g);
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                            {
                                Cx<T> const w{dir_flip(v[k][b].e[e].re, t), dir_flip(v[k][b].e[e].im, t)};
                                cfma(acc[b][e], c, w);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < IB; ++b)
            {
                uint32_t const l = 2u * (warp + kDirConsumerWarps * (i0 + b)) + half;
                Vec r;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    r.e[e] = acc[b][e];
                if (beta)
                {
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        r.e[e].re += old[i0 + b].e[e].re;
                        r.e[e].im += old[i0 + b].e[e].im;
                    }
                }
                out[static_cast<uint64_t>(base ^ s_comb[l]) * rowvecs + vcol] = r;
            }
        }
        // this warp's gathers of the buffer are done (their values were consumed above); order them before the
        // asynchronous-proxy refill and hand the buffer back
        DIR_T(2); // gathers + FMAs + stores
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            few_mbar_arrive(&s_empty[buf]);
        if (++buf == kFewTmaBufs)
        {
            buf = 0;
            ++round;
        }
        DIR_T(3); // fence + arrive
    }
#ifdef FP_DIR_PROFILE
    if (lane == 0)
        for (int k = 0; k < 4; ++k)
            atomicAdd(&g_dir_prof[k], prof_[k]);
#endif
}

} // namespace fpk
