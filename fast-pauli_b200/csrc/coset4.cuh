// K3j: TMA-fed coset kernel with direct stores for passes of EIGHT INDEPENDENT x-masks with any number of strings
// each -- the north-star operator (64 strings over 8 x-masks) in one pass, and every pass of 64 i.i.d. strings.
//
// Same ring and producer as K3i (coset3.cuh): a persistent CTA per SM, a producer warp streaming 256-row coset tiles
// (256 bytes per row) into three 64 KiB buffers with TMA tile::gather4, 16 consumer warps whose lanes run along the
// batch axis and store their accumulators straight to global memory.  Two things are new:
//
//  * PAIRED MASKS.  A lane owns 2^RB rows of the tile (its top RB local bits) x 8 / 2^RB vectors of the row segment.
//    The pass' basis is re-chosen on the host from the masks themselves, for RB = 2:
//        b_0..5 = m_0, m_2, m_4, m_5, m_6, m_7,   b_6 = m_0^m_1, b_7 = m_2^m_3
//    (RB = 3: b_0..4 = m_0, m_2, m_4, m_6, m_7, b_5..7 = m_0^m_1, m_2^m_3, m_4^m_5), so that in local coordinates
//    m_0 = e_0 and m_1 = e_0 ^ e_6: the vectors psi(l_i ^ e_0) gathered for mask 0 are, permuted over the lane's rows
//    i, exactly the ones mask 1 needs (psi(l_i ^ e_0 ^ e_6) = psi(l_{i^1} ^ e_0)).  RB pairs + (8 - 2 RB) single masks:
//    48 gathers (LDS.128) per lane and tile at RB = 2, 40 at RB = 3, instead of 64; the accumulation order per output
//    element stays mask 0, 1, ..., 7, so results are bit-identical to K3e / K3b on the same plan.
//  * ROW FACTORS FROM A TABLE.  D_g(l) = sum_{s in g} +-c_s (any number of strings per mask) depends on the coset and
//    the row, not on the column tile: the 8 x 256 factors of a coset are formed once (4 per consumer thread, strings in
//    plan order from the constant bank) into a 32 KiB shared-memory table whenever the CTA moves to a new coset (column
//    tiles run fastest); a warp forms exactly the 16 rows x 8 masks it reads itself, so the table costs no barrier among
//    the consumer warps (a CTA-wide barrier per coset was measured: the drain costs ~3 us per coset, 0.03 ms at 20 qubits
//    x 64).  The main loop reads a factor with one 16-byte load per (row, mask) -- 8 x 2^RB per lane and tile.  These
//    loads weigh like gathers (a 16-byte shared load is served a quarter-warp at a time whether or not the lanes share
//    the address), which is why RB = 2 (48 + 32 loads, 76.8 M shared-memory wavefronts per launch at 20 qubits x 64)
//    beats RB = 3 (40 + 64 loads, 94.7 M): 0.397 against 0.411 ms.
//
// Reference semantics: PauliOp::apply (PO:399-468): out(i,t) (+)= sum_s h_s m_s(i) psi(i ^ x_s, t).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "coset3.cuh"

namespace fpk
{

constexpr int kPairMasks = 8;
constexpr int kPairMaxStrings = 64;
constexpr int kPairConsumerWarps = 16;
constexpr int kPairThreads = kPairConsumerWarps * 32 + 128; // + the producer warpgroup (one live warp)

// One pass as kernel parameters (constant bank); masks in plan order, local coordinates in the re-chosen basis
template <typename T> struct PairStrings
{
    Cx<T> c[kPairMaxStrings];     // coefficient times (-i)^nY
    uint32_t z[kPairMaxStrings];  // full z-mask (<= 30 qubits): sign of a global row = parity(row & z)
    uint8_t gs[kPairMasks + 1];   // strings of mask g: gs[g] .. gs[g + 1]
    uint64_t basis[2][kPairMasks]; // the re-chosen basis (see above) for RB = 3 ([0]) and RB = 2 ([1]: two pairs)
};

template <typename T> constexpr size_t pair_table_bytes()
{
    return static_cast<size_t>(kPairMasks) * 256 * sizeof(Cx<T>);
}

// MODE 1: PauliOp::expectation_value partials (PO:482-549) instead of the store: e(t) += conj(psi(l, t)) (A psi)(l, t)
// with psi(l, t) read from the tile.  Per tile a warp folds its 16 rows (in-lane over the lane's rows, two shuffles over
// the lanes that share a column) and adds the column sums to ITS OWN row of `partials` ([CTA][warp][column], zeroed by
// the host) with fire-and-forget reductions: every address has exactly one writing thread, so the order of the additions
// is the program order and the result is reproducible; finalize_complex_kernel adds the rows in a fixed order.
template <typename T, int EPV, int RB = 2, int MODE = 0>
__global__ void __launch_bounds__(kPairThreads, 1)
    coset_pair_tma_kernel(uint64_t nonpivot_mask, uint64_t rowvecs, uint32_t nColTiles, uint64_t nTiles,
                          CVec<T, EPV> *__restrict__ out, int beta, const __grid_constant__ PairStrings<T> strs,
                          const __grid_constant__ CUtensorMap tm_in, Cx<T> *__restrict__ partials = nullptr,
                          uint32_t Bpad = 0)
{
    using Vec = CVec<T, EPV>;
    constexpr int TWC = 16, R = 8;
    constexpr uint32_t ROW_SHIFT = 8;

    extern __shared__ __align__(1024) unsigned char smem_pt[];
    __shared__ uint64_t s_full[kFewTmaBufs], s_empty[kFewTmaBufs];
    __shared__ uint32_t s_comb[256]; // XOR offsets of the 256 local rows
    Cx<T> *const tab = reinterpret_cast<Cx<T> *>(smem_pt + kFewTmaBufs * kFewTmaTile); // [mask][local row]

    uint32_t const tid = threadIdx.x;
    if (tid < 256)
        s_comb[tid] = static_cast<uint32_t>(comb_of<R>(strs.basis[RB == 3 ? 0 : 1], tid));
    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < kFewTmaBufs; ++b)
        {
            few_mbar_init(&s_full[b], 1);
            few_mbar_init(&s_empty[b], kPairConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's contiguous range of (coset, column tile) work items, column tiles fastest
    uint64_t const t0 = nTiles * blockIdx.x / gridDim.x, t1 = nTiles * (blockIdx.x + 1) / gridDim.x;

    if (tid >= kPairConsumerWarps * 32)
    {
        // ------------------------------------------------ producer warpgroup: hands its registers to the consumers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid >= kPairConsumerWarps * 32 + 32)
            return;
        uint32_t const lane = tid & 31u;
        uint64_t p_coset = 0;
        uint32_t p_ct = 0, p_base = 0;
        uint32_t buf = 0, round = 0; // ring position: tile i of this CTA sits in buffer i % 3, round = i / 3
        for (uint64_t t = t0; t < t1; ++t)
        {
            if (round)
                few_mbar_wait(&s_empty[buf], (round - 1) & 1u);
            if (t == t0 || ++p_ct == nColTiles)
            {
                p_coset = t == t0 ? t0 / nColTiles : p_coset + 1;
                p_ct = t == t0 ? static_cast<uint32_t>(t0 - p_coset * nColTiles) : 0u;
                p_base = t == t0 ? static_cast<uint32_t>(deposit_bits(p_coset, nonpivot_mask))
                                 : next_coset_base(p_base, static_cast<uint32_t>(nonpivot_mask));
            }
            if (lane == 0)
                few_mbar_expect_tx(&s_full[buf], static_cast<uint32_t>(kFewTmaTile));
            __syncwarp();
            int const c0 = static_cast<int>(p_ct) * TWC * static_cast<int>(16 / sizeof(T));
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                uint32_t const op = lane + 32 * h; // rows 4*op .. 4*op+3
                few_tma_gather4(smem_pt + buf * kFewTmaTile + (static_cast<size_t>(op) << (ROW_SHIFT + 2)), &tm_in, c0,
                                p_base ^ s_comb[4 * op], p_base ^ s_comb[4 * op + 1], p_base ^ s_comb[4 * op + 2],
                                p_base ^ s_comb[4 * op + 3], &s_full[buf]);
            }
            if (++buf == kFewTmaBufs)
            {
                buf = 0;
                ++round;
            }
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // A lane owns NR = 2^RB rows (the top RB local bits) x NV = 8 / NR vectors of the 256-byte row segment; LPR lanes
    // share a row, a warp instruction covers RW rows.  RB = 3: 8 rows x 1 vector, three mask pairs (40 gathers + 64 factor
    // loads per lane and tile); RB = 2: 4 rows x 2 vectors, two pairs (48 gathers + 32 factor loads).
    constexpr int NR = 1 << RB, NV = 8 / NR, LPR = 16 / NV, RW = 32 / LPR, RSTEP = RW * 16, NSINGLE = 8 - 2 * RB;
    uint32_t const warp = tid >> 5, lane = tid & 31u;
    uint32_t const rq = lane / LPR, jl = lane % LPR;
    uint32_t const l0 = RW * warp + rq;                      // local row of step i: l0 + RSTEP i
    uint32_t const own0 = (l0 << ROW_SHIFT) | (jl << 4);     // byte offset of (l0, this lane's first vector) in a buffer
    Cx<T> const *const my_tab = tab + l0;                    // factor of (mask g, step i): my_tab[g * 256 + RSTEP * i]
    // table builder: a warp forms exactly the 16 rows x 8 masks it reads itself (lane -> one row, masks g_lo .. g_lo + 3),
    // so the table needs no barrier among the warps and they keep drifting apart
    uint32_t const lb = RW * warp + ((lane >> 1) % RW) + RSTEP * ((lane >> 1) / RW), g_lo = (lane & 1u) * 4u;

    uint64_t coset = 0;
    uint32_t ct = 0, base = 0;
    uint32_t buf = 0, round = 0;
    for (uint64_t t = t0; t < t1; ++t)
    {
        if (t == t0 || ++ct == nColTiles)
        {
            coset = t == t0 ? t0 / nColTiles : coset + 1;
            ct = t == t0 ? static_cast<uint32_t>(t0 - coset * nColTiles) : 0u;
            base = t == t0 ? static_cast<uint32_t>(deposit_bits(coset, nonpivot_mask)) // launched for <= 30 qubits
                           : next_coset_base(base, static_cast<uint32_t>(nonpivot_mask));
            // new coset: row factors D_g(l) = sum_{s in g} (-1)^{popc(row & z_s)} c_s, strings in plan order
            __syncwarp(); // the warp's slice of the old table has been read
            uint32_t const row = base ^ s_comb[lb];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
            {
                uint32_t const g = g_lo + k;
                Cx<T> d{0, 0};
                uint32_t const s1 = strs.gs[g + 1];
                for (uint32_t s = strs.gs[g]; s < s1; ++s)
                {
                    Cx<T> const c = strs.c[s];
                    uint32_t const odd = __popc(row & strs.z[s]) & 1u;
                    d.re += flip_sign(c.re, odd);
                    d.im += flip_sign(c.im, odd);
                }
                tab[g * 256u + lb] = d;
            }
            __syncwarp();
        }
        uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jl;

        if (MODE == 0 && beta && t + 1 < t1)
        {
            // accumulating pass: the next tile's old output rows -> L2, a whole tile of gathers away from their use
            bool const same = ct + 1 < nColTiles;
            uint32_t const base_n = same ? base : next_coset_base(base, static_cast<uint32_t>(nonpivot_mask));
            uint64_t const vcol_n = static_cast<uint64_t>(same ? ct + 1 : 0u) * TWC + jl;
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(
                        &out[static_cast<uint64_t>(base_n ^ s_comb[l0 + RSTEP * i]) * rowvecs + vcol_n + c * LPR]));
        }
        few_mbar_wait(&s_full[buf], round & 1u);
        unsigned char const *const tb = smem_pt + buf * kFewTmaTile;

        Cx<T> acc[NR][NV][EPV];
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int c = 0; c < NV; ++c)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[i][c][e] = Cx<T>{0, 0};
        Vec v[NR][NV];
        // pairs (2p, 2p + 1): local x = e_p and e_p ^ e_{8 - RB + p}
#pragma unroll
        for (int p = 0; p < RB; ++p)
        {
            uint32_t const off = own0 ^ (1u << (ROW_SHIFT + p));
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    v[i][c] = *reinterpret_cast<Vec const *>(tb + off + i * (RSTEP << ROW_SHIFT) + c * (LPR * 16));
#pragma unroll
            for (int i = 0; i < NR; ++i)
            {
                Cx<T> const f = my_tab[(2 * p) * 256 + RSTEP * i];
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[i][c][e], f, v[i][c].e[e]);
            }
#pragma unroll
            for (int i = 0; i < NR; ++i)
            {
                Cx<T> const f = my_tab[(2 * p + 1) * 256 + RSTEP * i];
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[i][c][e], f, v[i ^ (1 << p)][c].e[e]);
            }
        }
        // single masks 2 RB .. 7: local x = e_{RB + q}
#pragma unroll
        for (int q = 0; q < NSINGLE; ++q)
        {
            uint32_t const off = own0 ^ (1u << (ROW_SHIFT + RB + q));
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    v[i][c] = *reinterpret_cast<Vec const *>(tb + off + i * (RSTEP << ROW_SHIFT) + c * (LPR * 16));
#pragma unroll
            for (int i = 0; i < NR; ++i)
            {
                Cx<T> const f = my_tab[(2 * RB + q) * 256 + RSTEP * i];
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[i][c][e], f, v[i][c].e[e]);
            }
        }
        Cx<T> esum[NV][EPV];
        if (MODE == 1)
        {
            // conj(psi(l, t)) * (A psi)(l, t), summed over the lane's rows; psi of the lane's own rows from the tile
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    v[i][c] = *reinterpret_cast<Vec const *>(tb + own0 + i * (RSTEP << ROW_SHIFT) + c * (LPR * 16));
#pragma unroll
            for (int c = 0; c < NV; ++c)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    Cx<T> sacc{0, 0};
#pragma unroll
                    for (int i = 0; i < NR; ++i)
                    {
                        Cx<T> const a = v[i][c].e[e];
                        sacc.re = fma(a.re, acc[i][c][e].re, sacc.re);
                        sacc.re = fma(a.im, acc[i][c][e].im, sacc.re);
                        sacc.im = fma(a.re, acc[i][c][e].im, sacc.im);
                        sacc.im = fma(-a.im, acc[i][c][e].re, sacc.im);
                    }
                    esum[c][e] = sacc;
                }
        }
        // this warp's gathers of the buffer are done: order them before the asynchronous-proxy refill and hand it back
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            few_mbar_arrive(&s_empty[buf]);
        if (++buf == kFewTmaBufs)
        {
            buf = 0;
            ++round;
        }

        if (MODE == 1)
        {
            // fold over the RW lanes that share a column, then one writer per (warp, column)
#pragma unroll
            for (int c = 0; c < NV; ++c)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
#pragma unroll
                    for (int sh = LPR; sh < 32; sh <<= 1)
                    {
                        esum[c][e].re += __shfl_xor_sync(0xffffffffu, esum[c][e].re, sh);
                        esum[c][e].im += __shfl_xor_sync(0xffffffffu, esum[c][e].im, sh);
                    }
            if (rq == 0)
            {
                Cx<T> *const prow = partials + (static_cast<uint64_t>(blockIdx.x) * kPairConsumerWarps + warp) * Bpad;
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        Cx<T> *const dst = prow + (vcol + c * LPR) * EPV + e;
                        atomicAdd(&dst->re, esum[c][e].re);
                        atomicAdd(&dst->im, esum[c][e].im);
                    }
            }
            continue;
        }
        if (beta)
        {
            // old output rows (in L2 since the previous tile) into the registers the gathers have left
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    v[i][c] = out[static_cast<uint64_t>(base ^ s_comb[l0 + RSTEP * i]) * rowvecs + vcol + c * LPR];
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        acc[i][c][e].re += v[i][c].e[e].re;
                        acc[i][c][e].im += v[i][c].e[e].im;
                    }
        }
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int c = 0; c < NV; ++c)
            {
                Vec r;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    r.e[e] = acc[i][c][e];
                out[static_cast<uint64_t>(base ^ s_comb[l0 + RSTEP * i]) * rowvecs + vcol + c * LPR] = r;
            }
    }
}

} // namespace fpk
