// K6b: SummedPauliOp::apply_weighted second stage (SPO:441-455) for registers of 11 qubits and more (complex64 below,
// complex128 at the end of the file):
//
//     out(l, t) (+)= sum_g [ sum_{s in g} (-1)^{popc(l & z_s)} W(s, t) ] * psi(l ^ x_g, t)
//
// One CTA owns a whole state column pair (2^n rows x 2 complex64 columns = one 16-byte vector per row, <= 64 KiB) in
// shared memory and evaluates every string against it -- a single pass whatever the operator (n <= 12); beyond 12
// qubits the tile is one rank-12 coset of a pass of the coset plan instead of the whole column.  Against the generic
// MODE 2 coset kernel this one removes the instruction overhead that bounded it (half its issue slots were sign
// generation and W address arithmetic):
//
//   * the tile is stored planar per pair, (re0, re1, im0, im1), so every FMA is a packed FFMA2 over the two columns;
//   * W(s, t0..t0+1) of a 128-string chunk is staged once per CTA (prefetched into registers during the previous
//     chunk) and read as one broadcast LDS.128 per string;
//   * a thread owns rows tid + q*NT, q < 8; the sign pattern of a string over q depends only on the string's three
//     top z bits, so a warp-uniform 8-way switch selects a fully unrolled body whose +-1 factors are compile-time
//     choices between the string's signed W pair and its negation: no per-row integer work at all, one FFMA2 per row and
//     component (the thread's own parity bit costs 4 LOP3 per string; FFMA2 negates its operand for free).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "coset.cuh"

namespace fpk
{

constexpr int kWtRpt = 8; // rows per thread (three top row bits)

template <int LOG_NT> struct WtileSmem
{
    static constexpr size_t tile = (static_cast<size_t>(kWtRpt) << LOG_NT) * 16;
    static constexpr size_t off_w = tile;                                      // float4 [CH]
    static constexpr size_t off_meta = off_w + kCosetChunkStrings * 16;        // uint32 [CH]
    static constexpr size_t off_gxl = off_meta + kCosetChunkStrings * 4;       // uint32 [CH]
    static constexpr size_t off_gstart = off_gxl + kCosetChunkGroups * 4;      // uint32 [CH + 2]
    static constexpr size_t bytes = off_gstart + (kCosetChunkGroups + 2) * 4 + 8;
};

#define FP_WT_ROW(K, Q)                                                                                                \
    {                                                                                                                  \
        bool const neg = __builtin_popcount((Q) & (K)) & 1;                                                            \
        dre[Q] = __ffma2_rn(neg ? make_float2(-wr.x, -wr.y) : wr, one2, dre[Q]);                                       \
        dim[Q] = __ffma2_rn(neg ? make_float2(-wi.x, -wi.y) : wi, one2, dim[Q]);                                       \
    }
#define FP_WT_CASE(K)                                                                                                  \
    case K:                                                                                                            \
        FP_WT_ROW(K, 0) FP_WT_ROW(K, 1) FP_WT_ROW(K, 2) FP_WT_ROW(K, 3) FP_WT_ROW(K, 4) FP_WT_ROW(K, 5)               \
        FP_WT_ROW(K, 6) FP_WT_ROW(K, 7) break;

// grid.x = (cosets of the pass) x (B / 2 column pairs), column pairs fastest.  For registers of at most 12 qubits the
// tile is the whole state column and the plan has one pass; larger registers are covered coset by coset (rank-12
// cosets of the pass' x-mask span, rows base ^ comb(basis, l)), one pass per launch, later passes accumulating.
template <int LOG_NT>
__global__ void __launch_bounds__(1 << LOG_NT, 1)
    wtile_kernel(CosetPassView<float> pass, uint64_t rowvecs, CVec<float, 2> const *__restrict__ in,
                 CVec<float, 2> *__restrict__ out, int beta, float const *__restrict__ Wre,
                 float const *__restrict__ Wim, uint64_t B)
{
    constexpr int NT = 1 << LOG_NT, RPT = kWtRpt;
    using S = WtileSmem<LOG_NT>;
    extern __shared__ __align__(16) unsigned char wt_smem[];
    float4 *tile = reinterpret_cast<float4 *>(wt_smem);
    float4 *s_w = reinterpret_cast<float4 *>(wt_smem + S::off_w);
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(wt_smem + S::off_meta);
    uint32_t *s_gxl = reinterpret_cast<uint32_t *>(wt_smem + S::off_gxl);
    uint32_t *s_gstart = reinterpret_cast<uint32_t *>(wt_smem + S::off_gstart);

    uint32_t const tid = threadIdx.x;
    uint64_t const coset = blockIdx.x / rowvecs;
    uint64_t const v = blockIdx.x - coset * rowvecs;
    uint64_t const t0 = v * 2;
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
    float4 const *in4 = reinterpret_cast<float4 const *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    uint64_t grow[RPT]; // global row of the thread's local rows tid + q * NT
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        grow[q] = base ^ comb_of<LOG_NT + 3>(pass.basis, tid + q * NT);

    uint32_t rowoff[RPT]; // byte offset of the thread's rows inside the tile
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        rowoff[q] = (tid + q * NT) * 16u;
    // state column pair -> shared memory, planar per pair: (re0, re1, im0, im1)
#pragma unroll
    for (int q = 0; q < RPT; ++q)
    {
        uint32_t const l = tid + q * NT;
        float4 const a = in4[grow[q] * rowvecs + v];
        tile[l] = make_float4(a.x, a.z, a.y, a.w);
    }

    // chunk metadata travels global -> registers (during the previous chunk's arithmetic) -> shared memory
    uint32_t r_meta = 0, r_gxl = 0, r_gstart = 0;
    float4 r_w = make_float4(0, 0, 0, 0);
    auto fetch = [&](uint32_t ci) {
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ns = ch.s_hi - ch.s_lo, ng = ch.g_hi - ch.g_lo;
        if (tid < ns)
        {
            uint32_t const zl = pass.szl[ch.s_lo + tid];
            r_meta = (zl & (NT - 1)) | ((zl >> LOG_NT) << 16);
            uint64_t const wrow = static_cast<uint64_t>(pass.sidx[ch.s_lo + tid]) * B + t0;
            float2 const re = *reinterpret_cast<float2 const *>(Wre + wrow);
            float2 const im = *reinterpret_cast<float2 const *>(Wim + wrow);
            // the sign of the coset base, (-1)^{par(base & z)}, is folded into the staged weights
            float const hs = parity64(base & pass.sz[ch.s_lo + tid]) ? -1.f : 1.f;
            r_w = make_float4(hs * re.x, hs * re.y, hs * im.x, hs * im.y);
        }
        if (tid <= ng)
        {
            r_gstart = pass.gstart[ch.g_lo + tid] - ch.s_lo;
            if (tid < ng)
                r_gxl = pass.gxl[ch.g_lo + tid];
        }
    };

    float2 acc_rp[RPT], acc_rm[RPT], acc_im[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        acc_rp[q] = acc_rm[q] = acc_im[q] = make_float2(0.f, 0.f);

    float2 const one2 = make_float2(1.f, 1.f);
    fetch(0);
    for (uint32_t ci = 0; ci < pass.n_chunks; ++ci)
    {
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ng = ch.g_hi - ch.g_lo;
        __syncthreads(); // the previous chunk's readers are done (first chunk: nothing to wait for)
        if (tid < kCosetChunkStrings)
        {
            s_meta[tid] = r_meta;
            s_w[tid] = r_w;
            s_gxl[tid] = r_gxl;
        }
        if (tid <= kCosetChunkGroups)
            s_gstart[tid] = r_gstart;
        __syncthreads(); // (first chunk: also publishes the tile)
        if (ci + 1 < pass.n_chunks)
            fetch(ci + 1);

        for (uint32_t gq = 0; gq < ng; ++gq)
        {
            uint32_t const xl = s_gxl[gq];
            uint32_t const s0 = s_gstart[gq], s1 = s_gstart[gq + 1];
            float2 dre[RPT], dim[RPT];
#pragma unroll
            for (int q = 0; q < RPT; ++q)
                dre[q] = dim[q] = make_float2(0.f, 0.f);
            for (uint32_t s = s0; s < s1; ++s)
            {
                uint32_t const m = s_meta[s];
                float4 const w = s_w[s];
                uint32_t const sgn = (__popc(tid & m & (NT - 1)) & 1u) << 31;
                float2 const wr = make_float2(__uint_as_float(__float_as_uint(w.x) ^ sgn),
                                              __uint_as_float(__float_as_uint(w.y) ^ sgn));
                float2 const wi = make_float2(__uint_as_float(__float_as_uint(w.z) ^ sgn),
                                              __uint_as_float(__float_as_uint(w.w) ^ sgn));
                switch ((m >> 16) & 7u) // warp-uniform: the three top z bits of the string
                {
                    FP_WT_CASE(0)
                    FP_WT_CASE(1)
                    FP_WT_CASE(2)
                    FP_WT_CASE(3)
                    FP_WT_CASE(4)
                    FP_WT_CASE(5)
                    FP_WT_CASE(6)
                    FP_WT_CASE(7)
                }
            }
            uint32_t const xl4 = xl << 4;
#pragma unroll
            for (int q = 0; q < RPT; ++q)
            {
                float4 const a = *reinterpret_cast<float4 const *>(wt_smem + (rowoff[q] ^ xl4));
                float2 const vr = make_float2(a.x, a.y), vi = make_float2(a.z, a.w);
                acc_rp[q] = __ffma2_rn(dre[q], vr, acc_rp[q]);
                acc_rm[q] = __ffma2_rn(dim[q], vi, acc_rm[q]);
                acc_im[q] = __ffma2_rn(dre[q], vi, acc_im[q]);
                acc_im[q] = __ffma2_rn(dim[q], vr, acc_im[q]);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < RPT; ++q)
    {
        float4 r = make_float4(acc_rp[q].x - acc_rm[q].x, acc_im[q].x, acc_rp[q].y - acc_rm[q].y, acc_im[q].y);
        float4 *dst = &out4[grow[q] * rowvecs + v];
        if (beta)
        {
            float4 const o = *dst;
            r.x += o.x;
            r.y += o.y;
            r.z += o.z;
            r.w += o.w;
        }
        *dst = r;
    }
}

// ---------------------------------------------------------------- complex128 variant
// Same plan with one complex128 column per CTA (a row is one 16-byte vector (re, im)): scalar FP64 arithmetic, the
// 8-way sign-pattern switch selects 16 DADD with free operand negation, the accumulation is 4 DFMA per row and group.
// (The reference's Python bindings are complex128-only, so this is the path a Python SummedPauliOp user takes.)
#define FP_WTD_ROW(K, Q)                                                                                               \
    {                                                                                                                  \
        bool const neg = __builtin_popcount((Q) & (K)) & 1;                                                            \
        dre[Q] += neg ? -wr : wr;                                                                                      \
        dim[Q] += neg ? -wi : wi;                                                                                      \
    }
#define FP_WTD_CASE(K)                                                                                                 \
    case K:                                                                                                            \
        FP_WTD_ROW(K, 0) FP_WTD_ROW(K, 1) FP_WTD_ROW(K, 2) FP_WTD_ROW(K, 3) FP_WTD_ROW(K, 4) FP_WTD_ROW(K, 5)         \
        FP_WTD_ROW(K, 6) FP_WTD_ROW(K, 7) break;

template <int LOG_NT>
__global__ void __launch_bounds__(1 << LOG_NT, 1)
    wtile_f64_kernel(CosetPassView<double> pass, uint64_t rowvecs, CVec<double, 1> const *__restrict__ in,
                     CVec<double, 1> *__restrict__ out, int beta, double const *__restrict__ Wre,
                     double const *__restrict__ Wim, uint64_t B)
{
    constexpr int NT = 1 << LOG_NT, RPT = kWtRpt;
    using S = WtileSmem<LOG_NT>;
    extern __shared__ __align__(16) unsigned char wtd_smem[];
    double2 *tile = reinterpret_cast<double2 *>(wtd_smem);
    double2 *s_w = reinterpret_cast<double2 *>(wtd_smem + S::off_w);
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(wtd_smem + S::off_meta);
    uint32_t *s_gxl = reinterpret_cast<uint32_t *>(wtd_smem + S::off_gxl);
    uint32_t *s_gstart = reinterpret_cast<uint32_t *>(wtd_smem + S::off_gstart);

    uint32_t const tid = threadIdx.x;
    uint64_t const coset = blockIdx.x / rowvecs;
    uint64_t const v = blockIdx.x - coset * rowvecs; // batch column
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
    double2 const *in2 = reinterpret_cast<double2 const *>(in);
    double2 *out2 = reinterpret_cast<double2 *>(out);
    uint64_t grow[RPT]; // global row of the thread's local rows tid + q * NT
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        grow[q] = base ^ comb_of<LOG_NT + 3>(pass.basis, tid + q * NT);

    uint32_t rowoff[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        rowoff[q] = (tid + q * NT) * 16u;
#pragma unroll
    for (int q = 0; q < RPT; ++q)
    {
        uint32_t const l = tid + q * NT;
        tile[l] = in2[grow[q] * rowvecs + v];
    }

    uint32_t r_meta = 0, r_gxl = 0, r_gstart = 0;
    double2 r_w = make_double2(0, 0);
    auto fetch = [&](uint32_t ci) {
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ns = ch.s_hi - ch.s_lo, ng = ch.g_hi - ch.g_lo;
        if (tid < ns)
        {
            uint32_t const zl = pass.szl[ch.s_lo + tid];
            r_meta = (zl & (NT - 1)) | ((zl >> LOG_NT) << 16);
            uint64_t const wrow = static_cast<uint64_t>(pass.sidx[ch.s_lo + tid]) * B + v;
            double const hs = parity64(base & pass.sz[ch.s_lo + tid]) ? -1.0 : 1.0; // coset-base sign
            r_w = make_double2(hs * Wre[wrow], hs * Wim[wrow]);
        }
        if (tid <= ng)
        {
            r_gstart = pass.gstart[ch.g_lo + tid] - ch.s_lo;
            if (tid < ng)
                r_gxl = pass.gxl[ch.g_lo + tid];
        }
    };

    double acc_re[RPT], acc_im[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
        acc_re[q] = acc_im[q] = 0.0;

    fetch(0);
    for (uint32_t ci = 0; ci < pass.n_chunks; ++ci)
    {
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ng = ch.g_hi - ch.g_lo;
        __syncthreads();
        if (tid < kCosetChunkStrings)
        {
            s_meta[tid] = r_meta;
            s_w[tid] = r_w;
            s_gxl[tid] = r_gxl;
        }
        if (tid <= kCosetChunkGroups)
            s_gstart[tid] = r_gstart;
        __syncthreads();
        if (ci + 1 < pass.n_chunks)
            fetch(ci + 1);

        for (uint32_t gq = 0; gq < ng; ++gq)
        {
            uint32_t const xl = s_gxl[gq];
            uint32_t const s0 = s_gstart[gq], s1 = s_gstart[gq + 1];
            double dre[RPT], dim[RPT];
#pragma unroll
            for (int q = 0; q < RPT; ++q)
                dre[q] = dim[q] = 0.0;
            for (uint32_t s = s0; s < s1; ++s)
            {
                uint32_t const m = s_meta[s];
                double2 const w = s_w[s];
                uint32_t const odd = __popc(tid & m & (NT - 1)) & 1u;
                double const wr = flip_sign(w.x, odd), wi = flip_sign(w.y, odd);
                switch ((m >> 16) & 7u) // warp-uniform: the three top z bits of the string
                {
                    FP_WTD_CASE(0)
                    FP_WTD_CASE(1)
                    FP_WTD_CASE(2)
                    FP_WTD_CASE(3)
                    FP_WTD_CASE(4)
                    FP_WTD_CASE(5)
                    FP_WTD_CASE(6)
                    FP_WTD_CASE(7)
                }
            }
            uint32_t const xl4 = xl << 4;
#pragma unroll
            for (int q = 0; q < RPT; ++q)
            {
                double2 const a = *reinterpret_cast<double2 const *>(wtd_smem + (rowoff[q] ^ xl4));
                acc_re[q] = fma(dre[q], a.x, acc_re[q]);
                acc_re[q] = fma(-dim[q], a.y, acc_re[q]);
                acc_im[q] = fma(dre[q], a.y, acc_im[q]);
                acc_im[q] = fma(dim[q], a.x, acc_im[q]);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < RPT; ++q)
    {
        double2 r = make_double2(acc_re[q], acc_im[q]);
        double2 *dst = &out2[grow[q] * rowvecs + v];
        if (beta)
        {
            double2 const o = *dst;
            r.x += o.x;
            r.y += o.y;
        }
        *dst = r;
    }
}

#undef FP_WTD_CASE
#undef FP_WTD_ROW
#undef FP_WT_CASE
#undef FP_WT_ROW

} // namespace fpk
