// K3c: register-resident coset kernels for operators whose x-masks span a GF(2) subspace of rank <= 4.
//
// Same algebra as the shared-memory coset kernels (coset_plan.hpp): with the basis b_0..b_{RR-1} of the x-mask
// span in reduced echelon form, the 2^RR rows {base ^ comb(l)} of one coset are mapped onto themselves by every
// gather of the operator, row ^ x_g = row(l ^ xl_g).  Here the coset is small enough that ONE THREAD holds all of
// it for one 16-byte vector of the batch axis:
//
//   load   2^RR vectors, one per coset row; the lanes of a warp run along the batch axis, so every load/store
//          instruction moves one contiguous row segment (512 B per warp): no shared-memory staging at all
//   table  the CTA's threads fill D[coset][xl][l] = sum_{s: x_s = xl} c_s (-1)^{par(base & z_s) ^ par(l & zl_s)}
//          in shared memory while the loads above are in flight: strings are first summed per local z-mask (one
//          pass over the strings per coset), then a 2^RR-point sign transform gives the 2^RR rows
//   fma    acc[l] += D[xl][l] * psi[l ^ xl] for every present xl: the gather is a compile-time register
//          permutation (the loop over xl is fully unrolled, absent masks are skipped by a uniform branch), the
//          factors are warp-broadcast LDS
//   store  2^RR vectors straight from registers (MODE 0), or conj(psi) . acc reduced per column (MODE 1)
//
// HBM traffic is the compulsory read + write and nothing else touches the memory pipes, so few-group operators
// run at streaming-copy speed (the shared-memory coset kernel spends G x 64 KiB of LDS per 64 KiB tile instead).
//
// MODE 0: PauliOp::apply (PO:399-468) / SummedPauliOp::apply (SPO:277-349)
// MODE 1: PauliOp::expectation_value partials (PO:482-549)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace fpk
{

constexpr int kRcMaxRank = 5;      // plans up to rank 5 (dcoset.cuh); the SIMT kernels below serve ranks <= 4
constexpr int kRcMaxSimtRank = 4;

template <typename T> struct RcPassView
{
    uint64_t basis[kRcMaxRank];
    uint32_t pivot[kRcMaxRank]; // ascending pivot bit positions of the basis
    uint32_t present;           // bit xl set <=> some string gathers with local mask xl
    uint32_t const *ustart;     // [4^RR + 1] strings sorted by (xl, zl): those with local masks (xl, zl) are
                                //            [ustart[(xl << RR) + zl], ustart[(xl << RR) + zl + 1])
    uint64_t const *sz;         // [S] full z-mask (coset-base parity)
    Cx<T> const *scoef;         // [S] h_s (-i)^nY_s
};

// coset number -> base row: spread the bits of c over the non-pivot positions (zeros at the pivots)
template <int RR> __device__ __forceinline__ uint64_t rc_base(uint64_t c, uint32_t const *pivot)
{
#pragma unroll
    for (int k = 0; k < RR; ++k)
    {
        uint32_t const p = pivot[k];
        uint64_t const low = c & ((1ull << p) - 1ull);
        c = ((c >> p) << (p + 1)) | low;
    }
    return c;
}

template <int RR> __device__ __forceinline__ uint64_t rc_comb(uint64_t const *basis, int l)
{
    uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < RR; ++k)
        if ((l >> k) & 1)
            c ^= basis[k];
    return c;
}

template <typename T, int RR> constexpr size_t rc_smem_bytes(uint32_t TY)
{
    return static_cast<size_t>(TY) * (1u << (2 * RR)) * sizeof(Cx<T>);
}

// Thread geometry: TW = 2^log2TW lanes along the batch axis x TY = NT / TW cosets per CTA iteration; the grid
// enumerates column tiles fastest.  MODE 1 CTAs walk `iters` consecutive coset sets and emit one partial row.
template <typename T, int EPV, int RR, int LOG_NT, int MODE>
__global__ void __launch_bounds__(1 << LOG_NT, (RR == 4 ? 512 : RR == 3 ? 640 : 1024) >> LOG_NT)
    rcoset_kernel(RcPassView<T> pass, uint64_t n_cosets, uint64_t rowvecs, uint32_t log2TW, uint32_t log2P,
                  uint32_t nColTiles, uint32_t iters, CVec<T, EPV> const *__restrict__ in, CVec<T, EPV> *__restrict__ out, int beta,
                  Cx<T> *__restrict__ partials, uint32_t Bpad)
{
    using Vec = CVec<T, EPV>;
    constexpr int NT = 1 << LOG_NT, ROWS = 1 << RR;
    extern __shared__ __align__(16) unsigned char rc_smem[];
    Cx<T> *Dt = reinterpret_cast<Cx<T> *>(rc_smem); // [TY][xl][l]
    Cx<T> *Ut = Dt + ((NT >> log2TW) << (2 * RR));  // [TY][xl][zl]

    uint32_t const tid = threadIdx.x;
    uint32_t const TW = 1u << log2TW, TY = NT >> log2TW;
    uint32_t const lx = tid & (TW - 1), ty = tid >> log2TW;
    uint64_t const cb = blockIdx.x / nColTiles;
    uint32_t const ct = static_cast<uint32_t>(blockIdx.x - cb * nColTiles);
    uint64_t const v = static_cast<uint64_t>(ct) * TW + lx;
    bool const vok = v < rowvecs;

    Cx<T> esum[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e)
        esum[e] = Cx<T>{0, 0};

    for (uint32_t it = 0; it < iters; ++it)
    {
        uint64_t const cs0 = (cb * iters + it) * TY;
        if (cs0 >= n_cosets)
            break; // uniform
        uint64_t const coset = cs0 + ty;
        bool const live = vok && coset < n_cosets;
        uint64_t const base = rc_base<RR>(coset, pass.pivot);

        Vec x[ROWS];
#pragma unroll
        for (int l = 0; l < ROWS; ++l)
        {
            if (live)
                x[l] = in[(base ^ rc_comb<RR>(pass.basis, l)) * rowvecs + v];
            else
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    x[l].e[e] = Cx<T>{0, 0};
            }
        }

        // row-only factors of this iteration's cosets (the loads above are still in flight), in two steps:
        //   U[c][xl][zl] = sum over the strings with local masks (xl, zl) of c_s (-1)^{par(base_c & z_s)}
        //   D[c][xl][l]  = sum_zl (-1)^{popc(l & zl)} U[c][xl][zl]
        // every string is touched once per coset (not once per coset row) and the second step is a tiny
        // fixed-size transform
        {
            // 2^log2P consecutive lanes share one U entry (small tables: keeps every thread busy)
            uint32_t const P = 1u << log2P, part = tid & (P - 1);
            for (uint32_t ent = tid >> log2P; ent < TY * ROWS * ROWS; ent += NT >> log2P)
            {
                uint32_t const c = ent >> (2 * RR), xz = ent & (ROWS * ROWS - 1);
                Cx<T> u{0, 0};
                uint32_t const s0 = __ldg(pass.ustart + xz), s1 = __ldg(pass.ustart + xz + 1);
                if (s0 < s1)
                {
                    uint64_t const b = rc_base<RR>(cs0 + c, pass.pivot);
                    for (uint32_t s = s0 + part; s < s1; s += P)
                    {
                        Cx<T> const cf = pass.scoef[s];
                        uint32_t const odd = parity64(b & __ldg(pass.sz + s));
                        u.re += flip_sign(cf.re, odd);
                        u.im += flip_sign(cf.im, odd);
                    }
                }
                for (uint32_t off = P >> 1; off > 0; off >>= 1)
                {
                    u.re += __shfl_xor_sync(0xffffffffu, u.re, off);
                    u.im += __shfl_xor_sync(0xffffffffu, u.im, off);
                }
                if (part == 0)
                    Ut[ent] = u;
            }
        }
        __syncthreads();
        for (uint32_t ent = tid; ent < TY * ROWS * ROWS; ent += NT)
        {
            uint32_t const xl = (ent >> RR) & (ROWS - 1), l = ent & (ROWS - 1);
            if ((pass.present >> xl) & 1u)
            {
                Cx<T> const *Ux = Ut + (ent & ~static_cast<uint32_t>(ROWS - 1));
                Cx<T> d{0, 0};
#pragma unroll
                for (int zl = 0; zl < ROWS; ++zl)
                {
                    Cx<T> const u = Ux[zl];
                    uint32_t const odd = __popc(l & zl) & 1u;
                    d.re += flip_sign(u.re, odd);
                    d.im += flip_sign(u.im, odd);
                }
                Dt[ent] = d;
            }
        }
        __syncthreads();

        // the 16-row case evaluates its output rows in two halves: 32 fewer accumulator registers buy a fourth
        // resident CTA per SM
        constexpr int SPLIT = RR == 4 ? 2 : 1, HR = ROWS / SPLIT;
        Cx<T> const *Dc = Dt + (ty << (2 * RR));
#pragma unroll
        for (int h = 0; h < SPLIT; ++h)
        {
            Cx<T> acc[HR][EPV];
#pragma unroll
            for (int l = 0; l < HR; ++l)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[l][e] = Cx<T>{0, 0};
#pragma unroll
            for (int xl = 0; xl < ROWS; ++xl)
            {
                if ((pass.present >> xl) & 1u)
                {
#pragma unroll
                    for (int lh = 0; lh < HR; ++lh)
                    {
                        int const l = h * HR + lh;
                        Cx<T> const d = Dc[xl * ROWS + l];
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            cfma(acc[lh][e], d, x[l ^ xl].e[e]);
                    }
                }
            }

            if (MODE == 0)
            {
                if (live)
                {
#pragma unroll
                    for (int lh = 0; lh < HR; ++lh)
                    {
                        int const l = h * HR + lh;
                        Vec *dst = &out[(base ^ rc_comb<RR>(pass.basis, l)) * rowvecs + v];
                        Vec r;
                        if (beta)
                        {
                            r = *dst;
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                            {
                                r.e[e].re += acc[lh][e].re;
                                r.e[e].im += acc[lh][e].im;
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                                r.e[e] = acc[lh][e];
                        }
                        *dst = r;
                    }
                }
            }
            else
            {
#pragma unroll
                for (int lh = 0; lh < HR; ++lh)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        Cx<T> const a = x[h * HR + lh].e[e]; // conj(a) * acc (dead lanes hold zeros)
                        esum[e].re = fma(a.re, acc[lh][e].re, esum[e].re);
                        esum[e].re = fma(a.im, acc[lh][e].im, esum[e].re);
                        esum[e].im = fma(a.re, acc[lh][e].im, esum[e].im);
                        esum[e].im = fma(-a.im, acc[lh][e].re, esum[e].im);
                    }
            }
        }
        __syncthreads(); // the table is rebuilt by the next iteration
    }

    if (MODE == 1)
    {
        // reduce over the TY cosets that share a vector column; the table area is free now
        Cx<T> *red = Dt; // needs NT * EPV entries: the host sizes the allocation for both uses
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            red[tid * EPV + e] = esum[e];
        __syncthreads();
        for (uint32_t half = TY >> 1; half > 0; half >>= 1)
        {
            if (ty < half)
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    Cx<T> const o = red[(tid + half * TW) * EPV + e];
                    red[tid * EPV + e].re += o.re;
                    red[tid * EPV + e].im += o.im;
                }
            }
            __syncthreads();
        }
        if (ty == 0 && vok)
        {
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                partials[cb * Bpad + v * EPV + e] = red[tid * EPV + e];
        }
    }
}

} // namespace fpk
