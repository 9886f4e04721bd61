// K3b / K6b: coset-blocked apply kernels -- the state tile lives in shared memory.
//
// One CTA owns the 2^r rows of one coset of the pass' x-mask subspace (see coset_plan.hpp) times a tile of TWc
// 16-byte vectors of the batch axis (4096 vectors = 64 KiB in every configuration):
//
//   load   2^r row segments, cp.async 16 B per lane, coalesced TWc*16-byte segments           (HBM -> smem, once)
//   loop   over the pass' groups: every thread owns RPT = 2^r/256 rows x TWc vectors; per group it forms the
//          row-only factor D_g(l) = sum_s c_s (-1)^par(l & zl_s) once per row and does TWc complex FMAs against
//          the gathered row (l ^ xl_g) read from shared memory                                  (smem only)
//   store  accumulators -> smem -> coalesced 16-byte stores (read-modify-write when accumulating) (smem -> HBM, once)
//
// Shared-memory rows are padded by one vector so that the 8 lanes of a quarter-warp, which read 8 different rows at
// the same column, hit 8 different 16-byte bank groups.  String metadata is staged per chunk with the coset-base
// sign par(base & z_s) already folded into the coefficient.
//
// MODE 0: PauliOp::apply (PO:399-468) / SummedPauliOp::apply (SPO:277-349)
// MODE 1: PauliOp::expectation_value partials (PO:482-549): e(t) += sum_l conj(psi(l,t)) (A psi)(l,t)
// MODE 2: SummedPauliOp::apply_weighted second stage (SPO:441-455): the per-string factor is W(s,t), read from the
//         planar contraction result, so D depends on the column as well.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "coset_plan.hpp"
#include "kernels.cuh"

namespace fpk
{

template <typename T> struct CosetPassView
{
    uint64_t basis[kCosetMaxRank];
    uint64_t nonpivot_mask;
    CosetChunk const *chunks;
    uint32_t const *gxl;
    uint32_t const *gstart;
    uint32_t const *szl;
    uint64_t const *sz;
    Cx<T> const *scoef;
    uint32_t const *sidx;
    uint32_t n_chunks;
};

__device__ __forceinline__ uint64_t deposit_bits(uint64_t src, uint64_t mask)
{
    uint64_t res = 0;
    for (uint64_t bb = 1; mask; bb <<= 1)
    {
        uint64_t low = mask & (~mask + 1);
        if (src & bb)
            res |= low;
        mask &= mask - 1;
    }
    return res;
}

template <int R> __device__ __forceinline__ uint64_t comb_of(uint64_t const *basis, uint32_t l)
{
    uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < R; ++k)
        if ((l >> k) & 1u)
            c ^= basis[k];
    return c;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, void const *gmem_src)
{
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int LOG_TWC> struct CosetCfg
{
    static constexpr int TWC = 1 << LOG_TWC;           // vectors per row in the tile
    static constexpr int RPT = 16 >> LOG_TWC;          // rows per thread
    static constexpr int R = 12 - LOG_TWC;             // tile rank: 2^R rows
    static constexpr int ROWS = 1 << R;
    static constexpr int PITCH = TWC + (TWC > 1 ? 1 : 0); // row pitch in vectors (padded)
    static constexpr size_t TILE_BYTES = static_cast<size_t>(ROWS) * PITCH * 16;
};

template <typename T> constexpr size_t coset_meta_bytes()
{
    // s_c[CH_S] + s_zl[CH_S] + s_sidx[CH_S] + s_gxl[CH_G] + s_gstart[CH_G+1] + comb_hi[16]
    return kCosetChunkStrings * (sizeof(Cx<T>) + 4 + 4) + kCosetChunkGroups * 4 + (kCosetChunkGroups + 1) * 4 + 16 * 8 +
           64;
}

template <typename T, int EPV, int LOG_TWC, int MODE>
__global__ void __launch_bounds__(kThreads)
    coset_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, CVec<T, EPV> const *__restrict__ in,
                 CVec<T, EPV> *__restrict__ out, int beta, Cx<T> *__restrict__ partials, uint32_t Bpad,
                 T const *__restrict__ Wre, T const *__restrict__ Wim, uint64_t B)
{
    using Cfg = CosetCfg<LOG_TWC>;
    using Vec = CVec<T, EPV>;
    constexpr int TWC = Cfg::TWC, RPT = Cfg::RPT, R = Cfg::R, PITCH = Cfg::PITCH;
    constexpr int ROWS_PER_STEP = kThreads >> LOG_TWC; // rows covered by one cooperative load/store step

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Vec *tile = reinterpret_cast<Vec *>(smem_raw);
    unsigned char *meta = smem_raw + Cfg::TILE_BYTES;
    Cx<T> *s_c = reinterpret_cast<Cx<T> *>(meta);
    uint32_t *s_zl = reinterpret_cast<uint32_t *>(s_c + kCosetChunkStrings);
    uint32_t *s_aux = s_zl + kCosetChunkStrings; // MODE 2: (sidx << 1) | sigma
    uint32_t *s_gxl = s_aux + kCosetChunkStrings;
    uint32_t *s_gstart = s_gxl + kCosetChunkGroups;
    uint64_t *s_comb_hi = reinterpret_cast<uint64_t *>(s_gstart + kCosetChunkGroups + 2);

    uint32_t const tid = threadIdx.x;
    uint64_t const blk = blockIdx.x;
    uint64_t const coset = blk / nColTiles;
    uint32_t const ct = static_cast<uint32_t>(blk - coset * nColTiles);
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);

    // ---- cooperative load of the coset tile
    if (tid < 16)
        s_comb_hi[tid] = comb_of<R>(pass.basis, tid * ROWS_PER_STEP);
    uint32_t const l_lo = tid >> LOG_TWC;
    uint32_t const jv = tid & (TWC - 1);
    uint64_t const row_lo = base ^ comb_of<R>(pass.basis, l_lo);
    uint64_t const col = static_cast<uint64_t>(ct) * TWC + jv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        uint64_t row = row_lo ^ s_comb_hi[k];
        uint32_t l = l_lo + k * ROWS_PER_STEP;
        cp_async16(&tile[l * PITCH + jv], &in[row * rowvecs + col]);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- accumulate all groups of the pass out of shared memory
    Cx<T> acc[RPT][TWC][EPV];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int j = 0; j < TWC; ++j)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                acc[q][j][e] = Cx<T>{0, 0};
    uint64_t const t0 = static_cast<uint64_t>(ct) * TWC * EPV; // first batch column of the tile (MODE 2)

    for (uint32_t ci = 0; ci < pass.n_chunks; ++ci)
    {
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ns = ch.s_hi - ch.s_lo, ng = ch.g_hi - ch.g_lo;
        for (uint32_t s = tid; s < ns; s += kThreads)
        {
            uint32_t sigma = parity64(base & pass.sz[ch.s_lo + s]);
            s_zl[s] = pass.szl[ch.s_lo + s];
            if (MODE == 2)
                s_aux[s] = (pass.sidx[ch.s_lo + s] << 1) | sigma;
            else
            {
                Cx<T> c = pass.scoef[ch.s_lo + s];
                s_c[s] = Cx<T>{flip_sign(c.re, sigma), flip_sign(c.im, sigma)};
            }
        }
        for (uint32_t gq = tid; gq <= ng; gq += kThreads)
        {
            s_gstart[gq] = pass.gstart[ch.g_lo + gq] - ch.s_lo;
            if (gq < ng)
                s_gxl[gq] = pass.gxl[ch.g_lo + gq];
        }
        __syncthreads();

        for (uint32_t gq = 0; gq < ng; ++gq)
        {
            uint32_t const xl = s_gxl[gq];
            uint32_t const s0 = s_gstart[gq], s1 = s_gstart[gq + 1];
            if (MODE != 2)
            {
                Cx<T> d[RPT];
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                    d[q] = Cx<T>{0, 0};
                for (uint32_t s = s0; s < s1; ++s)
                {
                    uint32_t const zl = s_zl[s];
                    Cx<T> const c = s_c[s];
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                    {
                        uint32_t odd = __popc((tid + q * kThreads) & zl) & 1u;
                        d[q].re += flip_sign(c.re, odd);
                        d[q].im += flip_sign(c.im, odd);
                    }
                }
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Vec const *src = &tile[((tid + q * kThreads) ^ xl) * PITCH];
#pragma unroll
                    for (int j = 0; j < TWC; ++j)
                    {
                        Vec v = src[j];
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            cfma(acc[q][j][e], d[q], v.e[e]);
                    }
                }
            }
            else
            {
                Cx<T> d[RPT][TWC][EPV];
#pragma unroll
                for (int q = 0; q < RPT; ++q)
#pragma unroll
                    for (int j = 0; j < TWC; ++j)
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            d[q][j][e] = Cx<T>{0, 0};
                for (uint32_t s = s0; s < s1; ++s)
                {
                    uint32_t const zl = s_zl[s];
                    uint32_t const aux = s_aux[s];
                    uint64_t const wrow = static_cast<uint64_t>(aux >> 1) * B + t0;
                    T wre[TWC * EPV], wim[TWC * EPV];
#pragma unroll
                    for (int c = 0; c < TWC * EPV; ++c)
                    {
                        wre[c] = __ldg(Wre + wrow + c);
                        wim[c] = __ldg(Wim + wrow + c);
                    }
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                    {
                        uint32_t odd = (__popc((tid + q * kThreads) & zl) + aux) & 1u;
#pragma unroll
                        for (int j = 0; j < TWC; ++j)
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                            {
                                d[q][j][e].re += flip_sign(wre[j * EPV + e], odd);
                                d[q][j][e].im += flip_sign(wim[j * EPV + e], odd);
                            }
                    }
                }
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Vec const *src = &tile[((tid + q * kThreads) ^ xl) * PITCH];
#pragma unroll
                    for (int j = 0; j < TWC; ++j)
                    {
                        Vec v = src[j];
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            cfma(acc[q][j][e], d[q][j][e], v.e[e]);
                    }
                }
            }
        }
        __syncthreads(); // metadata staging area is reused by the next chunk
    }

    if (MODE == 1)
    {
        // e(col) = sum over this CTA's rows of conj(psi) * acc ; warp shuffle, then 8 warps through smem
        Cx<T> e_col[TWC][EPV];
#pragma unroll
        for (int j = 0; j < TWC; ++j)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                Cx<T> sum{0, 0};
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Cx<T> a = tile[(tid + q * kThreads) * PITCH + j].e[e];
                    sum.re = fma(a.re, acc[q][j][e].re, sum.re);
                    sum.re = fma(a.im, acc[q][j][e].im, sum.re);
                    sum.im = fma(a.re, acc[q][j][e].im, sum.im);
                    sum.im = fma(-a.im, acc[q][j][e].re, sum.im);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                {
                    sum.re += __shfl_xor_sync(0xffffffffu, sum.re, off);
                    sum.im += __shfl_xor_sync(0xffffffffu, sum.im, off);
                }
                e_col[j][e] = sum;
            }
        __syncthreads(); // everyone is done reading the tile: reuse its first bytes for the cross-warp reduction
        Cx<T> *red = reinterpret_cast<Cx<T> *>(smem_raw);
        if ((tid & 31) == 0)
        {
#pragma unroll
            for (int j = 0; j < TWC; ++j)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    red[(tid >> 5) * (TWC * EPV) + j * EPV + e] = e_col[j][e];
        }
        __syncthreads();
        if (tid < TWC * EPV)
        {
            Cx<T> sum{0, 0};
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w)
            {
                sum.re += red[w * (TWC * EPV) + tid].re;
                sum.im += red[w * (TWC * EPV) + tid].im;
            }
            partials[coset * Bpad + t0 + tid] = sum;
        }
        return;
    }

    // ---- accumulators -> shared memory -> coalesced global stores
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int j = 0; j < TWC; ++j)
        {
            Vec v;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                v.e[e] = acc[q][j][e];
            tile[(tid + q * kThreads) * PITCH + j] = v;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        uint64_t row = row_lo ^ s_comb_hi[k];
        uint32_t l = l_lo + k * ROWS_PER_STEP;
        Vec v = tile[l * PITCH + jv];
        Vec *dst = &out[row * rowvecs + col];
        if (beta)
        {
            Vec o = *dst;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                v.e[e].re += o.e[e].re;
                v.e[e].im += o.e[e].im;
            }
        }
        *dst = v;
    }
}

} // namespace fpk
