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
    uint32_t n_groups; // (sub)groups of the pass = entries of gxl
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

// Tile shape: NT threads per CTA, every thread owns VPT vectors (RPT rows x TWC vectors), so a tile holds
// NT * VPT vectors (64 KiB at NT = 256, VPT = 16) and spans rank R = log2(NT * VPT / TWC) row bits.  VPT = 8 halves
// the accumulator registers (twice as many resident warps) at the price of one rank bit.
template <int LOG_TWC, int LOG_NT, int VPT = 16> struct CosetCfg
{
    static_assert(VPT == 8 || VPT == 16, "8 or 16 vectors per thread");
    static_assert((VPT >> LOG_TWC) >= 1, "a thread owns at least one whole row segment");
    static constexpr int NT = 1 << LOG_NT;
    static constexpr int TWC = 1 << LOG_TWC;              // vectors per row in the tile
    static constexpr int RPT = VPT >> LOG_TWC;            // rows per thread
    static constexpr int LOG_VPT = VPT == 16 ? 4 : 3;
    static constexpr int R = LOG_VPT + LOG_NT - LOG_TWC;  // tile rank: 2^R rows
    static constexpr int ROWS = 1 << R;
    static constexpr int PITCH = TWC + (TWC > 1 ? 1 : 0); // row pitch in vectors (padded against bank conflicts)
    static constexpr size_t TILE_BYTES = static_cast<size_t>(ROWS) * PITCH * 16;
};

// staged string metadata (per chunk) + small tables + cross-warp reduction scratch
template <typename T> struct CosetSmemLayout
{
    static constexpr size_t off_c = 0;                                                // Cx<T>  [CH_S]
    static constexpr size_t off_zl = off_c + kCosetChunkStrings * sizeof(Cx<T>);     // uint32 [CH_S] local z (low 16 bits) | row-slot sign mask (high 16)
    static constexpr size_t off_aux = off_zl + kCosetChunkStrings * 4;               // (spare)
    static constexpr size_t off_sidx = off_aux + kCosetChunkStrings * 4;             // uint32 [CH_S] W row (MODE 2)
    static constexpr size_t off_gxl = off_sidx + kCosetChunkStrings * 4;             // uint32 [CH_G]
    static constexpr size_t off_gstart = off_gxl + kCosetChunkGroups * 4;            // uint32 [CH_G + 2]
    static constexpr size_t off_comb_hi = off_gstart + (kCosetChunkGroups + 2) * 4;  // uint64 [16] load/store steps
    static constexpr size_t off_red = (off_comb_hi + 16 * 8 + 15) / 16 * 16;         // Cx<T>  [8 warps][32 columns]
    static constexpr size_t bytes = off_red + 8 * 32 * sizeof(Cx<float>); // 8 warps x (16 c128 | 32 c64) columns
};

template <typename T, int LOG_TWC, int LOG_NT, int VPT = 16> constexpr size_t coset_smem_bytes()
{
    return CosetCfg<LOG_TWC, LOG_NT, VPT>::TILE_BYTES + CosetSmemLayout<T>::bytes;
}

__device__ __forceinline__ float sign_mul(float, uint32_t odd)
{
    return __int_as_float(0x3f800000 | static_cast<int>(odd << 31));
}
__device__ __forceinline__ double sign_mul(double, uint32_t odd)
{
    return __hiloint2double(0x3ff00000 | static_cast<int>(odd << 31), 0);
}

// One CTA per (coset, column tile).  Several CTAs are resident per SM (2 at NT = 256, 4 at NT = 128) so that some
// are streaming their tile in or out while the others evaluate groups; measured on B200 this beats a persistent
// single-CTA double-buffered variant (too few warps to cover the LDS -> FMA latency) by 1.3-1.5x.
template <typename T, int EPV, int LOG_TWC, int LOG_NT, int MODE, int VPT = 16>
__global__ void __launch_bounds__(1 << LOG_NT)
    coset_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, CVec<T, EPV> const *__restrict__ in,
                 CVec<T, EPV> *__restrict__ out, int beta, Cx<T> *__restrict__ partials, uint32_t Bpad,
                 T const *__restrict__ Wre, T const *__restrict__ Wim, uint64_t B)
{
    using Cfg = CosetCfg<LOG_TWC, LOG_NT, VPT>;
    using L = CosetSmemLayout<T>;
    using Vec = CVec<T, EPV>;
    constexpr int NT = Cfg::NT, TWC = Cfg::TWC, RPT = Cfg::RPT, R = Cfg::R, PITCH = Cfg::PITCH;
    constexpr int ROWS_PER_STEP = NT >> LOG_TWC; // rows covered by one cooperative load/store step (VPT steps)
    constexpr int NCOL = TWC * EPV;              // batch columns per tile

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Vec *tile = reinterpret_cast<Vec *>(smem_raw);
    unsigned char *meta = smem_raw + Cfg::TILE_BYTES;
    Cx<T> *s_c = reinterpret_cast<Cx<T> *>(meta + L::off_c);
    uint32_t *s_zl = reinterpret_cast<uint32_t *>(meta + L::off_zl);
    uint32_t *s_sidx = reinterpret_cast<uint32_t *>(meta + L::off_sidx);
    uint32_t *s_gxl = reinterpret_cast<uint32_t *>(meta + L::off_gxl);
    uint32_t *s_gstart = reinterpret_cast<uint32_t *>(meta + L::off_gstart);
    uint64_t *s_comb_hi = reinterpret_cast<uint64_t *>(meta + L::off_comb_hi);
    Cx<T> *s_red = reinterpret_cast<Cx<T> *>(meta + L::off_red);

    uint32_t const tid = threadIdx.x;
    uint64_t const blk = blockIdx.x;
    uint64_t const coset = blk / nColTiles;
    uint32_t const ct = static_cast<uint32_t>(blk - coset * nColTiles);
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
    uint64_t const t0 = static_cast<uint64_t>(ct) * NCOL; // first batch column of the tile
    uint64_t const vcol0 = static_cast<uint64_t>(ct) * TWC;

    // ---- cooperative load of the coset tile: vector jv of rows l_lo + k * ROWS_PER_STEP
    if (tid < VPT)
        s_comb_hi[tid] = comb_of<R>(pass.basis, tid * ROWS_PER_STEP);
    uint32_t const l_lo = tid >> LOG_TWC;
    uint32_t const jv = tid & (TWC - 1);
    uint64_t const row_lo = base ^ comb_of<R>(pass.basis, l_lo);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPT; ++k)
    {
        uint64_t row = row_lo ^ s_comb_hi[k];
        uint32_t l = l_lo + k * ROWS_PER_STEP;
        cp_async16(&tile[l * PITCH + jv], &in[row * rowvecs + vcol0 + jv]);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    Cx<T> acc[RPT][TWC][EPV];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int j = 0; j < TWC; ++j)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                acc[q][j][e] = Cx<T>{0, 0};

    for (uint32_t ci = 0; ci < pass.n_chunks; ++ci)
    {
        // stage this chunk's metadata; the coset-base sign and the sign of the row bits above the thread index
        // (rows tid + NT*q, q < 16) are folded into one 16-bit mask per string
        CosetChunk const ch = pass.chunks[ci];
        uint32_t const ns = ch.s_hi - ch.s_lo, ng = ch.g_hi - ch.g_lo;
        for (uint32_t s = tid; s < ns; s += NT)
        {
            uint32_t const zl = pass.szl[ch.s_lo + s];
            uint32_t hp = parity64(base & pass.sz[ch.s_lo + s]) ? 0xffffu : 0u;
#pragma unroll
            for (uint32_t q = 0; q < RPT; ++q)
                hp ^= (__popc((q << LOG_NT) & zl) & 1u) << q;
            s_zl[s] = (zl & (NT - 1)) | (hp << 16); // one broadcast LDS per string: thread-index z bits | row-slot signs
            if (MODE == 2)
                s_sidx[s] = pass.sidx[ch.s_lo + s];
            else
                s_c[s] = pass.scoef[ch.s_lo + s];
        }
        for (uint32_t gq = tid; gq <= ng; gq += NT)
        {
            s_gstart[gq] = pass.gstart[ch.g_lo + gq] - ch.s_lo;
            if (gq < ng)
                s_gxl[gq] = pass.gxl[ch.g_lo + gq];
        }
        if (ci == 0)
            asm volatile("cp.async.wait_group 0;\n" ::: "memory"); // the tile landed while the metadata was staged
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
                    Cx<T> const c = s_c[s];
                    uint32_t const zm = s_zl[s];
                    uint32_t const par = (zm >> 16) ^ ((__popc(tid & zm & 0xffffu) & 1u) ? 0xffffu : 0u);
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                    {
                        uint32_t odd = (par >> q) & 1u;
                        d[q].re += flip_sign(c.re, odd);
                        d[q].im += flip_sign(c.im, odd);
                    }
                }
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Vec const *src = &tile[((tid + q * NT) ^ xl) * PITCH];
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
                    uint32_t const zm = s_zl[s];
                    uint32_t const par = (zm >> 16) ^ ((__popc(tid & zm & 0xffffu) & 1u) ? 0xffffu : 0u);
                    uint64_t const wrow = static_cast<uint64_t>(s_sidx[s]) * B + t0;
                    T wre[NCOL], wim[NCOL];
#pragma unroll
                    for (int c = 0; c < NCOL; ++c)
                    {
                        wre[c] = __ldg(Wre + wrow + c);
                        wim[c] = __ldg(Wim + wrow + c);
                    }
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                    {
                        T const sf = sign_mul(T(0), (par >> q) & 1u); // +-1: one FMA per component, no sign flips
#pragma unroll
                        for (int j = 0; j < TWC; ++j)
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                            {
                                d[q][j][e].re = fma(sf, wre[j * EPV + e], d[q][j][e].re);
                                d[q][j][e].im = fma(sf, wim[j * EPV + e], d[q][j][e].im);
                            }
                    }
                }
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Vec const *src = &tile[((tid + q * NT) ^ xl) * PITCH];
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
        // e(col) = sum over this tile's rows of conj(psi) * acc: warp shuffle, then the warps through smem
#pragma unroll
        for (int j = 0; j < TWC; ++j)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                Cx<T> sum{0, 0};
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                {
                    Cx<T> a = tile[(tid + q * NT) * PITCH + j].e[e];
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
                if ((tid & 31) == 0)
                    s_red[(tid >> 5) * NCOL + j * EPV + e] = sum;
            }
        __syncthreads();
        if (tid < NCOL)
        {
            Cx<T> sum{0, 0};
#pragma unroll
            for (int w = 0; w < NT / 32; ++w)
            {
                sum.re += s_red[w * NCOL + tid].re;
                sum.im += s_red[w * NCOL + tid].im;
            }
            partials[coset * Bpad + t0 + tid] = sum;
        }
    }
    else
    {
    // ---- accumulators -> the (now dead) tile buffer -> coalesced 16-byte stores: a warp writes whole TWC*16-byte
    // row segments (direct register stores scatter 32 sixteen-byte pieces per instruction and saturate the LSU
    // queue: measured 2.7x slower)
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int j = 0; j < TWC; ++j)
        {
            Vec v;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                v.e[e] = acc[q][j][e];
            tile[(tid + q * NT) * PITCH + j] = v;
        }
    __syncthreads();
    if (beta)
    {
        // accumulating pass: all VPT read-modify-write loads are issued before the first add (one round trip to
        // memory per thread instead of VPT dependent ones)
        Vec o[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k)
            o[k] = out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol0 + jv];
#pragma unroll
        for (int k = 0; k < VPT; ++k)
        {
            uint32_t l = l_lo + k * ROWS_PER_STEP;
            Vec v = tile[l * PITCH + jv];
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                v.e[e].re += o[k].e[e].re;
                v.e[e].im += o[k].e[e].im;
            }
            out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol0 + jv] = v;
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < VPT; ++k)
        {
            uint64_t row = row_lo ^ s_comb_hi[k];
            uint32_t l = l_lo + k * ROWS_PER_STEP;
            out[row * rowvecs + vcol0 + jv] = tile[l * PITCH + jv];
        }
    }
    }
}

// ---------------------------------------------------------------- K4b: per-string expectation values, state in smem
// SummedPauliOp::expectation_value first stage (SPO:573-577) for small registers (n <= 12 qubits): one CTA stages a
// whole state column tile (2^n rows x one 16-byte vector = 1 complex128 / 2 complex64 columns, <= 64 KiB) in shared
// memory ONCE and then evaluates every string against it: warps take the x-mask chunks (<= kPairMS strings sharing a
// gather) round-robin, lanes stride over the unordered row pairs {i, i^x} of the chunk (same pairing identity as
// expval_pairs_kernel: one real accumulator per string and column), a warp shuffle finishes each string and lane 0
// writes E(s, t) directly -- no partial sums, no second stage, and the batch is read from HBM once instead of once
// per chunk.  blockIdx.y splits the chunk list when there are too few column tiles to fill the chip.
// (Guarding the dead slots of short chunks with a warp-uniform branch was measured slower: 47 vs 42.6 ms.)
template <typename T, int EPV, int MS>
__global__ void __launch_bounds__(kThreads)
    sop_expval_tile_kernel(PairChunk const *__restrict__ chunks, uint32_t n_chunks, uint64_t const *__restrict__ sz,
                           uint8_t const *__restrict__ sodd, uint32_t n_qubits, uint64_t rowvecs,
                           CVec<T, EPV> const *__restrict__ in, T *__restrict__ E /* [S][B] */, uint64_t B)
{
    using Vec = CVec<T, EPV>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Vec *tile = reinterpret_cast<Vec *>(smem_raw);
    uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t const rows = 1u << n_qubits;
    uint64_t const v = blockIdx.x; // column vector of this CTA
    for (uint32_t r = tid; r < rows; r += kThreads)
        cp_async16(&tile[r], &in[static_cast<uint64_t>(r) * rowvecs + v]);
    cp_async_wait_all();
    __syncthreads();

    uint32_t const n_warps_total = (kThreads / 32) * gridDim.y;
    for (uint32_t c = blockIdx.y * (kThreads / 32) + warp; c < n_chunks; c += n_warps_total)
    {
        PairChunk const ch = chunks[c];
        uint64_t zs[MS];
        uint32_t odd_ny[MS];
#pragma unroll
        for (int m = 0; m < MS; ++m)
        {
            bool live = static_cast<uint32_t>(m) < ch.count;
            zs[m] = live ? sz[ch.s0 + m] : 0;
            odd_ny[m] = live ? sodd[ch.s0 + m] : 0;
        }
        uint32_t const x = static_cast<uint32_t>(ch.x);
        uint32_t const npairs = ch.diag ? rows : (rows >> 1);
        uint32_t const low_mask = ch.diag ? 0xffffffffu : ((1u << ch.hbit) - 1u);
        T r[MS][EPV];
#pragma unroll
        for (int m = 0; m < MS; ++m)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                r[m][e] = 0;
        for (uint32_t p = lane; p < npairs; p += 32)
        {
            uint32_t const i = ch.diag ? p : (((p & ~low_mask) << 1) | (p & low_mask));
            Vec const a = tile[i];
            T qre[EPV], qim[EPV];
            if (ch.diag)
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    qre[e] = fma(a.e[e].re, a.e[e].re, a.e[e].im * a.e[e].im);
                    qim[e] = 0;
                }
            }
            else
            {
                Vec const b = tile[i ^ x];
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    qre[e] = fma(a.e[e].re, b.e[e].re, a.e[e].im * b.e[e].im);
                    qim[e] = fma(a.e[e].re, b.e[e].im, -a.e[e].im * b.e[e].re);
                }
            }
#pragma unroll
            for (int m = 0; m < MS; ++m)
            {
                uint32_t const sgn = __popc(i & static_cast<uint32_t>(zs[m])) & 1u;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    r[m][e] += flip_sign(odd_ny[m] ? qim[e] : qre[e], sgn);
            }
        }
#pragma unroll
        for (int m = 0; m < MS; ++m)
        {
            if (static_cast<uint32_t>(m) >= ch.count)
                break;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                T val = r[m][e];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                    val += __shfl_xor_sync(0xffffffffu, val, off);
                if (lane == 0)
                    E[static_cast<uint64_t>(ch.s0 + m) * B + v * EPV + e] = val;
            }
        }
    }
}

} // namespace fpk
