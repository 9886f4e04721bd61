// K3e: coset-blocked apply for passes with FEW x-masks (<= GMAX groups): the row factors live in registers.
//
// Same tiling as coset_kernel (coset.cuh): a CTA owns the 256 rows of one coset of the pass' x-mask subspace and
// stages them in shared memory one column tile (TWC 16-byte vectors per row) at a time.  What is different:
//
//   * the row factors D_g(l) = sum_{s in g} c_s (-1)^par(row & z_s) of the thread's row are formed ONCE per coset and
//     kept in registers while the CTA walks all its column tiles -- coset_kernel re-reads every string's metadata
//     from shared memory for every column tile (17 % of its shared-memory wavefronts at 64 strings / 8 masks);
//   * the tile is dense (no padded pitch) and bank conflicts are avoided by a column rotation keyed on the READER's
//     own row: accumulator j of the thread with local row l holds column j ^ (l & (TWC-1)), so the 8 lanes of a
//     quarter warp always read 8 different 16-byte bank groups whatever the gathered row (l ^ xl) is, and the XOR
//     key is a per-thread constant (one LOP3 per LDS.128, no per-group address table).
//
// Per complex FMA the kernel still needs one LDS.128 (16 B of shared-memory bandwidth per 4 DFMA); everything else
// the load/store unit did per tile -- metadata broadcasts, bank-conflict replays -- is gone.
// Reference semantics: PauliOp::apply (PO:399-468): out(i,t) (+)= sum_s h_s m_s(i) psi(i ^ x_s, t).
#pragma once
#include <cstdint>
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint on the host)
#include <cuda_runtime.h>

#include "coset.cuh"

namespace fpk
{

#ifdef FP_FEW_PROFILE
__device__ unsigned long long g_few_prof[8];
#define FEW_T(i)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        if (threadIdx.x == 0)                                                                                          \
        {                                                                                                              \
            long long now_ = clock64();                                                                                \
            atomicAdd(&g_few_prof[i], static_cast<unsigned long long>(now_ - t_prev_));                               \
            t_prev_ = now_;                                                                                            \
        }                                                                                                              \
    } while (0)
#else
#define FEW_T(i)
#endif

constexpr uint32_t kFewMaxStrings = 256; // strings staged in shared memory per round of the row-factor phase
constexpr uint32_t kFewParamStrings = 128; // strings that fit the kernel parameter block (constant bank)

// Strings of one pass as kernel parameters: coefficient (already times (-i)^nY) and full z-mask per string, group
// boundaries.  ~3 KiB for complex128: inside the 4 KiB parameter space every driver supports.
template <typename T> struct FewStrings
{
    Cx<T> c[kFewParamStrings];
    uint64_t z[kFewParamStrings];
    uint32_t gs[9]; // first string of each group (GMAX = 8), gs[n_groups] = number of strings
    uint32_t gxl[8];
};


// Row factors D_g = sum_{s in g} c_s (-1)^par(row & z_s) of one row from the strings in the constant bank.
// Registers of <= 32 qubits use 32-bit masks (one POPC per string instead of two).
template <typename T, int GMAX>
__device__ __forceinline__ void few_row_factors(FewStrings<T> const &strs, uint32_t ng, uint64_t my_row, bool narrow, Cx<T> (&D)[GMAX])
{
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
    {
        D[g] = Cx<T>{0, 0};
        if (static_cast<uint32_t>(g) < ng)
        {
            uint32_t const s1 = strs.gs[g + 1];
            if (narrow)
            {
                uint32_t const r32 = static_cast<uint32_t>(my_row);
#pragma unroll 4
                for (uint32_t s = strs.gs[g]; s < s1; ++s)
                {
                    Cx<T> const c = strs.c[s];
                    uint32_t const odd = __popc(r32 & static_cast<uint32_t>(strs.z[s])) & 1u;
                    D[g].re += flip_sign(c.re, odd);
                    D[g].im += flip_sign(c.im, odd);
                }
            }
            else
            {
#pragma unroll 4
                for (uint32_t s = strs.gs[g]; s < s1; ++s)
                {
                    Cx<T> const c = strs.c[s];
                    uint32_t const odd = parity64(my_row & strs.z[s]);
                    D[g].re += flip_sign(c.re, odd);
                    D[g].im += flip_sign(c.im, odd);
                }
            }
        }
    }
}

template <int LOG_TWC> struct FewCfg
{
    static constexpr int NT = 256;                  // threads = rows of the tile
    static constexpr int R = 8;                     // tile rank
    static constexpr int TWC = 1 << LOG_TWC;        // vectors per row segment
    static constexpr int RPS = NT >> LOG_TWC;       // rows covered by one cooperative load/store step
    static constexpr int STEPS = NT / RPS;          // = TWC
    static constexpr size_t TILE_BYTES = static_cast<size_t>(NT) * TWC * 16; // one buffer
    static_assert(LOG_TWC == 3 || LOG_TWC == 4, "row segments of 128 or 256 bytes");
};

// NBUF = 2: the next column tile is requested (cp.async) before the current one is evaluated, so the CTA never waits
// for HBM in steady state; MINB = resident CTAs per SM the register budget is compiled for.
// PSTR: the pass' strings (<= kFewParamStrings) travel in the kernel parameter block, i.e. the constant bank: the
// row-factor phase then needs no shared-memory staging, no barrier and -- the point -- no LDS at all, so it no longer
// queues behind the other resident CTA's gathers in the load/store unit (measured: 13 k -> 4 k cycles per CTA).
// RMWPF: accumulating passes (beta != 0) prefetch the OLD output rows of a tile with cp.async into one more
// shared-memory buffer while the tile's gathers run, instead of loading them (LDG, 8-16 vectors live in registers, a
// full memory round trip exposed) in the store phase.
// MODE 1: PauliOp::expectation_value partials (PO:482-549) instead of the store: e(t) = sum over the tile's rows of
// conj(psi(l,t)) (A psi)(l,t), written to partials[coset][t] (fixed summation order; finalize_complex_kernel sums the
// cosets).  The products are staged in the (dead) tile buffer and column-summed by the cooperative row mapping.
template <typename T, int EPV, int LOG_TWC, int GMAX, int NBUF = 1, int MINB = 2, bool PSTR = false, bool RMWPF = false,
          int MODE = 0>
__global__ void __launch_bounds__(256, MINB)
    coset_few_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, uint32_t ctPerCta, uint32_t nCtGroups,
                     CVec<T, EPV> const *__restrict__ in, CVec<T, EPV> *__restrict__ out, int beta,
                     const __grid_constant__ FewStrings<T> strs, Cx<T> *__restrict__ partials = nullptr,
                     uint32_t Bpad = 0)
{
    using Cfg = FewCfg<LOG_TWC>;
    using Vec = CVec<T, EPV>;
    constexpr int TWC = Cfg::TWC, RPS = Cfg::RPS, STEPS = Cfg::STEPS;
    constexpr uint32_t ROW_SHIFT = LOG_TWC + 4; // bytes per tile row = 1 << ROW_SHIFT

    extern __shared__ __align__(1024) unsigned char smem_few[];
    __shared__ uint64_t s_comb_hi[STEPS];
    __shared__ uint32_t s_gxl[GMAX];
    __shared__ uint32_t s_gs[GMAX + 1];
    __shared__ Cx<T> s_c[kFewMaxStrings];
    __shared__ uint32_t s_zl[kFewMaxStrings];
    __shared__ CVec<T, EPV> s_red[MODE == 1 ? 8 * (1 << LOG_TWC) : 1];

    uint32_t const tid = threadIdx.x;
#ifdef FP_FEW_PROFILE
    long long t_prev_ = clock64();
#endif
    uint32_t const ng = pass.n_groups;
    uint64_t const coset = blockIdx.x / nCtGroups;
    uint32_t const ctg = static_cast<uint32_t>(blockIdx.x - coset * nCtGroups);
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
    uint32_t ct = ctg * ctPerCta;
    uint32_t const ct_end = min(nColTiles, ct + ctPerCta);

    if (tid < STEPS)
        s_comb_hi[tid] = comb_of<Cfg::R>(pass.basis, tid * RPS);
    if (tid < ng)
        s_gxl[tid] = PSTR ? strs.gxl[tid] : pass.gxl[tid];
    uint32_t const l_lo = tid >> LOG_TWC;
    uint32_t const jv = tid & (TWC - 1);
    uint64_t const row_lo = base ^ comb_of<Cfg::R>(pass.basis, l_lo);
    __syncthreads();

    auto fill = [&](uint32_t c, uint32_t buf, bool commit) {
        uint64_t const vcol = static_cast<uint64_t>(c) * TWC + jv;
        Vec *tile = reinterpret_cast<Vec *>(smem_few + buf * Cfg::TILE_BYTES);
#pragma unroll
        for (int k = 0; k < STEPS; ++k)
        {
            uint64_t const row = row_lo ^ s_comb_hi[k];
            cp_async16(&tile[(l_lo + k * RPS) * TWC + jv], &in[row * rowvecs + vcol]);
        }
        if (commit)
            asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    bool const pf = RMWPF && beta != 0;
    unsigned char *const ob = smem_few + NBUF * Cfg::TILE_BYTES; // old output rows of the current tile (RMWPF)
    auto fill_old = [&](uint32_t c) {
        uint64_t const vcol = static_cast<uint64_t>(c) * TWC + jv;
        Vec *o = reinterpret_cast<Vec *>(ob);
#pragma unroll
        for (int k = 0; k < STEPS; ++k)
            cp_async16(&o[(l_lo + k * RPS) * TWC + jv], &out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol]);
    };
    fill(ct, 0, true);
    if (NBUF == 2 && ct + 1 < ct_end)
        fill(ct + 1, 1, false);
    if (pf)
        fill_old(ct);
    if (pf || (NBUF == 2 && ct + 1 < ct_end))
        asm volatile("cp.async.commit_group;\n" ::: "memory");

    // ---- row factors of this thread's row, once per coset (the first tile is in flight meanwhile).  The strings are
    // staged in shared memory in rounds of kFewMaxStrings with the coset-base sign par(base & z_s) folded into the
    // coefficient, so the per-thread loop is broadcast LDS + the row-local sign par(l & zl_s) only.
    Cx<T> D[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
        D[g] = Cx<T>{0, 0};
    if (PSTR)
    {
        // strings in the constant bank: sign from the thread's global row, no shared memory, no barrier
        few_row_factors<T, GMAX>(strs, ng, base ^ comb_of<Cfg::R>(pass.basis, tid), (pass.nonpivot_mask >> 32) == 0, D);
    }
    uint32_t const n_str = PSTR ? 0u : pass.gstart[ng];
    for (uint32_t r0 = 0; r0 < n_str; r0 += kFewMaxStrings)
    {
        uint32_t const cnt = min(kFewMaxStrings, n_str - r0);
        if (r0)
            __syncthreads();
        if (tid < cnt)
        {
            Cx<T> c = pass.scoef[r0 + tid];
            uint32_t const odd = parity64(base & pass.sz[r0 + tid]);
            c.re = flip_sign(c.re, odd);
            c.im = flip_sign(c.im, odd);
            s_c[tid] = c;
            s_zl[tid] = pass.szl[r0 + tid];
        }
        if (tid <= ng)
        {
            uint32_t const gs = pass.gstart[tid];
            s_gs[tid] = gs < r0 ? 0u : min(gs - r0, cnt);
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
        {
            if (static_cast<uint32_t>(g) < ng)
            {
                uint32_t const s1 = s_gs[g + 1];
#pragma unroll 4
                for (uint32_t s = s_gs[g]; s < s1; ++s)
                {
                    Cx<T> const c = s_c[s];
                    uint32_t const odd = __popc(tid & s_zl[s]) & 1u;
                    D[g].re += flip_sign(c.re, odd);
                    D[g].im += flip_sign(c.im, odd);
                }
            }
        }
    }

    FEW_T(0); // setup + row factors
    uint32_t const key_off = (tid & (TWC - 1)) << 4;             // column rotation of this thread, in bytes
    uint32_t const own_off = (tid << ROW_SHIFT) | key_off;       // own row, rotated column 0

    for (uint32_t it = 0; ct < ct_end; ++ct, ++it)
    {
        uint32_t const buf = NBUF == 2 ? (it & 1u) : 0u;
        unsigned char *const tb = smem_few + buf * Cfg::TILE_BYTES;
        Vec *const tile = reinterpret_cast<Vec *>(tb);
        if (pf || (NBUF == 2 && ct + 1 < ct_end))
            asm volatile("cp.async.wait_group 1;\n" ::: "memory"); // a newer group (next tile / old rows) may be in flight
        else
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        FEW_T(1); // waiting for the tile

        Cx<T> acc[TWC][EPV];
#pragma unroll
        for (int j = 0; j < TWC; ++j)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                acc[j][e] = Cx<T>{0, 0};
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
        {
            if (static_cast<uint32_t>(g) < ng)
            {
                uint32_t const src = own_off ^ (s_gxl[g] << ROW_SHIFT);
                Cx<T> const d = D[g];
#pragma unroll
                for (int j = 0; j < TWC; ++j)
                {
                    Vec const v = *reinterpret_cast<Vec const *>(tb + (src ^ (j << 4)));
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[j][e], d, v.e[e]);
                }
            }
        }
        __syncthreads(); // every gather of this tile is done: the buffer becomes the store staging area
        FEW_T(2); // gathers + FMAs

#pragma unroll
        for (int j = 0; j < TWC; ++j)
        {
            Vec v;
            if (MODE == 1)
            {
                // conj(psi) * acc for the thread's own row: read and rewritten in place (no other thread touches
                // these slots after the barrier above)
                Vec const a = *reinterpret_cast<Vec const *>(tb + (own_off ^ (j << 4)));
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    v.e[e].re = fma(a.e[e].re, acc[j][e].re, a.e[e].im * acc[j][e].im);
                    v.e[e].im = fma(a.e[e].re, acc[j][e].im, -a.e[e].im * acc[j][e].re);
                }
            }
            else
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    v.e[e] = acc[j][e];
            }
            *reinterpret_cast<Vec *>(tb + (own_off ^ (j << 4))) = v;
        }
        if (pf)
            asm volatile("cp.async.wait_group 0;\n" ::: "memory"); // the old output rows of this tile have landed
        __syncthreads();
        FEW_T(3); // accumulators -> staging

        if (MODE == 1)
        {
            // column sums: this thread adds rows l_lo + k * RPS of vector column jv, the warp's 32 / TWC row groups are
            // folded by shuffles, the 8 warps through shared memory (fixed order)
            Vec sum;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                sum.e[e] = Cx<T>{0, 0};
#pragma unroll
            for (int k = 0; k < STEPS; ++k)
            {
                Vec const v = tile[(l_lo + k * RPS) * TWC + jv];
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    sum.e[e].re += v.e[e].re;
                    sum.e[e].im += v.e[e].im;
                }
            }
#pragma unroll
            for (int off = TWC; off < 32; off <<= 1)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    sum.e[e].re += __shfl_xor_sync(0xffffffffu, sum.e[e].re, off);
                    sum.e[e].im += __shfl_xor_sync(0xffffffffu, sum.e[e].im, off);
                }
            __syncthreads(); // every staged product has been read: s_red may alias nothing, but the tile is reused below
            if ((tid & 31u) < static_cast<uint32_t>(TWC))
                s_red[(tid >> 5) * TWC + jv] = sum;
            __syncthreads();
            if (tid < static_cast<uint32_t>(TWC))
            {
                Vec tot = s_red[tid];
#pragma unroll
                for (int w = 1; w < 8; ++w)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        tot.e[e].re += s_red[w * TWC + tid].e[e].re;
                        tot.e[e].im += s_red[w * TWC + tid].e[e].im;
                    }
                uint64_t const col0 = (static_cast<uint64_t>(ct) * TWC + tid) * EPV;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    partials[coset * Bpad + col0 + e] = tot.e[e];
            }
        }
        else
        {

        uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jv;
        if (pf)
        {
            Vec const *o = reinterpret_cast<Vec const *>(ob);
#pragma unroll
            for (int k = 0; k < STEPS; ++k)
            {
                Vec v = tile[(l_lo + k * RPS) * TWC + jv];
                Vec const w = o[(l_lo + k * RPS) * TWC + jv];
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    v.e[e].re += w.e[e].re;
                    v.e[e].im += w.e[e].im;
                }
                out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol] = v;
            }
        }
        else if (beta)
        {
            Vec o[STEPS];
#pragma unroll
            for (int k = 0; k < STEPS; ++k)
                o[k] = out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol];
#pragma unroll
            for (int k = 0; k < STEPS; ++k)
            {
                Vec v = tile[(l_lo + k * RPS) * TWC + jv];
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    v.e[e].re += o[k].e[e].re;
                    v.e[e].im += o[k].e[e].im;
                }
                out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol] = v;
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < STEPS; ++k)
                out[(row_lo ^ s_comb_hi[k]) * rowvecs + vcol] = tile[(l_lo + k * RPS) * TWC + jv];
        }
        }
        FEW_T(4); // issuing the stores
        if (ct + NBUF < ct_end || (pf && ct + 1 < ct_end))
        {
            __syncthreads(); // staged (and old) rows have been read: refill this buffer / the old-row buffer
            if (ct + NBUF < ct_end)
                fill(ct + NBUF, buf, false);
            if (pf && ct + 1 < ct_end)
                fill_old(ct + 1);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            FEW_T(5);
        }
    }
}


// ================================================================ K3f: TMA-fed, warp-specialised variant of K3e
// One persistent CTA per SM: a producer warp streams coset tiles into a ring of three 64 KiB shared-memory buffers
// with TMA tile::gather4 (four scattered 256-byte row segments per operation, completion on an mbarrier), and two
// consumer groups of 256 threads evaluate alternate tiles exactly like coset_few_kernel.  Why: cp.async costs the
// load/store unit ~15 cycles per 512-byte warp operation (measured: the gather phase of a resident CTA shortens from
// 7.6 k to 5.8 k cycles when the OTHER CTA's fill moves to the TMA), and with the fill off the LSU and off the
// consumers' critical path the kernel is bound by its gathers and HBM only.
//
// Tile order of a CTA (deterministic, so buffer = tile index mod 3): coset pairs p = blockIdx.x + k * gridDim.x;
// consumer group g owns coset 2p + g; tiles alternate between the groups: i -> group i & 1, column tile (i / 2) % nct.
__device__ __forceinline__ uint32_t few_smem_u32(void const *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void few_mbar_init(uint64_t *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(few_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void few_mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(few_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void few_mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(few_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void few_mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\n"
                 "DONE_%=:\n}" ::"r"(few_smem_u32(b)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void few_tma_gather4(void *dst, CUtensorMap const *tm, int c0, int r0, int r1, int r2, int r3,
                                                uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
                 "%5, %6}], [%7];" ::"r"(few_smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(few_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void few_group_sync(uint32_t group)
{
    asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
}

constexpr int kFewTmaThreads = 2 * 256 + 128; // two consumer groups + the producer warpgroup (one live warp)
constexpr int kFewTmaBufs = 3;
constexpr size_t kFewTmaTile = 256 * 16 * 16; // 256 rows x 16 vectors x 16 bytes

template <typename T, int EPV, int GMAX>
__global__ void __launch_bounds__(kFewTmaThreads, 1)
    coset_few_tma_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, uint64_t nPairs,
                         CVec<T, EPV> *__restrict__ out, int beta, const __grid_constant__ FewStrings<T> strs,
                         const __grid_constant__ CUtensorMap tm_in)
{
    using Vec = CVec<T, EPV>;
    constexpr int TWC = 16, RPS = 16, STEPS = 16, R = 8;
    constexpr uint32_t ROW_SHIFT = 8;

    extern __shared__ __align__(1024) unsigned char smem_ft[];
    __shared__ uint64_t s_full[kFewTmaBufs], s_empty[kFewTmaBufs];
    __shared__ uint32_t s_comb[256];    // XOR offsets of the 256 local rows
    __shared__ uint64_t s_comb_hi[STEPS];

    uint32_t const tid = threadIdx.x;
    uint32_t const ng = pass.n_groups;
    if (tid < 256)
        s_comb[tid] = static_cast<uint32_t>(comb_of<R>(pass.basis, tid));
    if (tid < STEPS)
        s_comb_hi[tid] = comb_of<R>(pass.basis, tid * RPS);
    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < kFewTmaBufs; ++b)
        {
            few_mbar_init(&s_full[b], 1);
            few_mbar_init(&s_empty[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint64_t const my_pairs = blockIdx.x < nPairs ? (nPairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint64_t const n_tiles = my_pairs * 2 * nColTiles;

    if (tid >= 512)
    {
        // ------------------------------------------------ producer warpgroup: hand its registers to the consumers
        // (640 threads are launched with 96 registers each; 24 are enough here, the consumers grow to 112: a CTA can only re-use what its own warps released)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid >= 544)
            return;
        uint32_t const lane = tid & 31u;
        for (uint64_t i = 0; i < n_tiles; ++i)
        {
            uint32_t const buf = static_cast<uint32_t>(i % kFewTmaBufs);
            if (i >= kFewTmaBufs)
                few_mbar_wait(&s_empty[buf], static_cast<uint32_t>((i / kFewTmaBufs) - 1) & 1u);
            uint64_t const v = i >> 1;
            uint64_t const pair = blockIdx.x + (v / nColTiles) * gridDim.x;
            uint64_t const coset = 2 * pair + (i & 1u);
            uint32_t const ct = static_cast<uint32_t>(v % nColTiles);
            uint32_t const base = static_cast<uint32_t>(deposit_bits(coset, pass.nonpivot_mask));
            if (lane == 0)
                few_mbar_expect_tx(&s_full[buf], static_cast<uint32_t>(kFewTmaTile));
            __syncwarp();
            int const c0 = static_cast<int>(ct) * TWC * static_cast<int>(16 / sizeof(T));
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                uint32_t const op = lane + 32 * h; // rows 4*op .. 4*op+3
                few_tma_gather4(smem_ft + buf * kFewTmaTile + (static_cast<size_t>(op) << (ROW_SHIFT + 2)), &tm_in, c0,
                                base ^ s_comb[4 * op], base ^ s_comb[4 * op + 1], base ^ s_comb[4 * op + 2],
                                base ^ s_comb[4 * op + 3], &s_full[buf]);
            }
        }
        return;
    }

    // ---------------------------------------------------- consumer groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    uint32_t const grp = tid >> 8;
    uint32_t const l = tid & 255u; // local row of this thread
    uint32_t const l_lo = l >> 4, jv = l & 15u;
    uint32_t const key_off = (l & 15u) << 4;
    uint32_t const own_off = (l << ROW_SHIFT) | key_off;

    uint64_t const n_q = my_pairs * nColTiles; // tiles of this group
    uint64_t const shift = 0;
    uint64_t cur_k = ~0ull, row_lo = 0;
    Cx<T> D[GMAX];
    for (uint64_t q = 0; q < n_q; ++q)
    {
        uint64_t const v = q + shift;
        uint64_t const k = (v / nColTiles) % my_pairs;
        uint32_t const ct = static_cast<uint32_t>(v % nColTiles);
        if (k != cur_k)
        {
            cur_k = k;
            uint64_t const coset = 2 * (blockIdx.x + k * gridDim.x) + grp;
            uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
            row_lo = base ^ s_comb[l_lo];
            few_row_factors<T, GMAX>(strs, ng, base ^ s_comb[l], true, D); // launched for <= 30 qubits only
        }
        {
            uint64_t const i = q * 2 + grp; // position in the CTA's tile order
            uint32_t const buf = static_cast<uint32_t>(i % kFewTmaBufs);
            unsigned char *const tb = smem_ft + buf * kFewTmaTile;
            Vec *const tile = reinterpret_cast<Vec *>(tb);
            few_mbar_wait(&s_full[buf], static_cast<uint32_t>(i / kFewTmaBufs) & 1u);

            Cx<T> acc[TWC][EPV];
#pragma unroll
            for (int j = 0; j < TWC; ++j)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[j][e] = Cx<T>{0, 0};
#pragma unroll
            for (int g = 0; g < GMAX; ++g)
            {
                if (static_cast<uint32_t>(g) < ng)
                {
                    uint32_t const src = own_off ^ (strs.gxl[g] << ROW_SHIFT);
#pragma unroll
                    for (int j = 0; j < TWC; ++j)
                    {
                        Vec const v = *reinterpret_cast<Vec const *>(tb + (src ^ (j << 4)));
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            cfma(acc[j][e], D[g], v.e[e]);
                    }
                }
            }
            few_group_sync(grp); // every gather of this tile is done: the buffer becomes the store staging area
#pragma unroll
            for (int j = 0; j < TWC; ++j)
            {
                Vec v;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    v.e[e] = acc[j][e];
                *reinterpret_cast<Vec *>(tb + (own_off ^ (j << 4))) = v;
            }
            few_group_sync(grp);
            uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jv;
            if (beta)
            {
                Vec o[STEPS];
#pragma unroll
                for (int q = 0; q < STEPS; ++q)
                    o[q] = out[(row_lo ^ s_comb_hi[q]) * rowvecs + vcol];
#pragma unroll
                for (int q = 0; q < STEPS; ++q)
                {
                    Vec v = tile[(l_lo + q * RPS) * TWC + jv];
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        v.e[e].re += o[q].e[e].re;
                        v.e[e].im += o[q].e[e].im;
                    }
                    out[(row_lo ^ s_comb_hi[q]) * rowvecs + vcol] = v;
                }
            }
            else
            {
#pragma unroll
                for (int q = 0; q < STEPS; ++q)
                    out[(row_lo ^ s_comb_hi[q]) * rowvecs + vcol] = tile[(l_lo + q * RPS) * TWC + jv];
            }
            // the staged rows were read through the generic proxy; the TMA refill writes through the asynchronous one
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            few_group_sync(grp);
            if (l == 0)
                few_mbar_arrive(&s_empty[buf]);
        }
    }
}


// ================================================================ K3g: K3f for passes with ANY number of x-masks
// Same persistent producer / two-consumer-group structure and tile ring as K3f.  The row factors cannot live in
// registers (a pass of BASELINE config 3 has 25-50 x-masks), so they are formed per (tile, x-mask) from the pass'
// strings, which are staged in shared memory ONCE per CTA (the CTA is persistent: once per launch) -- the general coset
// kernel K3b re-stages them for every tile.  Work items are (coset pair, chunk of column tiles), dealt round-robin to
// the CTAs, so small registers with wide batches (16 qubits x 1024 columns: 128 coset pairs) still fill 148 SMs.
// Tile order inside an item: column tile ct of the chunk for group 0 (coset 2p), then for group 1 (coset 2p + 1), ...
constexpr uint32_t kGenMaxStrings = 768;
constexpr uint32_t kGenMaxGroups = 256;
constexpr size_t kGenMetaBytes = kGenMaxStrings * 24 + (kGenMaxGroups + 1) * 4 + kGenMaxGroups * 4;

// The strings of one pass as a (large) kernel-parameter block: the row-factor loop then reads the constant bank only.
// ~21 KiB for complex128: needs the 32 KiB parameter space of CUDA >= 12.1 on sm_70+ (the library targets sm_100a).
template <typename T> struct GenStrings
{
    Cx<T> c[kGenMaxStrings];
    uint32_t z[kGenMaxStrings]; // <= 30 qubits: one word
    uint16_t gs[kGenMaxGroups + 1];
    uint8_t gxl[kGenMaxGroups];
};

template <typename T, int EPV, bool CMETA = false>
__global__ void __launch_bounds__(kFewTmaThreads, 1)
    coset_gen_tma_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, uint64_t nPairs, uint32_t chunkTiles,
                         CVec<T, EPV> *__restrict__ out, int beta, const __grid_constant__ CUtensorMap tm_in,
                         const __grid_constant__ GenStrings<T> gstr)
{
    using Vec = CVec<T, EPV>;
    constexpr int TWC = 16, RPS = 16, STEPS = 16, R = 8;
    constexpr uint32_t ROW_SHIFT = 8;

    extern __shared__ __align__(1024) unsigned char smem_gt[];
    __shared__ uint64_t s_full[kFewTmaBufs], s_empty[kFewTmaBufs];
    __shared__ uint32_t s_comb[256];
    __shared__ uint64_t s_comb_hi[STEPS];
    unsigned char *const meta = smem_gt + kFewTmaBufs * kFewTmaTile;
    uint64_t *const s_z = reinterpret_cast<uint64_t *>(meta);                              // [S]
    Cx<double> *const s_c = reinterpret_cast<Cx<double> *>(meta + kGenMaxStrings * 8);     // [S] (as double pairs)
    uint32_t *const s_gs = reinterpret_cast<uint32_t *>(meta + kGenMaxStrings * 24);       // [G + 1]
    uint32_t *const s_gxl = s_gs + kGenMaxGroups + 1;                                      // [G]

    uint32_t const tid = threadIdx.x;
    uint32_t const ng = pass.n_groups;
    uint32_t const n_str = pass.gstart[ng];
    if (tid < 256)
        s_comb[tid] = static_cast<uint32_t>(comb_of<R>(pass.basis, tid));
    if (tid < STEPS)
        s_comb_hi[tid] = comb_of<R>(pass.basis, tid * RPS);
    if (!CMETA)
    {
        for (uint32_t s = tid; s < n_str; s += blockDim.x)
        {
            s_z[s] = pass.sz[s];
            Cx<T> const c = pass.scoef[s];
            s_c[s] = Cx<double>{static_cast<double>(c.re), static_cast<double>(c.im)};
        }
        for (uint32_t g = tid; g <= ng; g += blockDim.x)
        {
            s_gs[g] = pass.gstart[g];
            if (g < ng)
                s_gxl[g] = pass.gxl[g];
        }
    }
    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < kFewTmaBufs; ++b)
        {
            few_mbar_init(&s_full[b], 1);
            few_mbar_init(&s_empty[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t const n_chunks = nColTiles / chunkTiles; // the host picks a divisor
    uint64_t const n_items = nPairs * n_chunks;
    uint64_t const my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint64_t const tiles_per_item = 2ull * chunkTiles;
    uint64_t const n_tiles = my_items * tiles_per_item;

    if (tid >= 512)
    {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid >= 544)
            return;
        uint32_t const lane = tid & 31u;
        for (uint64_t i = 0; i < n_tiles; ++i)
        {
            uint32_t const buf = static_cast<uint32_t>(i % kFewTmaBufs);
            if (i >= kFewTmaBufs)
                few_mbar_wait(&s_empty[buf], static_cast<uint32_t>((i / kFewTmaBufs) - 1) & 1u);
            uint64_t const item = blockIdx.x + (i / tiles_per_item) * gridDim.x;
            uint32_t const within = static_cast<uint32_t>(i % tiles_per_item);
            uint64_t const coset = 2 * (item / n_chunks) + (within & 1u);
            uint32_t const ct = static_cast<uint32_t>(item % n_chunks) * chunkTiles + (within >> 1);
            uint32_t const base = static_cast<uint32_t>(deposit_bits(coset, pass.nonpivot_mask));
            if (lane == 0)
                few_mbar_expect_tx(&s_full[buf], static_cast<uint32_t>(kFewTmaTile));
            __syncwarp();
            int const c0 = static_cast<int>(ct) * TWC * static_cast<int>(16 / sizeof(T));
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                uint32_t const op = lane + 32 * h;
                few_tma_gather4(smem_gt + buf * kFewTmaTile + (static_cast<size_t>(op) << (ROW_SHIFT + 2)), &tm_in, c0,
                                base ^ s_comb[4 * op], base ^ s_comb[4 * op + 1], base ^ s_comb[4 * op + 2],
                                base ^ s_comb[4 * op + 3], &s_full[buf]);
            }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    uint32_t const grp = tid >> 8;
    uint32_t const l = tid & 255u;
    uint32_t const l_lo = l >> 4, jv = l & 15u;
    uint32_t const own_off = (l << ROW_SHIFT) | ((l & 15u) << 4);

    for (uint64_t k = 0; k < my_items; ++k)
    {
        uint64_t const item = blockIdx.x + k * gridDim.x;
        uint64_t const coset = 2 * (item / n_chunks) + grp;
        uint32_t const ct0 = static_cast<uint32_t>(item % n_chunks) * chunkTiles;
        uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
        uint32_t const my_row = static_cast<uint32_t>(base) ^ s_comb[l]; // launched for <= 30 qubits only
        uint64_t const row_lo = base ^ s_comb[l_lo];
        for (uint32_t q = 0; q < chunkTiles; ++q)
        {
            uint32_t const ct = ct0 + q;
            uint64_t const i = k * tiles_per_item + 2ull * q + grp;
            uint32_t const buf = static_cast<uint32_t>(i % kFewTmaBufs);
            unsigned char *const tb = smem_gt + buf * kFewTmaTile;
            Vec *const tile = reinterpret_cast<Vec *>(tb);
            if (beta)
            {
                // accumulating pass: pull the old output rows of this tile into L2 now, so that the read-modify-write
                // loads of the store phase (many thousand cycles from here) do not wait for HBM
                uint64_t const vc = static_cast<uint64_t>(ct) * TWC + jv;
#pragma unroll
                for (int t = 0; t < STEPS; ++t)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(&out[(row_lo ^ s_comb_hi[t]) * rowvecs + vc]));
            }
            few_mbar_wait(&s_full[buf], static_cast<uint32_t>(i / kFewTmaBufs) & 1u);

            Cx<T> acc[TWC][EPV];
#pragma unroll
            for (int j = 0; j < TWC; ++j)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[j][e] = Cx<T>{0, 0};
            for (uint32_t g = 0; g < ng; ++g)
            {
                Cx<T> d{0, 0};
                uint32_t xl;
                if (CMETA)
                {
                    uint32_t const s1 = gstr.gs[g + 1];
                    for (uint32_t s = gstr.gs[g]; s < s1; ++s)
                    {
                        Cx<T> const c = gstr.c[s];
                        uint32_t const odd = __popc(my_row & gstr.z[s]) & 1u;
                        d.re += flip_sign(c.re, odd);
                        d.im += flip_sign(c.im, odd);
                    }
                    xl = gstr.gxl[g];
                }
                else
                {
                    uint32_t const s1 = s_gs[g + 1];
                    for (uint32_t s = s_gs[g]; s < s1; ++s)
                    {
                        Cx<double> const c = s_c[s];
                        uint32_t const odd = __popc(my_row & static_cast<uint32_t>(s_z[s])) & 1u;
                        d.re += flip_sign(static_cast<T>(c.re), odd);
                        d.im += flip_sign(static_cast<T>(c.im), odd);
                    }
                    xl = s_gxl[g];
                }
                uint32_t const src = own_off ^ (xl << ROW_SHIFT);
#pragma unroll
                for (int j = 0; j < TWC; ++j)
                {
                    Vec const v = *reinterpret_cast<Vec const *>(tb + (src ^ (j << 4)));
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[j][e], d, v.e[e]);
                }
            }
            few_group_sync(grp);
#pragma unroll
            for (int j = 0; j < TWC; ++j)
            {
                Vec v;
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    v.e[e] = acc[j][e];
                *reinterpret_cast<Vec *>(tb + (own_off ^ (j << 4))) = v;
            }
            few_group_sync(grp);
            uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jv;
            if (beta)
            {
                Vec o[STEPS];
#pragma unroll
                for (int t = 0; t < STEPS; ++t)
                    o[t] = out[(row_lo ^ s_comb_hi[t]) * rowvecs + vcol];
#pragma unroll
                for (int t = 0; t < STEPS; ++t)
                {
                    Vec v = tile[(l_lo + t * RPS) * TWC + jv];
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        v.e[e].re += o[t].e[e].re;
                        v.e[e].im += o[t].e[e].im;
                    }
                    out[(row_lo ^ s_comb_hi[t]) * rowvecs + vcol] = v;
                }
            }
            else
            {
#pragma unroll
                for (int t = 0; t < STEPS; ++t)
                    out[(row_lo ^ s_comb_hi[t]) * rowvecs + vcol] = tile[(l_lo + t * RPS) * TWC + jv];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            few_group_sync(grp);
            if (l == 0)
                few_mbar_arrive(&s_empty[buf]);
        }
    }
}

} // namespace fpk
