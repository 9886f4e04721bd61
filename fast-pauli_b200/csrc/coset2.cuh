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
#include <cuda_runtime.h>

#include "coset.cuh"

namespace fpk
{

template <int LOG_TWC> struct FewCfg
{
    static constexpr int NT = 256;                  // threads = rows of the tile
    static constexpr int R = 8;                     // tile rank
    static constexpr int TWC = 1 << LOG_TWC;        // vectors per row segment
    static constexpr int RPS = NT >> LOG_TWC;       // rows covered by one cooperative load/store step
    static constexpr int STEPS = NT / RPS;          // = TWC
    static constexpr size_t TILE_BYTES = static_cast<size_t>(NT) * TWC * 16;
    static_assert(LOG_TWC == 3 || LOG_TWC == 4, "row segments of 128 or 256 bytes");
};

template <typename T, int EPV, int LOG_TWC, int GMAX>
__global__ void __launch_bounds__(256, 2)
    coset_few_kernel(CosetPassView<T> pass, uint64_t rowvecs, uint32_t nColTiles, uint32_t ctPerCta, uint32_t nCtGroups,
                     CVec<T, EPV> const *__restrict__ in, CVec<T, EPV> *__restrict__ out, int beta)
{
    using Cfg = FewCfg<LOG_TWC>;
    using Vec = CVec<T, EPV>;
    constexpr int TWC = Cfg::TWC, RPS = Cfg::RPS, STEPS = Cfg::STEPS;
    constexpr uint32_t ROW_SHIFT = LOG_TWC + 4; // bytes per tile row = 1 << ROW_SHIFT

    extern __shared__ __align__(1024) unsigned char smem_few[];
    __shared__ uint64_t s_comb_hi[STEPS];
    __shared__ uint32_t s_gxl[GMAX];
    Vec *tile = reinterpret_cast<Vec *>(smem_few);

    uint32_t const tid = threadIdx.x;
    uint32_t const ng = pass.n_groups;
    uint64_t const coset = blockIdx.x / nCtGroups;
    uint32_t const ctg = static_cast<uint32_t>(blockIdx.x - coset * nCtGroups);
    uint64_t const base = deposit_bits(coset, pass.nonpivot_mask);
    uint32_t ct = ctg * ctPerCta;
    uint32_t const ct_end = min(nColTiles, ct + ctPerCta);

    if (tid < STEPS)
        s_comb_hi[tid] = comb_of<Cfg::R>(pass.basis, tid * RPS);
    if (tid < ng)
        s_gxl[tid] = pass.gxl[tid];
    uint32_t const l_lo = tid >> LOG_TWC;
    uint32_t const jv = tid & (TWC - 1);
    uint64_t const row_lo = base ^ comb_of<Cfg::R>(pass.basis, l_lo);
    uint64_t const my_row = base ^ comb_of<Cfg::R>(pass.basis, tid);
    __syncthreads();

    auto fill = [&](uint32_t c) {
        uint64_t const vcol = static_cast<uint64_t>(c) * TWC + jv;
#pragma unroll
        for (int k = 0; k < STEPS; ++k)
        {
            uint64_t const row = row_lo ^ s_comb_hi[k];
            cp_async16(&tile[(l_lo + k * RPS) * TWC + jv], &in[row * rowvecs + vcol]);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    fill(ct);

    // ---- row factors of this thread's row, once per coset (the first tile is in flight meanwhile)
    Cx<T> D[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
    {
        D[g] = Cx<T>{0, 0};
        if (static_cast<uint32_t>(g) < ng)
        {
            uint32_t const s0 = pass.gstart[g], s1 = pass.gstart[g + 1];
            for (uint32_t s = s0; s < s1; ++s)
            {
                Cx<T> const c = pass.scoef[s];
                uint32_t const odd = parity64(my_row & pass.sz[s]);
                D[g].re += flip_sign(c.re, odd);
                D[g].im += flip_sign(c.im, odd);
            }
        }
    }

    uint32_t const key_off = (tid & (TWC - 1)) << 4;             // column rotation of this thread, in bytes
    uint32_t const own_off = (tid << ROW_SHIFT) | key_off;       // own row, rotated column 0

    for (; ct < ct_end; ++ct)
    {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();

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
#pragma unroll
                for (int j = 0; j < TWC; ++j)
                {
                    Vec const v = *reinterpret_cast<Vec const *>(smem_few + (src ^ (j << 4)));
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[j][e], D[g], v.e[e]);
                }
            }
        }
        __syncthreads(); // every gather of this tile is done: the buffer becomes the store staging area

#pragma unroll
        for (int j = 0; j < TWC; ++j)
        {
            Vec v;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                v.e[e] = acc[j][e];
            *reinterpret_cast<Vec *>(smem_few + (own_off ^ (j << 4))) = v;
        }
        __syncthreads();

        uint64_t const vcol = static_cast<uint64_t>(ct) * TWC + jv;
        if (beta)
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
        if (ct + 1 < ct_end)
        {
            __syncthreads(); // staged rows have been read: refill the buffer with the next column tile
            fill(ct + 1);
        }
    }
}

} // namespace fpk
