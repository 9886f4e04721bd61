// K3d: dense-coset tensor-core kernel for complex128 operators whose x-masks span a GF(2) subspace of rank 4 or 5
// (every k-local dense term with k = 4, 5: up to 1024 strings over 32 masks).
//
// On one coset the operator is a dense 2^RR x 2^RR complex matrix M_c (entry (l', l) = D[l ^ l'][l'], the same row
// factors as rcoset.cuh) applied to 2^RR rows x all batch columns -- GEMM-shaped work, so it runs on the FP64
// tensor cores (mma.sync.m8n8k4.f64: measured 63.7 FMA/clk/SM on B200, the full FP64 rate) with M_c held in
// REGISTERS as A fragments.  The SIMT register kernel (K3c) is bound at rank 4 by one broadcast LDS.128 of D per
// complex FMA on the SM's single load/store unit; here D is read once per coset and warp.
//
// Real formulation: [out_re; out_im] = [[Re M, -Im M], [Im M, Re M]] [x_re; x_im], planar ordering (all real parts,
// then all imaginary parts) on both sides, so that
//   * a lane's 16-byte global load psi(row 4*kb + lane%4, column n0 + lane/4) feeds two B fragments (k-block kb with
//     its real part, k-block kb + ROWS/4 with its imaginary part): every LDG.128 instruction moves four contiguous
//     128-byte row segments;
//   * the real and imaginary parts of an output element land in the same lane (m-tiles rt and rt + ROWS/8), and one
//     pair exchange per row tile lets every STG.128 instruction write whole 32-byte sectors.
// WPC warps share one coset: each owns ROWS/8/WPC row tiles of the output and loads the full input tile.
//
// Covers PauliOp::apply (PO:399-468) / SummedPauliOp::apply (SPO:277-349) for std::complex<double> -- and, through the
// TIO = float instance, for std::complex<float> batches: the 16-byte load then holds TWO batch columns, which feed two
// n-tiles of MMAs (even / odd columns) after an exact float -> double conversion; arithmetic and row factors stay
// FP64 and the result is rounded to float once, in the store.  The complex64 SIMT forms of the same operators are
// bound by the load/store unit (all 1024 Paulis on 5 qubits, 20 q x 64: 1.53 ms through the shared-memory coset
// kernel against 0.93 ms for the SAME operator on a complex128 batch of twice the bytes through this kernel).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rcoset.cuh"

namespace fpk
{

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// batch element type of the global arrays: the 16-byte vector of one lane and the n-tiles (batch columns) it carries
template <typename TIO> struct DcIo;
template <> struct DcIo<double>
{
    using V = double2;
    static constexpr int NH = 1;
};
template <> struct DcIo<float>
{
    using V = float4;
    static constexpr int NH = 2;
};
__device__ __forceinline__ double dc_re(double2 const &v, int) { return v.x; }
__device__ __forceinline__ double dc_im(double2 const &v, int) { return v.y; }
__device__ __forceinline__ double dc_re(float4 const &v, int h) { return h ? v.z : v.x; }
__device__ __forceinline__ double dc_im(float4 const &v, int h) { return h ? v.w : v.y; }
template <typename V> __device__ __forceinline__ V dc_zero();
template <> __device__ __forceinline__ double2 dc_zero<double2>() { return make_double2(0, 0); }
template <> __device__ __forceinline__ float4 dc_zero<float4>() { return make_float4(0, 0, 0, 0); }

// barrier among the WPC warps that share a coset (named barrier 1 + group index; a single warp needs __syncwarp only)
template <int WPC> __device__ __forceinline__ void group_sync(uint32_t group)
{
    if (WPC == 1)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;\n" ::"r"(group + 1), "n"(WPC * 32) : "memory");
}

template <int RR, int WPC> struct DcosetCfg
{
    static constexpr int NT = 128, WARPS = NT / 32;
    static constexpr int ROWS = 1 << RR;
    static constexpr int CPI = WARPS / WPC;          // cosets per CTA iteration
    static constexpr int RT = ROWS / 8;              // row tiles of the output (real part; the same again imaginary)
    static constexpr int RT_OWN = RT / WPC;          // row tiles per warp
    static constexpr int KB = ROWS / 4;              // k-blocks per part (real / imaginary)
    static constexpr int PITCH = ROWS + 1;           // table row pitch (entries): column accesses stay conflict-free
    static constexpr size_t tables = 2 * static_cast<size_t>(CPI) * ROWS * PITCH * sizeof(Cx<double>); // D and U tables
    static_assert(WPC * 32 >= 2 * ROWS, "the sign transform maps one (local x-mask, re/im) pair to a thread");
    static constexpr int MAX_STAGED = 1024; // strings whose (coefficient, z-mask) are staged in shared memory
    static constexpr size_t smem = tables + MAX_STAGED * (sizeof(Cx<double>) + 8) + (ROWS * ROWS + 1) * 4 + 12;
    static constexpr int ECOLS = 128; // MODE 1: 16-byte batch vectors per CTA (blockIdx.y picks the range)
    // per-warp column sums: one entry per batch column (two per vector for complex64 batches)
    static constexpr size_t smem_expval(int nh)
    {
        return (smem + 15) / 16 * 16 + static_cast<size_t>(WARPS) * ECOLS * nh * sizeof(Cx<double>);
    }
    static_assert(RT % WPC == 0 && WARPS % WPC == 0, "warps must split the row tiles evenly");
};

// MODE 0: apply (store / accumulate).  MODE 1: expectation-value partials (PauliOp::expectation_value, PO:482-549):
// every warp sums conj(psi) . (M_c psi) per batch column over its cosets into a private shared-memory array, the CTA
// folds its warps at the end and writes one partial row per CTA: partials[blockIdx.x][column] (fixed order:
// deterministic); blockIdx.y selects a range of ECOLS columns.
template <int RR, int WPC, int PFD, int MODE, typename TIO = double>
__global__ void __launch_bounds__(128)
    dcoset_kernel(RcPassView<TIO> pass, uint32_t n_strings, uint64_t n_cosets, uint64_t rowvecs_total,
                  CVec<TIO, 16 / (2 * sizeof(TIO))> const *__restrict__ in, CVec<TIO, 16 / (2 * sizeof(TIO))> *__restrict__ out,
                  int beta, Cx<double> *__restrict__ partials, uint32_t Bpad)
{
    using Cfg = DcosetCfg<RR, WPC>;
    using V = typename DcIo<TIO>::V;
    constexpr int NH = DcIo<TIO>::NH;
    constexpr int ROWS = Cfg::ROWS, CPI = Cfg::CPI, PITCH = Cfg::PITCH, RT_OWN = Cfg::RT_OWN, KB = Cfg::KB;
    extern __shared__ __align__(16) unsigned char dc_smem[];
    Cx<double> *Dt = reinterpret_cast<Cx<double> *>(dc_smem); // [CPI][xl][l], rows PITCH entries apart
    Cx<double> *Ut = Dt + CPI * ROWS * PITCH;                 // [CPI][xl][zl]
    // operator metadata staged once per (persistent) CTA: the per-coset table build then never waits on global loads
    Cx<double> *s_coef = Ut + CPI * ROWS * PITCH;
    uint64_t *s_z = reinterpret_cast<uint64_t *>(s_coef + Cfg::MAX_STAGED);
    uint32_t *s_ustart = reinterpret_cast<uint32_t *>(s_z + Cfg::MAX_STAGED);
    bool const staged = n_strings <= static_cast<uint32_t>(Cfg::MAX_STAGED);
    uint64_t const *zs = pass.sz;
    uint32_t const *ustart = pass.ustart;
    if (staged)
    {
        for (uint32_t i = threadIdx.x; i < n_strings; i += Cfg::NT)
        {
            Cx<TIO> const c = pass.scoef[i];
            s_coef[i] = Cx<double>{static_cast<double>(c.re), static_cast<double>(c.im)};
            s_z[i] = pass.sz[i];
        }
        for (uint32_t i = threadIdx.x; i <= ROWS * ROWS; i += Cfg::NT)
            s_ustart[i] = pass.ustart[i];
        zs = s_z;
        ustart = s_ustart;
        __syncthreads();
    }
    auto coef_at = [&](uint32_t s) {
        if (staged)
            return s_coef[s];
        Cx<TIO> const c = pass.scoef[s];
        return Cx<double>{static_cast<double>(c.re), static_cast<double>(c.im)};
    };

    uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t const cw = warp / WPC, part = warp % WPC; // coset slot of this warp, its share of the row tiles
    uint32_t const lq = lane >> 2, lr = lane & 3u;
    // MODE 1 CTAs work on the column range [col0, col0 + rowvecs); the row pitch is always rowvecs_total
    uint64_t const col0 = MODE == 1 ? static_cast<uint64_t>(blockIdx.y) * Cfg::ECOLS : 0;
    uint64_t const rowvecs = MODE == 1 ? (rowvecs_total - col0 < Cfg::ECOLS ? rowvecs_total - col0 : Cfg::ECOLS) : rowvecs_total;
    V const *in2 = reinterpret_cast<V const *>(in) + col0;
    V *out2 = reinterpret_cast<V *>(out);
    Cx<double> *esm = reinterpret_cast<Cx<double> *>(dc_smem + (Cfg::smem + 15) / 16 * 16) + warp * Cfg::ECOLS * NH;
    if (MODE == 1)
    {
        for (uint32_t i = lane; i < Cfg::ECOLS * NH; i += 32)
            esm[i] = Cx<double>{0, 0};
        __syncwarp();
    }

    constexpr int PF = PFD; // input tiles in flight per warp (one 8-column tile of MMAs is ~270 ns, an HBM load ~800 ns)
    uint64_t const n_tiles = (rowvecs + 7) / 8;

    for (uint64_t cs0 = static_cast<uint64_t>(blockIdx.x) * CPI; cs0 < n_cosets; cs0 += static_cast<uint64_t>(gridDim.x) * CPI)
    {
        uint64_t const coset = cs0 + cw;
        bool const live = coset < n_cosets; // warp-uniform
        uint64_t const base = rc_base<RR>(live ? coset : 0, pass.pivot);
        uint64_t rowoff[KB]; // element offset of input row 4*kb + lane%4
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
        {
            uint64_t comb = 0;
#pragma unroll
            for (int k = 0; k < RR; ++k)
                if (((4 * kb + lr) >> k) & 1u)
                    comb ^= pass.basis[k];
            rowoff[kb] = (base ^ comb) * rowvecs_total;
        }
        // ---- the first PF input tiles are requested before the factor tables are built: their latency hides there
        V x[PF][KB];
#pragma unroll
        for (int u = 0; u < PF; ++u)
        {
            uint64_t const col = static_cast<uint64_t>(u) * 8 + lq;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
                x[u][kb] = (live && col < rowvecs) ? in2[rowoff[kb] + col] : dc_zero<V>();
        }

        // ---- row factors of this warp group's coset (same two-step build as rcoset.cuh).  Every group of WPC warps
        // builds its own table and synchronises only within itself, so the groups of a CTA drift freely through
        // their table and tensor-core phases instead of meeting at CTA-wide barriers.
        Cx<double> *Dc = Dt + cw * ROWS * PITCH;
        Cx<double> *Uc = Ut + cw * ROWS * PITCH;
        uint32_t const gtid = part * 32 + lane; // thread index inside the group
        if (live)
        {
            for (uint32_t xz = gtid; xz < ROWS * ROWS; xz += WPC * 32)
            {
                Cx<double> u{0, 0};
                uint32_t const s0 = ustart[xz], s1 = ustart[xz + 1];
                for (uint32_t s = s0; s < s1; ++s)
                {
                    Cx<double> const cf = coef_at(s);
                    uint32_t const odd = parity64(base & zs[s]);
                    u.re += flip_sign(cf.re, odd);
                    u.im += flip_sign(cf.im, odd);
                }
                Uc[(xz >> RR) * PITCH + (xz & (ROWS - 1))] = u;
            }
        }
        group_sync<WPC>(cw);
        if (live && gtid < 2 * ROWS)
        {
            // D[xl][.] = Walsh-Hadamard transform of U[xl][.]: one (xl, re/im) pair per thread, in registers
            uint32_t const xl = gtid >> 1, ri = gtid & 1u;
            double const *Ud = reinterpret_cast<double const *>(Uc + xl * PITCH) + ri;
            double v[ROWS];
#pragma unroll
            for (int z = 0; z < ROWS; ++z)
                v[z] = Ud[2 * z];
#pragma unroll
            for (int h = 1; h < ROWS; h <<= 1)
#pragma unroll
                for (int i = 0; i < ROWS; ++i)
                    if ((i & h) == 0)
                    {
                        double const a = v[i], b = v[i + h];
                        v[i] = a + b;
                        v[i + h] = a - b;
                    }
            double *Dd = reinterpret_cast<double *>(Dc + xl * PITCH) + ri;
#pragma unroll
            for (int l = 0; l < ROWS; ++l)
                Dd[2 * l] = v[l];
        }
        group_sync<WPC>(cw);

        if (live)
        {
            // ---- A fragments of M_c for this warp's m-tiles: row kk' = 8*mt + lane/4, column kk = 4*kb + lane%4
            double A[2 * RT_OWN][2 * KB];
#pragma unroll
            for (int r = 0; r < RT_OWN; ++r)
            {
                uint32_t const lp = 8 * (part * RT_OWN + r) + lq; // output row l'
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
                {
                    uint32_t const l = 4 * kb + lr; // input row
                    Cx<double> const m = Dc[(l ^ lp) * PITCH + lp];
                    A[r][kb] = m.re;               // Re block
                    A[r][KB + kb] = -m.im;         // -Im block
                    A[RT_OWN + r][kb] = m.im;      // Im block
                    A[RT_OWN + r][KB + kb] = m.re; // Re block
                }
            }
            uint64_t orow[RT_OWN]; // element offset of output row 8*rt + lane/4
#pragma unroll
            for (int r = 0; r < RT_OWN; ++r)
            {
                uint32_t const lp = 8 * (part * RT_OWN + r) + lq;
                uint64_t comb = 0;
#pragma unroll
                for (int k = 0; k < RR; ++k)
                    if ((lp >> k) & 1u)
                        comb ^= pass.basis[k];
                orow[r] = (base ^ comb) * rowvecs_total;
            }

            // ---- sweep the batch in tiles of 8 columns through a ring of PF register buffers: the buffer a tile has
            // just consumed is refilled with the tile PF steps ahead
            for (uint64_t nt0 = 0; nt0 < n_tiles; nt0 += PF)
            {
#pragma unroll
                for (int u = 0; u < PF; ++u)
                {
                    uint64_t const nt = nt0 + u;
                    if (nt >= n_tiles)
                        break; // warp-uniform
                    uint64_t const n0 = nt * 8;
                    double C[NH][2 * RT_OWN][2];
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int m = 0; m < 2 * RT_OWN; ++m)
                            C[h][m][0] = C[h][m][1] = 0.0;
                    // consecutive MMAs go to different accumulators (the m-tiles), never back to back into one
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                    {
#pragma unroll
                        for (int h = 0; h < NH; ++h)
                        {
                            double const xr = dc_re(x[u][kb], h), xi = dc_im(x[u][kb], h);
#pragma unroll
                            for (int m = 0; m < 2 * RT_OWN; ++m)
                                dmma884(C[h][m][0], C[h][m][1], A[m][kb], xr);
#pragma unroll
                            for (int m = 0; m < 2 * RT_OWN; ++m)
                                dmma884(C[h][m][0], C[h][m][1], A[m][KB + kb], xi);
                        }
                    }
                    if (nt + PF < n_tiles)
                    {
                        uint64_t const col = n0 + PF * 8 + lq;
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb)
                            x[u][kb] = col < rowvecs ? in2[rowoff[kb] + col] : dc_zero<V>();
                    }
                    if (MODE == 1)
                    {
                        // e(col) += sum over this lane's rows of conj(psi(l', col)) * out(l', col); the psi values are
                        // re-read (cache hits: the tile was just loaded), then the 8 row lanes are folded by shuffles
                        double er[2][NH], ei[2][NH];
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int h = 0; h < NH; ++h)
                                er[c][h] = ei[c][h] = 0.0;
#pragma unroll
                        for (int r = 0; r < RT_OWN; ++r)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                            {
                                uint64_t const col = n0 + 2 * lr + c;
                                V const a = col < rowvecs ? in2[orow[r] + col] : dc_zero<V>();
#pragma unroll
                                for (int h = 0; h < NH; ++h)
                                {
                                    double const ar = dc_re(a, h), ai = dc_im(a, h);
                                    double const o_re = C[h][r][c], o_im = C[h][RT_OWN + r][c];
                                    er[c][h] = fma(ar, o_re, er[c][h]);
                                    er[c][h] = fma(ai, o_im, er[c][h]);
                                    ei[c][h] = fma(ar, o_im, ei[c][h]);
                                    ei[c][h] = fma(-ai, o_re, ei[c][h]);
                                }
                            }
#pragma unroll
                        for (int off = 4; off < 32; off <<= 1)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int h = 0; h < NH; ++h)
                                {
                                    er[c][h] += __shfl_xor_sync(0xffffffffu, er[c][h], off);
                                    ei[c][h] += __shfl_xor_sync(0xffffffffu, ei[c][h], off);
                                }
                        if (lq == 0)
                        {
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                            {
                                uint64_t const col = n0 + 2 * lr + c;
                                if (col < rowvecs)
                                {
#pragma unroll
                                    for (int h = 0; h < NH; ++h)
                                    {
                                        esm[col * NH + h].re += er[c][h];
                                        esm[col * NH + h].im += ei[c][h];
                                    }
                                }
                            }
                        }
                    }
                    else
                    {
                    // lane holds rows l' (real: C[.][r], imaginary: C[.][RT_OWN + r]) x vectors n0 + 2*lr + {0, 1}
                    if constexpr (NH == 1)
                    {
#pragma unroll
                    for (int r = 0; r < RT_OWN; ++r)
                    {
                        double2 first = make_double2(C[0][r][0], C[0][RT_OWN + r][0]);  // column 2*lr
                        double2 second = make_double2(C[0][r][1], C[0][RT_OWN + r][1]); // column 2*lr + 1
                        // pair exchange: even lanes end up with columns (4p, 4p+2), odd lanes with (4p+1, 4p+3), so
                        // each store instruction writes adjacent columns from adjacent lanes = whole 32-byte sectors
                        bool const oddl = lr & 1u;
                        double2 const send = oddl ? first : second;
                        double2 recv;
                        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                        double2 va = oddl ? recv : first;  // column 2*(lr & ~1) + (lr & 1)
                        double2 vb = oddl ? second : recv; // that column + 2
                        uint64_t const ca = n0 + 2 * (lr & ~1u) + (lr & 1u), cb = ca + 2;
                        if (ca < rowvecs)
                        {
                            V *dst = &out2[orow[r] + ca];
                            if (beta)
                            {
                                V const o = *dst;
                                va.x += o.x;
                                va.y += o.y;
                            }
                            *dst = va;
                        }
                        if (cb < rowvecs)
                        {
                            V *dst = &out2[orow[r] + cb];
                            if (beta)
                            {
                                V const o = *dst;
                                vb.x += o.x;
                                vb.y += o.y;
                            }
                            *dst = vb;
                        }
                    }
                    }
                    else
                    {
                        // complex64: both columns of a 16-byte vector sit in this lane (n-tiles 0 / 1), the lane's two
                        // vectors are neighbours: 32 contiguous bytes per lane, 128 per row and instruction pair
#pragma unroll
                        for (int r = 0; r < RT_OWN; ++r)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                            {
                                uint64_t const col = n0 + 2 * lr + c;
                                if (col < rowvecs)
                                {
                                    V *dst = &out2[orow[r] + col];
                                    double v0 = C[0][r][c], v1 = C[0][RT_OWN + r][c], v2 = C[NH - 1][r][c],
                                           v3 = C[NH - 1][RT_OWN + r][c];
                                    if (beta)
                                    {
                                        V const o = *dst;
                                        v0 += dc_re(o, 0);
                                        v1 += dc_im(o, 0);
                                        v2 += dc_re(o, 1);
                                        v3 += dc_im(o, 1);
                                    }
                                    float4 const w = make_float4(static_cast<float>(v0), static_cast<float>(v1),
                                                                 static_cast<float>(v2), static_cast<float>(v3));
                                    *reinterpret_cast<float4 *>(dst) = w;
                                }
                            }
                    }
                    }
                }
            }
        }
        group_sync<WPC>(cw); // the group's tables are rebuilt by its next iteration
    }
    if (MODE == 1)
    {
        __syncthreads();
        Cx<double> const *all = reinterpret_cast<Cx<double> const *>(dc_smem + (Cfg::smem + 15) / 16 * 16);
        for (uint32_t i = tid; i < rowvecs * NH; i += Cfg::NT)
        {
            Cx<double> sum{0, 0};
#pragma unroll
            for (int w = 0; w < Cfg::WARPS; ++w)
            {
                sum.re += all[w * Cfg::ECOLS * NH + i].re;
                sum.im += all[w * Cfg::ECOLS * NH + i].im;
            }
            partials[static_cast<uint64_t>(blockIdx.x) * Bpad + col0 * NH + i] = sum;
        }
    }
}

} // namespace fpk
