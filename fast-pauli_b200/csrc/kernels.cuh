// Hand-written sm_100a kernels of the fast-pauli hot path.
//
// Everything here evaluates the PauliComposer closed form on the fly --
//     (P psi)(i,t) = (-i)^nY (-1)^popcount(i & z) psi(i ^ x, t)
// (reference: get_sparse_repr, __pauli_string.hpp:49-118) -- with no materialised
// (k, m) tables.  State batches are row-major (dim, n_states), batch axis
// contiguous, and every global access is a 16-byte vector along that axis
// (1 complex128 or 2 complex64 per lane; an 8-byte variant covers complex64
// rows that are not 16-byte multiples).
//
// Thread geometry shared by all kernels ("Geom"): a CTA is 256 threads laid
// out as TW lanes along the row (TW = 2^log2TW consecutive 16-byte vectors)
// times TY = 256/TW rows; each thread owns V rows (stride TY) of one vector
// column, so a CTA covers TY*V rows x TW vectors.  The 1-D grid enumerates row
// blocks fastest and column tiles slowest, so concurrently resident CTAs sweep
// all rows of ONE narrow column tile: for multi-group operators the host picks
// TW such that dim x TW x 16 B fits the L2 budget and every gather after the
// first touch is an L2 hit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fpk
{

// ---------------------------------------------------------------- small numeric helpers
template <typename T> struct alignas(2 * sizeof(T)) Cx
{
    T re, im;
};

// EPV complex elements moved as one aligned vector (16 B, or 8 B for <float,1>)
template <typename T, int EPV> struct alignas(2 * sizeof(T) * EPV) CVec
{
    Cx<T> e[EPV];
};

__device__ __forceinline__ double flip_sign(double v, uint32_t odd)
{
    return __longlong_as_double(__double_as_longlong(v) ^ (static_cast<long long>(odd) << 63));
}
__device__ __forceinline__ float flip_sign(float v, uint32_t odd)
{
    return __int_as_float(__float_as_int(v) ^ static_cast<int>(odd << 31));
}
__device__ __forceinline__ uint32_t parity64(uint64_t v)
{
    return static_cast<uint32_t>(__popcll(v)) & 1u;
}

template <typename T> __device__ __forceinline__ void cfma(Cx<T> &acc, Cx<T> a, Cx<T> b)
{
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(a.im, b.re, acc.im);
}

// ---------------------------------------------------------------- operator view / geometry
// Device view of a packed operator: strings sorted by (x, z), duplicates merged, grouped by x.
// scoef[s] already contains h_s * (-i)^nY_s, so only the popcount sign remains per row.
template <typename T> struct OpView
{
    uint64_t const *gx;     // [G]   x-mask of each group
    uint32_t const *gstart; // [G+1] first packed string of each group
    uint64_t const *sz;     // [S]   z-mask per packed string
    Cx<T> const *scoef;     // [S]   h_s * (-i)^nY
    uint32_t G;
    // single-string form passed by value (INLINE1 kernels): no device metadata at all
    uint64_t x0, z0;
    Cx<T> c0;
};

struct Geom
{
    uint64_t N;          // rows (= dim, or pair-rows for the paired expectation kernel)
    uint64_t rowvecs;    // vectors per row
    uint64_t nRowBlocks; // row blocks per column tile
    uint32_t nColTiles;
    uint32_t log2TW;
    uint32_t iters; // row iterations per CTA (reduction kernels)
    uint32_t Bpad;  // padded batch length of the partials rows (reduction kernels)
};

constexpr int kThreads = 256;

// FP64 peak probe (fp_measure_fp64_tflops): 16 independent DFMA chains per thread
static __global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
    double const a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        c[i] = i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- K1/K3: (grouped) operator apply
//   out(i,t) (+)= sum_g D_g(i) psi(i ^ x_g, t),  D_g(i) = sum_{s in g} scoef_s (-1)^popc(i & z_s)
// Covers PauliString::apply / apply_batch (PS:296-436; INLINE1, one group of one string),
// PauliOp::apply 1-D/2-D (PO:362-468) and SummedPauliOp::apply (SPO:277-349, with c_j = sum_k coeffs(j,k)
// folded on the host).
// Each thread owns V rows x J vectors (vectors strided by TW so every warp-level access is contiguous); the
// row-only factor D_g(i) is formed once per (row, group) and reused for the J vectors.
// MODE 0: store.  MODE 1: expectation partials  e(t) = sum_i conj(bra(i,t)) * (A psi)(i,t)
// (PauliOp::expectation_value, PO:482-549; bra = psi) reduced over the CTA's rows into partials[rb][t].
template <typename T, int EPV, int V, int J, int MODE, bool INLINE1>
__global__ void __launch_bounds__(kThreads)
    op_kernel(OpView<T> op, Geom g, CVec<T, EPV> const *__restrict__ in, CVec<T, EPV> *__restrict__ out,
              Cx<T> *__restrict__ partials, int beta, CVec<T, EPV> const *__restrict__ bra)
{
    using Vec = CVec<T, EPV>;
    uint32_t const TW = 1u << g.log2TW;
    uint32_t const TY = kThreads >> g.log2TW;
    uint32_t const lx = threadIdx.x & (TW - 1);
    uint32_t const ty = threadIdx.x >> g.log2TW;
    uint64_t const blk = blockIdx.x;
    uint64_t const ct = blk / g.nRowBlocks;
    uint64_t const rb = blk - ct * g.nRowBlocks;
    uint64_t vcol[J];
    bool vok[J];
#pragma unroll
    for (int j = 0; j < J; ++j)
    {
        uint64_t v = ct * (static_cast<uint64_t>(TW) * J) + lx + static_cast<uint64_t>(j) * TW;
        vok[j] = v < g.rowvecs;
        vcol[j] = vok[j] ? v : 0;
    }
    if (MODE == 0 && !vok[0])
        return;

    uint32_t const n_it = (MODE == 0) ? 1u : g.iters;
    Cx<T> esum[J][EPV];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            esum[j][e] = Cx<T>{0, 0};

    for (uint32_t it = 0; it < n_it; ++it)
    {
        uint64_t const row0 = (rb * n_it + it) * (static_cast<uint64_t>(TY) * V) + ty;
        uint64_t rows[V];
        bool rok[V];
#pragma unroll
        for (int k = 0; k < V; ++k)
        {
            uint64_t r = row0 + static_cast<uint64_t>(k) * TY;
            rok[k] = r < g.N;
            rows[k] = rok[k] ? r : (g.N - 1);
        }

        Cx<T> acc[V][J][EPV];
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[k][j][e] = Cx<T>{0, 0};

        if (INLINE1)
        {
            Vec src[V][J];
#pragma unroll
            for (int k = 0; k < V; ++k)
#pragma unroll
                for (int j = 0; j < J; ++j)
                    src[k][j] = in[(rows[k] ^ op.x0) * g.rowvecs + vcol[j]];
#pragma unroll
            for (int k = 0; k < V; ++k)
            {
                uint32_t odd = parity64(rows[k] & op.z0);
                Cx<T> d{flip_sign(op.c0.re, odd), flip_sign(op.c0.im, odd)};
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        cfma(acc[k][j][e], d, src[k][j].e[e]);
            }
        }
        else
        {
            uint32_t s0 = __ldg(op.gstart);
            for (uint32_t gi = 0; gi < op.G; ++gi)
            {
                uint64_t const x = __ldg(op.gx + gi);
                uint32_t const s1 = __ldg(op.gstart + gi + 1);
                Vec src[V][J];
#pragma unroll
                for (int k = 0; k < V; ++k)
#pragma unroll
                    for (int j = 0; j < J; ++j)
                        src[k][j] = in[(rows[k] ^ x) * g.rowvecs + vcol[j]];
                Cx<T> d[V];
#pragma unroll
                for (int k = 0; k < V; ++k)
                    d[k] = Cx<T>{0, 0};
                for (uint32_t s = s0; s < s1; ++s)
                {
                    uint64_t const z = __ldg(op.sz + s);
                    Cx<T> const c = op.scoef[s];
#pragma unroll
                    for (int k = 0; k < V; ++k)
                    {
                        uint32_t odd = parity64(rows[k] & z);
                        d[k].re += flip_sign(c.re, odd);
                        d[k].im += flip_sign(c.im, odd);
                    }
                }
#pragma unroll
                for (int k = 0; k < V; ++k)
#pragma unroll
                    for (int j = 0; j < J; ++j)
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            cfma(acc[k][j][e], d[k], src[k][j].e[e]);
                s0 = s1;
            }
        }

        if (MODE == 0)
        {
#pragma unroll
            for (int k = 0; k < V; ++k)
            {
                if (!rok[k])
                    continue;
#pragma unroll
                for (int j = 0; j < J; ++j)
                {
                    if (!vok[j])
                        continue;
                    uint64_t const o = rows[k] * g.rowvecs + vcol[j];
                    Vec r;
                    if (beta)
                    {
                        r = out[o];
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                        {
                            r.e[e].re += acc[k][j][e].re;
                            r.e[e].im += acc[k][j][e].im;
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int e = 0; e < EPV; ++e)
                            r.e[e] = acc[k][j][e];
                    }
                    out[o] = r;
                }
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < V; ++k)
            {
                if (!rok[k])
                    continue;
#pragma unroll
                for (int j = 0; j < J; ++j)
                {
                    if (!vok[j])
                        continue;
                    Vec a = bra[rows[k] * g.rowvecs + vcol[j]]; // bra == in except for sharded states
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        // conj(a) * acc
                        esum[j][e].re = fma(a.e[e].re, acc[k][j][e].re, esum[j][e].re);
                        esum[j][e].re = fma(a.e[e].im, acc[k][j][e].im, esum[j][e].re);
                        esum[j][e].im = fma(a.e[e].re, acc[k][j][e].im, esum[j][e].im);
                        esum[j][e].im = fma(-a.e[e].im, acc[k][j][e].re, esum[j][e].im);
                    }
                }
            }
        }
    }

    if (MODE == 1)
    {
        // reduce over the TY row-lanes that share a vector column
        __shared__ Cx<T> red[kThreads * EPV];
#pragma unroll
        for (int j = 0; j < J; ++j)
        {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                red[threadIdx.x * EPV + e] = esum[j][e];
            __syncthreads();
            for (uint32_t half = TY >> 1; half > 0; half >>= 1)
            {
                if (ty < half)
                {
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                    {
                        Cx<T> o = red[(threadIdx.x + half * TW) * EPV + e];
                        red[threadIdx.x * EPV + e].re += o.re;
                        red[threadIdx.x * EPV + e].im += o.im;
                    }
                }
                __syncthreads();
            }
            if (ty == 0 && vok[j])
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    partials[rb * g.Bpad + vcol[j] * EPV + e] = red[threadIdx.x * EPV + e];
            }
        }
    }
}

// Second-stage reductions run with 32 x 32 threads: x indexes 32 consecutive batch columns (coalesced), y strides
// over the row blocks; partial sums are combined in double precision in a fixed order (deterministic).
constexpr int kFinX = 32, kFinY = 32;

__device__ __forceinline__ double fin_reduce_y(double v, double (*sm)[kFinX + 1])
{
    sm[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    for (int half = kFinY / 2; half > 0; half >>= 1)
    {
        if (static_cast<int>(threadIdx.y) < half)
            sm[threadIdx.y][threadIdx.x] += sm[threadIdx.y + half][threadIdx.x];
        __syncthreads();
    }
    double r = sm[0][threadIdx.x];
    __syncthreads();
    return r;
}

// out[t] = (beta ? out[t] : 0) + sum_rb partials[rb][t]
template <typename T, typename TO = T> // TO: element type of the result (partials may be kept in double for float batches)
__global__ void __launch_bounds__(kFinX *kFinY)
    finalize_complex_kernel(Cx<T> const *__restrict__ partials, uint64_t nRowBlocks, uint32_t Bpad, uint64_t B,
                            Cx<TO> *__restrict__ out, int beta)
{
    __shared__ double sm[kFinY][kFinX + 1];
    uint64_t const t = blockIdx.x * static_cast<uint64_t>(kFinX) + threadIdx.x;
    double re = 0, im = 0;
    if (t < B)
    {
        for (uint64_t rb = threadIdx.y; rb < nRowBlocks; rb += kFinY)
        {
            Cx<T> p = partials[rb * Bpad + t];
            re += p.re;
            im += p.im;
        }
    }
    re = fin_reduce_y(re, sm);
    im = fin_reduce_y(im, sm);
    if (threadIdx.y == 0 && t < B)
    {
        if (beta)
        {
            re += out[t].re;
            im += out[t].im;
        }
        out[t] = Cx<TO>{static_cast<TO>(re), static_cast<TO>(im)};
    }
}

// First stage of a two-stage column sum for long partial lists (thousands of rows: the single-stage finaliser has only
// B / 32 CTAs): CTA (x, s) adds the rows [s * rowsPerSlice, (s + 1) * rowsPerSlice) of its 32 columns in double and
// writes slice row s; finalize_complex_kernel<double, T> then folds the slices.  Fixed order: deterministic.
template <typename T>
__global__ void __launch_bounds__(kFinX *kFinY)
    fold_partials_kernel(Cx<T> const *__restrict__ partials, uint64_t nRowBlocks, uint64_t rowsPerSlice, uint32_t Bpad,
                         uint64_t B, Cx<double> *__restrict__ slices)
{
    __shared__ double sm[kFinY][kFinX + 1];
    uint64_t const t = blockIdx.x * static_cast<uint64_t>(kFinX) + threadIdx.x;
    uint64_t const r0 = blockIdx.y * rowsPerSlice;
    uint64_t const r1 = r0 + rowsPerSlice < nRowBlocks ? r0 + rowsPerSlice : nRowBlocks;
    double re = 0, im = 0;
    if (t < B)
    {
        for (uint64_t rb = r0 + threadIdx.y; rb < r1; rb += kFinY)
        {
            Cx<T> p = partials[rb * Bpad + t];
            re += p.re;
            im += p.im;
        }
    }
    re = fin_reduce_y(re, sm);
    im = fin_reduce_y(im, sm);
    if (threadIdx.y == 0 && t < B)
        slices[static_cast<uint64_t>(blockIdx.y) * Bpad + t] = Cx<double>{re, im};
}

// ---------------------------------------------------------------- K2/K4: paired expectation values
// For strings sharing one x-mask, rows are visited as unordered pairs {i, j = i ^ x} (i has the top bit of x
// clear), so every amplitude is read ONCE (16 B/amp for complex128 instead of 32):
//   conj(psi_i) m_i psi_j + conj(psi_j) m_j psi_i = m_i (q + (-1)^nY conj(q)),  q = conj(psi_i) psi_j,
// i.e. base * sign_i * 2 Re q (nY even) or base * sign_i * 2i Im q (nY odd); for x = 0 it is sign_i |psi_i|^2.
// Only ONE real accumulator per (string, column) is needed; the complex factor is applied by the finaliser
// (PauliString::expectation_value, PS:470-538) or folded into the coefficient matrix (SummedPauliOp, SPO:520-614).
struct PairChunk
{
    uint64_t x;
    uint32_t s0;    // first string of the chunk in sz / sodd
    uint32_t count; // <= MS strings
    uint32_t hbit;  // index of the top set bit of x (unused when x == 0)
    uint32_t diag;  // x == 0
};

template <typename T, int EPV, int V, int MS>
__global__ void __launch_bounds__(kThreads)
    expval_pairs_kernel(PairChunk const *__restrict__ chunks, uint64_t const *__restrict__ sz,
                        uint8_t const *__restrict__ sodd, PairChunk inline_chunk, uint64_t inline_z, uint32_t inline_odd,
                        int use_inline, Geom g, uint64_t dimN, CVec<T, EPV> const *__restrict__ in,
                        T *__restrict__ partials /* [slot][rb][Bpad] */, uint64_t slot_stride)
{
    using Vec = CVec<T, EPV>;
    uint32_t const TW = 1u << g.log2TW;
    uint32_t const TY = kThreads >> g.log2TW;
    uint32_t const lx = threadIdx.x & (TW - 1);
    uint32_t const ty = threadIdx.x >> g.log2TW;
    uint64_t const per_chunk = g.nRowBlocks * g.nColTiles;
    uint64_t const chunk_id = blockIdx.x / per_chunk;
    uint64_t const blk = blockIdx.x - chunk_id * per_chunk;
    uint64_t const ct = blk / g.nRowBlocks;
    uint64_t const rb = blk - ct * g.nRowBlocks;
    uint64_t v = ct * TW + lx;
    bool const vok = v < g.rowvecs;
    if (!vok)
        v = 0;

    PairChunk const ch = use_inline ? inline_chunk : chunks[chunk_id];
    uint64_t zs[MS];
    uint32_t odd_ny[MS];
#pragma unroll
    for (int m = 0; m < MS; ++m)
    {
        bool live = static_cast<uint32_t>(m) < ch.count;
        zs[m] = use_inline ? inline_z : (live ? sz[ch.s0 + m] : 0);
        odd_ny[m] = use_inline ? inline_odd : (live ? sodd[ch.s0 + m] : 0);
    }
    // pair-rows of this chunk: dim/2 when x != 0, dim when x == 0
    uint64_t const nrows = ch.diag ? dimN : (dimN >> 1);
    uint64_t const low_mask = ch.diag ? ~0ull : ((1ull << ch.hbit) - 1);

    T r[MS][EPV];
#pragma unroll
    for (int m = 0; m < MS; ++m)
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            r[m][e] = 0;

    for (uint32_t it = 0; it < g.iters; ++it)
    {
        uint64_t const row0 = (rb * g.iters + it) * (static_cast<uint64_t>(TY) * V) + ty;
        if (row0 - ty >= nrows)
            break; // CTA-uniform: the geometry may be sized for dim rows while this chunk has dim/2 pair-rows
        Vec a[V], b[V];
        uint64_t irow[V];
        bool rok[V];
#pragma unroll
        for (int k = 0; k < V; ++k)
        {
            uint64_t p = row0 + static_cast<uint64_t>(k) * TY;
            rok[k] = p < nrows;
            if (!rok[k])
                p = nrows - 1;
            uint64_t i = ch.diag ? p : (((p & ~low_mask) << 1) | (p & low_mask));
            irow[k] = i;
            a[k] = in[i * g.rowvecs + v];
            if (!ch.diag)
                b[k] = in[(i ^ ch.x) * g.rowvecs + v];
        }
#pragma unroll
        for (int k = 0; k < V; ++k)
        {
            if (!rok[k])
                continue;
            T qre[EPV], qim[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                if (ch.diag)
                {
                    qre[e] = fma(a[k].e[e].re, a[k].e[e].re, a[k].e[e].im * a[k].e[e].im);
                    qim[e] = 0;
                }
                else
                {
                    qre[e] = fma(a[k].e[e].re, b[k].e[e].re, a[k].e[e].im * b[k].e[e].im);
                    qim[e] = fma(a[k].e[e].re, b[k].e[e].im, -a[k].e[e].im * b[k].e[e].re);
                }
            }
#pragma unroll
            for (int m = 0; m < MS; ++m)
            {
                uint32_t sgn = parity64(irow[k] & zs[m]);
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    r[m][e] += flip_sign(odd_ny[m] ? qim[e] : qre[e], sgn);
            }
        }
    }

    __shared__ T red[kThreads * EPV];
#pragma unroll
    for (int m = 0; m < MS; ++m)
    {
        if (static_cast<uint32_t>(m) >= ch.count)
            break; // uniform across the CTA
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            red[threadIdx.x * EPV + e] = r[m][e];
        __syncthreads();
        for (uint32_t half = TY >> 1; half > 0; half >>= 1)
        {
            if (ty < half)
            {
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    red[threadIdx.x * EPV + e] += red[(threadIdx.x + half * TW) * EPV + e];
            }
            __syncthreads();
        }
        if (ty == 0 && vok)
        {
            uint64_t slot = use_inline ? 0 : (ch.s0 + m);
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                partials[slot * slot_stride + rb * g.Bpad + v * EPV + e] = red[threadIdx.x * EPV + e];
        }
    }
}

// Single string: out[t] = (beta ? out[t] : 0) + factor * sum_rb partials[rb][t]
template <typename T>
__global__ void __launch_bounds__(kFinX *kFinY)
    finalize_pairs_string_kernel(T const *__restrict__ partials, uint64_t nRowBlocks, uint32_t Bpad, uint64_t B,
                                 double fre, double fim, Cx<T> *__restrict__ out, int beta)
{
    __shared__ double sm[kFinY][kFinX + 1];
    uint64_t const t = blockIdx.x * static_cast<uint64_t>(kFinX) + threadIdx.x;
    double r = 0;
    if (t < B)
        for (uint64_t rb = threadIdx.y; rb < nRowBlocks; rb += kFinY)
            r += partials[rb * Bpad + t];
    r = fin_reduce_y(r, sm);
    if (threadIdx.y == 0 && t < B)
    {
        double re = fre * r, im = fim * r;
        if (beta)
        {
            re += out[t].re;
            im += out[t].im;
        }
        out[t] = Cx<T>{static_cast<T>(re), static_cast<T>(im)};
    }
}

// SummedPauliOp: E(s,t) = sum_rb partials[s][rb][t]  -> dense real (S, B) matrix feeding the contraction
template <typename T>
__global__ void __launch_bounds__(kFinX *kFinY)
    finalize_pairs_matrix_kernel(T const *__restrict__ partials, uint64_t slot_stride, uint64_t nRowBlocks,
                                 uint32_t Bpad, uint64_t B, T *__restrict__ E /* [S][B] */)
{
    __shared__ double sm[kFinY][kFinX + 1];
    uint64_t const t = blockIdx.x * static_cast<uint64_t>(kFinX) + threadIdx.x;
    uint64_t const s = blockIdx.y;
    double r = 0;
    if (t < B)
        for (uint64_t rb = threadIdx.y; rb < nRowBlocks; rb += kFinY)
            r += partials[s * slot_stride + rb * Bpad + t];
    r = fin_reduce_y(r, sm);
    if (threadIdx.y == 0 && t < B)
        E[s * B + t] = static_cast<T>(r);
}

// ---------------------------------------------------------------- K6: weighted apply (SummedPauliOp::apply_weighted)
//   out(l,t) (+)= sum_g [ sum_{s in g} (-1)^popc(l & z_s) W(s,t) ] psi(l ^ x_g, t)          (SPO:441-455)
// W is the planar contraction result: Wre/Wim are (S_packed, B) real matrices whose rows already contain
// (-i)^nY_s sum_k coeffs(s,k) data(k,t).  The thread's W vector is loaded once per string and reused for its V rows.
template <typename T, int EPV, int V>
__global__ void __launch_bounds__(kThreads)
    weighted_apply_kernel(OpView<T> op, Geom g, T const *__restrict__ Wre, T const *__restrict__ Wim, uint64_t B,
                          CVec<T, EPV> const *__restrict__ in, CVec<T, EPV> *__restrict__ out, int beta)
{
    using Vec = CVec<T, EPV>;
    uint32_t const TW = 1u << g.log2TW;
    uint32_t const TY = kThreads >> g.log2TW;
    uint32_t const lx = threadIdx.x & (TW - 1);
    uint32_t const ty = threadIdx.x >> g.log2TW;
    uint64_t const blk = blockIdx.x;
    uint64_t const ct = blk / g.nRowBlocks;
    uint64_t const rb = blk - ct * g.nRowBlocks;
    uint64_t const v = ct * TW + lx;
    if (v >= g.rowvecs)
        return;
    uint64_t const row0 = rb * (static_cast<uint64_t>(TY) * V) + ty;
    uint64_t rows[V];
    bool rok[V];
#pragma unroll
    for (int k = 0; k < V; ++k)
    {
        uint64_t r = row0 + static_cast<uint64_t>(k) * TY;
        rok[k] = r < g.N;
        rows[k] = rok[k] ? r : (g.N - 1);
    }
    Cx<T> acc[V][EPV];
#pragma unroll
    for (int k = 0; k < V; ++k)
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            acc[k][e] = Cx<T>{0, 0};

    uint64_t const t0 = v * EPV;
    uint32_t s0 = __ldg(op.gstart);
    for (uint32_t gi = 0; gi < op.G; ++gi)
    {
        uint64_t const x = __ldg(op.gx + gi);
        uint32_t const s1 = __ldg(op.gstart + gi + 1);
        Vec src[V];
#pragma unroll
        for (int k = 0; k < V; ++k)
            src[k] = in[(rows[k] ^ x) * g.rowvecs + v];
        Cx<T> d[V][EPV];
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                d[k][e] = Cx<T>{0, 0};
        for (uint32_t s = s0; s < s1; ++s)
        {
            uint64_t const z = __ldg(op.sz + s);
            T wre[EPV], wim[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                wre[e] = Wre[static_cast<uint64_t>(s) * B + t0 + e];
                wim[e] = Wim[static_cast<uint64_t>(s) * B + t0 + e];
            }
#pragma unroll
            for (int k = 0; k < V; ++k)
            {
                uint32_t odd = parity64(rows[k] & z);
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                {
                    d[k][e].re += flip_sign(wre[e], odd);
                    d[k][e].im += flip_sign(wim[e], odd);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                cfma(acc[k][e], d[k][e], src[k].e[e]);
        s0 = s1;
    }
#pragma unroll
    for (int k = 0; k < V; ++k)
    {
        if (!rok[k])
            continue;
        uint64_t const o = rows[k] * g.rowvecs + v;
        Vec r;
        if (beta)
        {
            r = out[o];
#pragma unroll
            for (int e = 0; e < EPV; ++e)
            {
                r.e[e].re += acc[k][e].re;
                r.e[e].im += acc[k][e].im;
            }
        }
        else
        {
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                r.e[e] = acc[k][e];
        }
        out[o] = r;
    }
}

// ---------------------------------------------------------------- K5 (SIMT engine): real GEMM  C = A * B
// C[M x N] = A[M x Kd] * Bm[Kd x N], all row-major.  A is a planar-stacked coefficient matrix prepared at plan
// creation ([Re; Im] rows), Bm is the real data / expectation matrix (DT = float or double, converted to T).
// splitK > 1 writes partial products to C + z * M * N (reduced by the caller's finaliser).
// Used for complex128 plans and as the verification / fallback engine for the tcgen05 path.
template <typename T, typename DT>
__global__ void __launch_bounds__(256) gemm_simt_kernel(T const *__restrict__ A, DT const *__restrict__ Bm,
                                                         T *__restrict__ C, uint32_t M, uint64_t N, uint32_t Kd,
                                                         uint32_t kchunk)
{
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN];
    uint32_t const tx = threadIdx.x & 15, tyy = threadIdx.x >> 4;
    uint64_t const n0 = static_cast<uint64_t>(blockIdx.x) * BN;
    uint32_t const m0 = blockIdx.y * BM;
    uint32_t const kbeg = blockIdx.z * kchunk;
    uint32_t const kend = min(Kd, kbeg + kchunk);
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = 0;
    for (uint32_t k0 = kbeg; k0 < kend; k0 += BK)
    {
        // A tile: BM x BK (coalesced along k)
        for (uint32_t idx = threadIdx.x; idx < BM * BK; idx += 256)
        {
            uint32_t m = idx / BK, k = idx % BK;
            T val = 0;
            if (m0 + m < M && k0 + k < kend)
                val = A[static_cast<uint64_t>(m0 + m) * Kd + k0 + k];
            As[k][m] = val;
        }
        for (uint32_t idx = threadIdx.x; idx < BK * BN; idx += 256)
        {
            uint32_t k = idx / BN, n = idx % BN;
            T val = 0;
            if (k0 + k < kend && n0 + n < N)
                val = static_cast<T>(Bm[static_cast<uint64_t>(k0 + k) * N + n0 + n]);
            Bs[k][n] = val;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k)
        {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = As[k][tyy * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    T *Cz = C + static_cast<uint64_t>(blockIdx.z) * M * N;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        uint32_t m = m0 + tyy * 4 + i;
        if (m >= M)
            continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            uint64_t n = n0 + tx + 16 * j;
            if (n < N)
                Cz[static_cast<uint64_t>(m) * N + n] = acc[i][j];
        }
    }
}

// out(k,t) = (beta ? out : 0) + sum_z (Cre[z][k][t] + i Cim[z][k][t]); C is the planar [2K x B] split-K stack
template <typename T>
__global__ void finalize_sop_expval_kernel(T const *__restrict__ C, uint32_t splitK, uint32_t K, uint64_t B,
                                           Cx<T> *__restrict__ out, int beta)
{
    uint64_t idx = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<uint64_t>(K) * B)
        return;
    uint64_t const plane = 2ull * K * B;
    double re = 0, im = 0;
    for (uint32_t z = 0; z < splitK; ++z)
    {
        re += C[z * plane + idx];
        im += C[z * plane + static_cast<uint64_t>(K) * B + idx];
    }
    if (beta)
    {
        re += out[idx].re;
        im += out[idx].im;
    }
    out[idx] = Cx<T>{static_cast<T>(re), static_cast<T>(im)};
}

// ---------------------------------------------------------------- synthetic input generator (bench / tests)
__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline double u01_from_bits(uint64_t h)
{
    return static_cast<double>(h >> 11) * (1.0 / 9007199254740992.0);
}

template <typename T>
__global__ void fill_uniform_kernel(T *__restrict__ dst, uint64_t n_real, uint64_t first_real, uint64_t seed)
{
    uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; e < n_real; e += stride)
        dst[e] = static_cast<T>(u01_from_bits(splitmix64(seed * 0xD1342543DE82EF95ull + first_real + e)));
}

} // namespace fpk
