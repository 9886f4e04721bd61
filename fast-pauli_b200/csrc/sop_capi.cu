// C ABI, second translation unit: SummedPauliOp (apply / apply_weighted / expectation_value / square), the tensor-core
// contraction engine, the process-wide default context and the one-shot entry points.  Shares capi_internal.hpp and
// coset_launch.hpp with capi.cu; calls the PauliOp path only through the public C ABI.
#include "capi_internal.hpp"
#include "coset_launch.hpp"
#include "etile.cuh"
#include "gemm_tc.cuh"
#include "square.cuh"
#include "wtile.cuh"

namespace
{
// C[M x N] = A[M x Kd] * Bm[Kd x N]  (+ split-K planes)
template <typename T, typename DT>
int run_gemm(fp_ctx *ctx, T const *A, DT const *Bm, T *C, uint32_t M, uint64_t N, uint32_t Kd, uint32_t splitK,
             uint32_t kchunk)
{
    if (M == 0 || N == 0)
        return FP_OK;
    if constexpr (std::is_same<T, float>::value && std::is_same<DT, float>::value)
    {
        if (ctx->tensor_core && gemm_tc_supported(M, N, Kd, splitK))
        {
            int rc = gemm_tc_3xtf32(ctx->stream, A, Bm, C, M, N, Kd, splitK, kchunk);
            if (rc == 0)
            {
                ctx->launches++;
                ctx->last_gemm_engine = 1;
                return FP_OK;
            }
        }
    }
    dim3 grid(static_cast<unsigned>((N + 63) / 64), (M + 63) / 64, splitK);
    if (grid.y > 65535 || grid.z > 65535)
        return set_err(FP_UNSUPPORTED, "contraction too large");
    gemm_simt_kernel<T, DT><<<grid, 256, 0, ctx->stream>>>(A, Bm, C, M, N, Kd, kchunk);
    ctx->launches++;
    ctx->last_gemm_engine = 0;
    return FP_OK;
}

} // namespace

// ================================================================ SummedPauliOp
namespace
{
template <typename T>
int sop_create_t(fp_ctx *ctx, int dtype, int n, size_t S, uint8_t const *codes, size_t K,
                 std::complex<T> const *coeffs, fp_sop **out)
{
    std::unique_ptr<fp_sop> sop(new fp_sop);
    sop->dtype = dtype;
    sop->device = ctx->device;
    sop->n_qubits = n;
    sop->n_strings = S;
    sop->n_ops = K;
    // (1) SummedPauliOp::apply: c_j = sum_k coeffs(j,k), summed in k order in T like SPO:312-317 / 341-345
    std::vector<std::complex<T>> csum(S);
    for (size_t j = 0; j < S; ++j)
    {
        std::complex<T> c(0, 0);
        for (size_t k = 0; k < K; ++k)
            c += coeffs[j * K + k];
        csum[j] = c;
    }
    FP_TRY(op_create_t<T>(ctx, dtype, n, S, codes, csum.data(), true, &sop->summed));
    // (2) unmerged packed strings with unit coefficients: masks + order for the W / E matrices
    std::vector<std::complex<T>> ones(S, std::complex<T>(1, 0));
    int rc = op_create_t<T>(ctx, dtype, n, S, codes, ones.data(), false, &sop->strings);
    if (rc != FP_OK)
    {
        fp_op_destroy(sop->summed);
        return rc;
    }
    PackedOp<T> const &pk = dop<T>(sop->strings).host;
    // (3) planar coefficient matrices in packed order
    std::vector<T> Aw(2 * S * K), Ae(2 * K * S);
    for (size_t p = 0; p < S; ++p)
    {
        size_t j = pk.perm[p];
        uint32_t ny = pk.sny[p];
        bool diag = false;
        // group lookup is not needed: x == 0 iff the string has no X/Y; recompute from the masks
        {
            StringMasks mk = make_masks(n, codes + j * static_cast<size_t>(n));
            diag = mk.x == 0;
        }
        for (size_t k = 0; k < K; ++k)
        {
            std::complex<T> c = times_phase(coeffs[j * K + k], ny); // coeffs(j,k) * (-i)^nY
            Aw[p * K + k] = c.real();
            Aw[(S + p) * K + k] = c.imag();
            // pair factor of the expectation kernel: 1 (x == 0), 2 (nY even), 2i (nY odd)
            std::complex<T> e = diag ? c : ((ny & 1u) ? std::complex<T>(-2 * c.imag(), 2 * c.real()) : T(2) * c);
            Ae[k * S + p] = e.real();
            Ae[(K + k) * S + p] = e.imag();
        }
    }
    T *dAw = nullptr, *dAe = nullptr;
    rc = upload_vec(&dAw, Aw);
    if (rc == FP_OK)
        rc = upload_vec(&dAe, Ae);
    if (rc != FP_OK)
    {
        cudaFree(dAw);
        fp_op_destroy(sop->summed);
        fp_op_destroy(sop->strings);
        return rc;
    }
    sop->A_w = dAw;
    sop->A_e = dAe;
    *out = sop.release();
    return FP_OK;
}

int sop_check(fp_ctx *ctx, fp_sop const *sop)
{
    if (!ctx || !sop)
        return set_err(FP_INVALID_ARGUMENT, "null context or operator");
    if (sop->device != ctx->device)
        return set_err(FP_INVALID_ARGUMENT, "operator plan was created on a different device than the context");
    return FP_OK;
}

// K6b (wtile.cuh): whole state column (pair) in shared memory; complex64 (packed FP32) or complex128, 11-12 qubits
template <typename T, int LOG_NT>
int launch_wtile(fp_ctx *ctx, CosetPassView<T> const &view, uint64_t rowvecs, uint64_t grid, void const *in, void *out,
                 int beta, T const *Wre, T const *Wim, uint64_t B)
{
    constexpr size_t smem = WtileSmem<LOG_NT>::bytes;
    static PerDevice configured;
    if constexpr (sizeof(T) == 4)
    {
        if (!configured.done(ctx->device))
        {
            FP_CU(cudaFuncSetAttribute(wtile_kernel<LOG_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
            configured.set(ctx->device);
        }
        wtile_kernel<LOG_NT><<<static_cast<unsigned>(grid), 1 << LOG_NT, smem, ctx->stream>>>(
            view, rowvecs, static_cast<CVec<float, 2> const *>(in), static_cast<CVec<float, 2> *>(out), beta, Wre, Wim, B);
    }
    else
    {
        if (!configured.done(ctx->device))
        {
            FP_CU(cudaFuncSetAttribute(wtile_f64_kernel<LOG_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
            configured.set(ctx->device);
        }
        wtile_f64_kernel<LOG_NT><<<static_cast<unsigned>(grid), 1 << LOG_NT, smem, ctx->stream>>>(
            view, rowvecs, static_cast<CVec<double, 1> const *>(in), static_cast<CVec<double, 1> *>(out), beta, Wre, Wim,
            B);
    }
    ctx->launches++;
    return FP_OK;
}

template <typename T>
int try_wtile(fp_ctx *ctx, DeviceOp<T> const &op, int n_qubits, void *out, void const *in, uint64_t dim, uint64_t B,
              int beta, T const *Wre, T const *Wim, bool *used)
{
    *used = false;
    constexpr int EPV = sizeof(T) == 4 ? 2 : 1;
    if (!ctx->wtile || ctx->coset_mode != 1 || ctx->coset_log_twc >= 0 || n_qubits < 11 ||
        dim != (1ull << n_qubits) || pick_epv<T>(in, out, B) != EPV)
        return FP_OK;
    if (n_qubits > 12 && op.host.gx.size() > 20000)
        return FP_OK; // pass planning is quadratic in the number of x-groups
    int const rank = std::min(n_qubits, 12);
    std::vector<typename DeviceOp<T>::CosetPassDev> const *passes = nullptr;
    FP_TRY(get_coset_plan<T>(op, n_qubits, rank, 2, &passes));
    // every pass re-streams the batch: only worth it while passes << groups
    if (passes->size() > 1 && passes->size() * 3 > op.host.gx.size())
        return FP_OK;
    uint64_t const rowvecs = B / EPV;
    uint64_t const grid = (1ull << (n_qubits - rank)) * rowvecs;
    FP_TRY(check_grid(grid));
    for (size_t p = 0; p < passes->size(); ++p)
    {
        CosetPassView<T> const &view = (*passes)[p].view;
        int const b = p == 0 ? beta : 1;
        if (rank == 12)
            FP_TRY((launch_wtile<T, 9>(ctx, view, rowvecs, grid, in, out, b, Wre, Wim, B)));
        else
            FP_TRY((launch_wtile<T, 8>(ctx, view, rowvecs, grid, in, out, b, Wre, Wim, B)));
    }
    *used = true;
    return FP_OK;
}

template <typename T, typename DT>
int run_sop_weighted(fp_ctx *ctx, fp_sop const *sop, void *out, void const *in, DT const *data, uint64_t dim, uint64_t B,
                     int beta)
{
    uint32_t const S = static_cast<uint32_t>(sop->n_strings), K = static_cast<uint32_t>(sop->n_ops);
    DeviceOp<T> const &op = dop<T>(sop->strings);
    // W[2S x B] = A_w[2S x K] * data[K x B]                                   (SPO:413-432, the tensor-core step)
    FP_TRY(ctx->work_a.ensure(2ull * S * B * sizeof(T)));
    T *W = static_cast<T *>(ctx->work_a.p);
    FP_TRY((run_gemm<T, DT>(ctx, static_cast<T const *>(sop->A_w), data, W, 2 * S, B, K, 1, K ? K : 1)));
    FP_TRY(check_align(in, 2 * sizeof(T), "states"));
    FP_TRY(check_align(out, 2 * sizeof(T), "new_states"));
    T const *Wre = W, *Wim = W + static_cast<uint64_t>(S) * B;
    if (op.host.gx.size() > 1)
    {
        bool used = false;
        FP_TRY((try_wtile<T>(ctx, op, sop->n_qubits, out, in, dim, B, beta, Wre, Wim, &used)));
        if (used)
            return FP_OK;
        FP_TRY((try_coset<T, 2>(ctx, op, sop->n_qubits, out, in, dim, B, beta, Wre, Wim, &used)));
        if (used)
            return FP_OK;
    }
    int const epv = pick_epv<T>(in, out, B);
    uint64_t const rowvecs = B / epv;
    GeomSel gs = choose_geom(ctx, dim, dim, rowvecs, 2 * sizeof(T) * epv, op.host.gx.size() > 1, false);
    FP_TRY(check_grid(gs.grid));
    OpView<T> view = op.view();
    dim3 grid(static_cast<unsigned>(gs.grid));
    bool done = false;
    if constexpr (sizeof(T) == 4)
    {
        if (epv == 2)
        {
            auto const *din = static_cast<CVec<T, 2> const *>(in);
            auto *dout = static_cast<CVec<T, 2> *>(out);
            if (gs.V == 4)
                weighted_apply_kernel<T, 2, 4>
                    <<<grid, kThreads, 0, ctx->stream>>>(view, gs.g, Wre, Wim, B, din, dout, beta);
            else
                weighted_apply_kernel<T, 2, 1>
                    <<<grid, kThreads, 0, ctx->stream>>>(view, gs.g, Wre, Wim, B, din, dout, beta);
            done = true;
        }
    }
    if (!done)
    {
        auto const *din = static_cast<CVec<T, 1> const *>(in);
        auto *dout = static_cast<CVec<T, 1> *>(out);
        if (gs.V == 4)
            weighted_apply_kernel<T, 1, 4><<<grid, kThreads, 0, ctx->stream>>>(view, gs.g, Wre, Wim, B, din, dout, beta);
        else
            weighted_apply_kernel<T, 1, 1><<<grid, kThreads, 0, ctx->stream>>>(view, gs.g, Wre, Wim, B, din, dout, beta);
    }
    ctx->launches++;
    return FP_OK;
}

template <typename T>
int run_sop_expval(fp_ctx *ctx, fp_sop const *sop, void *out, void const *in, uint64_t dim, uint64_t B, int beta)
{
    uint32_t const S = static_cast<uint32_t>(sop->n_strings), K = static_cast<uint32_t>(sop->n_ops);
    DeviceOp<T> const &op = dop<T>(sop->strings);
    FP_TRY(check_align(in, 2 * sizeof(T), "states"));
    int const epv = pick_epv<T>(in, in, B);
    uint64_t const rowvecs = B / epv;
    // stage 1: E(s,t) per packed string                                      (SPO:573-577)
    FP_TRY(ctx->work_a.ensure(static_cast<uint64_t>(S) * B * sizeof(T)));
    T *E = static_cast<T *>(ctx->work_a.p);
    bool stage1_done = false;
    constexpr int EPV_FULL = sizeof(T) == 4 ? 2 : 1;
    if (ctx->coset_mode != 0 && sop->n_qubits >= 5 && sop->n_qubits <= 12 && epv == EPV_FULL)
    {
        // K4b: the whole state column lives in shared memory and every string is evaluated against it
        uint32_t splits = 1;
        while (rowvecs * splits < static_cast<uint64_t>(ctx->sm_count) * 2 && splits * 8 < op.n_chunks)
            splits *= 2;
        size_t const smem = (static_cast<size_t>(1) << sop->n_qubits) * 16;
        static PerDevice configured;
        if (!configured.done(ctx->device))
        {
            FP_CU(cudaFuncSetAttribute(sop_expval_tile_kernel<T, EPV_FULL, kPairMS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            configured.set(ctx->device);
        }
        bool launched = false;
        if (ctx->etile && sop->n_qubits >= 9 && rowvecs <= 0x7fffffffull)
        {
            // K4c (etile.cuh): planar pair tile / packed FP32 (complex64) or FP64 (complex128), compile-time sign patterns
            using P = typename std::conditional<sizeof(T) == 4, EtF32, EtF64>::type;
            static PerDevice configured2;
            if (!configured2.done(ctx->device))
            {
                FP_CU(cudaFuncSetAttribute(sop_expval_tile2_kernel<P, kPairMS, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
                configured2.set(ctx->device);
            }
            EtStrings st{};
            st.chunks = op.chunks;
            st.n_chunks = op.n_chunks;
            st.sz = op.sz;
            st.sodd = op.sodd;
            st.n_cosets = 1;
            dim3 grid(static_cast<unsigned>(rowvecs), splits);
            sop_expval_tile2_kernel<P, kPairMS, false><<<grid, kThreads, smem, ctx->stream>>>(
                st, static_cast<uint32_t>(sop->n_qubits), rowvecs, static_cast<CVec<T, EPV_FULL> const *>(in), E, B);
            ctx->launches++;
            stage1_done = launched = true;
        }
        if (!launched && rowvecs <= 0x7fffffffull)
        {
            dim3 grid(static_cast<unsigned>(rowvecs), splits);
            sop_expval_tile_kernel<T, EPV_FULL, kPairMS><<<grid, kThreads, smem, ctx->stream>>>(
                op.chunks, op.n_chunks, op.sz, op.sodd, static_cast<uint32_t>(sop->n_qubits), rowvecs,
                static_cast<CVec<T, EPV_FULL> const *>(in), E, B);
            ctx->launches++;
            stage1_done = true;
        }
    }
    if (!stage1_done && ctx->etile && ctx->coset_mode == 1 && sop->n_qubits > 12 && epv == EPV_FULL &&
        dim == (1ull << sop->n_qubits) && rowvecs <= 0x7fffffffull && op.host.gx.size() <= 20000)
    {
        // K4c over rank-12 coset tiles: every CTA walks all cosets of a pass for its column (pair)
        using P = typename std::conditional<sizeof(T) == 4, EtF32, EtF64>::type;
        std::vector<typename DeviceOp<T>::CosetPassDev> const *passes = nullptr;
        FP_TRY(get_coset_plan<T>(op, sop->n_qubits, 12, 2, &passes));
        if (passes->size() == 1 || passes->size() * 3 <= op.host.gx.size())
        {
            static PerDevice configured3;
            if (!configured3.done(ctx->device))
            {
                FP_CU(cudaFuncSetAttribute(sop_expval_tile2_kernel<P, kPairMS, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
                configured3.set(ctx->device);
            }
            for (auto const &pd : *passes)
            {
                EtStrings st{};
                st.chunks = pd.echunks;
                st.n_chunks = pd.n_echunks;
                st.sz = pd.view.sz;
                st.szl = pd.view.szl;
                st.sodd = pd.esodd;
                st.sidx = pd.view.sidx;
                for (int k = 0; k < 12; ++k)
                    st.basis[k] = pd.view.basis[k];
                st.nonpivot_mask = pd.view.nonpivot_mask;
                st.n_cosets = 1ull << (sop->n_qubits - 12);
                uint32_t splits = 1;
                while (rowvecs * splits < static_cast<uint64_t>(ctx->sm_count) * 2 && splits * 8 < st.n_chunks)
                    splits *= 2;
                dim3 grid(static_cast<unsigned>(rowvecs), splits);
                sop_expval_tile2_kernel<P, kPairMS, true><<<grid, kThreads, 65536, ctx->stream>>>(
                    st, static_cast<uint32_t>(sop->n_qubits), rowvecs, static_cast<CVec<T, EPV_FULL> const *>(in), E, B);
                ctx->launches++;
            }
            stage1_done = true;
        }
    }
    uint64_t const rows = op.any_diag ? dim : dim / 2;
    GeomSel gs = choose_geom(ctx, rows, dim, rowvecs, 2 * sizeof(T) * epv, false, true, op.n_chunks);
    FP_TRY(check_grid(gs.grid));
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    gs.g.Bpad = Bpad;
    uint64_t const slot_stride = gs.g.nRowBlocks * Bpad;
    // generic stage 1 (any register size): paired kernel with per-row-block partial sums
    T *part = E;
    if (!stage1_done)
    {
    if (gs.g.nRowBlocks > 1 || Bpad != B)
    {
        FP_TRY(ctx->partials.ensure(static_cast<uint64_t>(S) * slot_stride * sizeof(T)));
        part = static_cast<T *>(ctx->partials.p);
    }
    PairChunk none{};
    bool done = false;
    if constexpr (sizeof(T) == 4)
    {
        if (epv == 2)
        {
            launch_pairs_v<T, 2, kPairMS>(ctx, gs, op.chunks, op.sz, op.sodd, none, 0, 0, 0, dim, in, part, slot_stride);
            done = true;
        }
    }
    if (!done)
        launch_pairs_v<T, 1, kPairMS>(ctx, gs, op.chunks, op.sz, op.sodd, none, 0, 0, 0, dim, in, part, slot_stride);
    if (part != E)
    {
        dim3 fgrid(static_cast<unsigned>((B + kFinX - 1) / kFinX), S);
        if (S > 65535)
            return set_err(FP_UNSUPPORTED, "too many strings for the expectation finaliser");
        finalize_pairs_matrix_kernel<T>
            <<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(part, slot_stride, gs.g.nRowBlocks, Bpad, B, E);
        ctx->launches++;
    }
    }
    // stage 2: out[2K x B] = A_e[2K x S] * E[S x B], split over S              (SPO:579-591)
    uint32_t kchunk = 512;
    uint32_t splitK = std::max<uint32_t>(1, (S + kchunk - 1) / kchunk);
    FP_TRY(ctx->work_b.ensure(static_cast<uint64_t>(splitK) * 2 * K * B * sizeof(T)));
    T *Cst = static_cast<T *>(ctx->work_b.p);
    FP_TRY((run_gemm<T, T>(ctx, static_cast<T const *>(sop->A_e), E, Cst, 2 * K, B, S, splitK, kchunk)));
    uint64_t total = static_cast<uint64_t>(K) * B;
    finalize_sop_expval_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(
        Cst, splitK, K, B, static_cast<Cx<T> *>(out), beta);
    ctx->launches++;
    return FP_OK;
}
} // namespace

extern "C"
{

    int fp_sop_create(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes, size_t n_operators,
                      const void *coeffs, fp_sop **sop)
    {
        if (!ctx || !sop || !coeffs || (n_qubits > 0 && !codes))
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        if (n_strings == 0) // the reference dereferences pauli_strings.front() (SPO:50): undefined; reject instead
            return set_err(FP_INVALID_ARGUMENT, "SummedPauliOp needs at least one PauliString");
        if (n_strings > 0x7fffffffull || n_operators > 0x3fffffffull)
            return set_err(FP_UNSUPPORTED, "operator too large");
        DeviceGuard g(ctx->device);
        if (dtype == FP_C128)
            return sop_create_t<double>(ctx, dtype, n_qubits, n_strings, codes, n_operators,
                                        static_cast<std::complex<double> const *>(coeffs), sop);
        return sop_create_t<float>(ctx, dtype, n_qubits, n_strings, codes, n_operators,
                                   static_cast<std::complex<float> const *>(coeffs), sop);
    }

    int fp_sop_destroy(fp_sop *sop)
    {
        if (!sop)
            return FP_OK;
        DeviceGuard g(sop->device);
        fp_op_destroy(sop->summed);
        fp_op_destroy(sop->strings);
        cudaFree(sop->A_w);
        cudaFree(sop->A_e);
        delete sop;
        return FP_OK;
    }

    int fp_sop_apply(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, size_t dim, size_t n_states,
                     int accumulate)
    {
        FP_TRY(sop_check(ctx, sop));
        if (dim != dim_of(sop->n_qubits))
            return set_err(FP_INVALID_ARGUMENT, "state size must match the dimension of the operators");
        return fp_op_apply(ctx, sop->summed, out, in, dim, n_states, accumulate);
    }

    int fp_sop_apply_weighted(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, const void *data,
                              int data_is_f64, size_t dim, size_t n_states, int accumulate)
    {
        FP_TRY(sop_check(ctx, sop));
        if (dim != dim_of(sop->n_qubits)) // SPO:396-399
            return set_err(FP_INVALID_ARGUMENT, "state size must match the dimension of the operators");
        if (dim == 0 || n_states == 0)
            return FP_OK;
        if (!data && sop->n_ops)
            return set_err(FP_INVALID_ARGUMENT, "null data pointer");
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        size_t const bytes = dim * n_states * csize(sop->dtype);
        Staged sin, sout, sdat;
        FP_TRY(stage_in(ctx, ctx->stage_in, in, bytes, true, sin));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, bytes, accumulate != 0, sout));
        FP_TRY(stage_in(ctx, ctx->stage_data, data, sop->n_ops * n_states * (data_is_f64 ? 8 : 4), true, sdat));
        int rc;
        if (sop->dtype == FP_C128)
            rc = data_is_f64 ? run_sop_weighted<double, double>(ctx, sop, sout.dev, sin.dev,
                                                                static_cast<double const *>(sdat.dev), dim, n_states,
                                                                accumulate)
                             : run_sop_weighted<double, float>(ctx, sop, sout.dev, sin.dev,
                                                               static_cast<float const *>(sdat.dev), dim, n_states,
                                                               accumulate);
        else
            rc = data_is_f64 ? run_sop_weighted<float, double>(ctx, sop, sout.dev, sin.dev,
                                                               static_cast<double const *>(sdat.dev), dim, n_states,
                                                               accumulate)
                             : run_sop_weighted<float, float>(ctx, sop, sout.dev, sin.dev,
                                                              static_cast<float const *>(sdat.dev), dim, n_states,
                                                              accumulate);
        FP_TRY(rc);
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged || sdat.staged);
    }

    int fp_sop_expval(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, size_t dim, size_t n_states,
                      int accumulate)
    {
        FP_TRY(sop_check(ctx, sop));
        if (dim != dim_of(sop->n_qubits)) // SPO:546-551
            return set_err(FP_INVALID_ARGUMENT, "states must have the same dimension (" + std::to_string(dim) +
                                                    ") as the SummedPauliOp (" +
                                                    std::to_string(dim_of(sop->n_qubits)) + ")");
        if (n_states == 0 || sop->n_ops == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        Staged sin, sout;
        FP_TRY(stage_in(ctx, ctx->stage_in, in, dim * n_states * csize(sop->dtype), true, sin));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, sop->n_ops * n_states * csize(sop->dtype), accumulate != 0, sout));
        if (sop->dtype == FP_C128)
            FP_TRY(run_sop_expval<double>(ctx, sop, sout.dev, sin.dev, dim, n_states, accumulate));
        else
            FP_TRY(run_sop_expval<float>(ctx, sop, sout.dev, sin.dev, dim, n_states, accumulate));
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged);
    }

    // ------------------------------------------------------------ peer memory (one process per GPU, NVLink P2P)
    int fp_ipc_export(fp_ctx *ctx, const void *dev_ptr, unsigned char *handle)
    {
        if (!ctx || !dev_ptr || !handle)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        static_assert(sizeof(cudaIpcMemHandle_t) == FP_IPC_HANDLE_BYTES, "IPC handle size");
        DeviceGuard g(ctx->device);
        cudaIpcMemHandle_t h;
        FP_CU(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
        memcpy(handle, &h, sizeof h);
        return FP_OK;
    }

    int fp_ipc_open(fp_ctx *ctx, const unsigned char *handle, void **peer_ptr)
    {
        if (!ctx || !handle || !peer_ptr)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof h);
        *peer_ptr = nullptr;
        FP_CU(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
        return FP_OK;
    }

    int fp_ipc_close(fp_ctx *ctx, void *peer_ptr)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        if (peer_ptr)
            FP_CU(cudaIpcCloseMemHandle(peer_ptr));
        return FP_OK;
    }

    // ------------------------------------------------------------ diagnostics
    int fp_debug_gemm_f32(fp_ctx *ctx, int engine, const float *A, const float *B, float *C, uint32_t M, uint64_t N,
                          uint32_t Kd, uint32_t split_k)
    {
        if (!ctx || !A || !B || !C || M == 0 || N == 0 || Kd == 0 || split_k == 0)
            return set_err(FP_INVALID_ARGUMENT, "bad gemm arguments");
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        Staged sa, sb, sc;
        FP_TRY(stage_in(ctx, ctx->stage_in, A, static_cast<size_t>(M) * Kd * 4, true, sa));
        FP_TRY(stage_in(ctx, ctx->stage_data, B, static_cast<size_t>(Kd) * N * 4, true, sb));
        FP_TRY(stage_in(ctx, ctx->stage_out, C, static_cast<size_t>(split_k) * M * N * 4, false, sc));
        uint32_t kchunk = (Kd + split_k - 1) / split_k;
        if (split_k > 1)
            kchunk = (kchunk + 31) / 32 * 32;
        bool const saved = ctx->tensor_core;
        ctx->tensor_core = engine == 1;
        uint64_t const before = ctx->launches;
        int rc = run_gemm<float, float>(ctx, static_cast<float const *>(sa.dev), static_cast<float const *>(sb.dev),
                                        static_cast<float *>(sc.dev), M, N, Kd, split_k, kchunk);
        ctx->tensor_core = saved;
        FP_TRY(rc);
        (void)before;
        FP_TRY(stage_back(ctx, sc));
        return finish(ctx, true);
    }

    // ------------------------------------------------------------ one-shot entry points (oracle-shaped)
    int fp_default_ctx(fp_ctx **out)
    {
        static std::mutex mu;
        static fp_ctx *ctx = nullptr;
        std::lock_guard<std::mutex> lk(mu);
        if (!ctx)
        {
            // same rule as the Python package's default_context(): FASTPAULI_DEVICE, else the launcher's LOCAL_RANK
            // (one process per GPU under torchrun / mpirun wrappers), else device 0
            int dev = 0;
            if (char const *env = getenv("FASTPAULI_DEVICE"))
                dev = atoi(env);
            else if (char const *lr = getenv("LOCAL_RANK"))
            {
                int n = 0;
                if (fp_device_count(&n) == FP_OK && n > 0)
                    dev = atoi(lr) % n;
            }
            FP_TRY(fp_ctx_create(dev, &ctx));
        }
        *out = ctx;
        return FP_OK;
    }

#define FP_DEFINE_ONESHOT(SFX, T, DT)                                                                                  \
    int fp_string_apply1d_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, int)         \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        return fp_string_apply(ctx, DT, n, codes, c, out, in, dim, 1, 1);                                              \
    }                                                                                                                  \
    int fp_string_apply_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, size_t B, int) \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        return fp_string_apply(ctx, DT, n, codes, c, out, in, dim, B, 1);                                              \
    }                                                                                                                  \
    int fp_string_expval_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, size_t B,     \
                               int)                                                                                    \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        return fp_string_expval(ctx, DT, n, codes, c, out, in, dim, B, 1);                                             \
    }                                                                                                                  \
    int fp_op_apply_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,     \
                          size_t B, int)                                                                               \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        fp_op *op = nullptr;                                                                                           \
        FP_TRY(fp_op_create(ctx, DT, n, S, codes, coeffs, &op));                                                       \
        int rc = fp_op_apply(ctx, op, out, in, dim, B, 1);                                                             \
        fp_op_destroy(op);                                                                                             \
        return rc;                                                                                                     \
    }                                                                                                                  \
    int fp_op_apply1d_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,   \
                            int par)                                                                                   \
    {                                                                                                                  \
        return fp_op_apply_##SFX(n, S, codes, coeffs, out, in, dim, 1, par);                                           \
    }                                                                                                                  \
    int fp_op_expval_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,    \
                           size_t B, int)                                                                              \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        fp_op *op = nullptr;                                                                                           \
        FP_TRY(fp_op_create(ctx, DT, n, S, codes, coeffs, &op));                                                       \
        int rc = fp_op_expval(ctx, op, out, in, dim, B, 1);                                                            \
        fp_op_destroy(op);                                                                                             \
        return rc;                                                                                                     \
    }                                                                                                                  \
    int fp_sop_apply_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out, const T *in,      \
                           size_t dim, size_t B, int)                                                                  \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        fp_sop *sop = nullptr;                                                                                         \
        FP_TRY(fp_sop_create(ctx, DT, n, S, codes, K, coeffs, &sop));                                                  \
        int rc = fp_sop_apply(ctx, sop, out, in, dim, B, 1);                                                           \
        fp_sop_destroy(sop);                                                                                           \
        return rc;                                                                                                     \
    }                                                                                                                  \
    int fp_sop_apply_weighted_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out,          \
                                    const T *in, const void *data, int data_is_f64, size_t dim, size_t B, int)         \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        fp_sop *sop = nullptr;                                                                                         \
        FP_TRY(fp_sop_create(ctx, DT, n, S, codes, K, coeffs, &sop));                                                  \
        int rc = fp_sop_apply_weighted(ctx, sop, out, in, data, data_is_f64, dim, B, 1);                               \
        fp_sop_destroy(sop);                                                                                           \
        return rc;                                                                                                     \
    }                                                                                                                  \
    int fp_sop_expval_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out, const T *in,     \
                            size_t dim, size_t B, int)                                                                 \
    {                                                                                                                  \
        fp_ctx *ctx;                                                                                                   \
        FP_TRY(fp_default_ctx(&ctx));                                                                                  \
        fp_sop *sop = nullptr;                                                                                         \
        FP_TRY(fp_sop_create(ctx, DT, n, S, codes, K, coeffs, &sop));                                                  \
        int rc = fp_sop_expval(ctx, sop, out, in, dim, B, 1);                                                          \
        fp_sop_destroy(sop);                                                                                           \
        return rc;                                                                                                     \
    }

    FP_DEFINE_ONESHOT(c128, double, FP_C128)
    FP_DEFINE_ONESHOT(c64, float, FP_C64)

} // extern "C"

namespace
{
// SummedPauliOp::square() on the device (square.cuh): host side = duplicate merge, partner table, staging
template <typename T>
int run_sop_square(fp_ctx *ctx, int n, size_t S, uint8_t const *codes, size_t K, std::complex<T> const *coeffs,
                   size_t n_sq, uint8_t const *sq_codes, std::complex<T> *coeffs_sq)
{
    // ---- merge duplicate input strings (sum of their coefficient rows: the same operators A_k)
    std::map<std::pair<uint64_t, uint64_t>, uint32_t> uniq;
    std::vector<uint64_t> xs, zs;
    std::vector<std::complex<T>> h;
    for (size_t s = 0; s < S; ++s)
    {
        StringMasks mk = make_masks(n, codes + s * static_cast<size_t>(n));
        auto key = std::make_pair(mk.x, mk.z);
        auto it = uniq.find(key);
        uint32_t u;
        if (it == uniq.end())
        {
            u = static_cast<uint32_t>(xs.size());
            uniq.emplace(key, u);
            xs.push_back(mk.x);
            zs.push_back(mk.z);
            h.resize(h.size() + K, std::complex<T>(0));
        }
        else
            u = it->second;
        for (size_t k = 0; k < K; ++k)
            h[static_cast<size_t>(u) * K + k] += coeffs[s * K + k];
    }
    uint32_t const Su = static_cast<uint32_t>(xs.size());
    uint32_t tsize = 16;
    while (tsize < 2 * Su)
        tsize <<= 1;
    std::vector<SqEntry> table(tsize, SqEntry{0, 0, 0xffffffffu, 0, 0});
    auto host_hash = [](uint64_t x, uint64_t z) {
        uint64_t v = (x * 0x9E3779B97F4A7C15ull) ^ (z * 0xC2B2AE3D27D4EB4Full);
        v ^= v >> 29;
        v *= 0xBF58476D1CE4E5B9ull;
        v ^= v >> 32;
        return static_cast<uint32_t>(v);
    };
    for (uint32_t u = 0; u < Su; ++u)
    {
        uint32_t slot = host_hash(xs[u], zs[u]) & (tsize - 1);
        while (table[slot].idx != 0xffffffffu)
            slot = (slot + 1) & (tsize - 1);
        table[slot] = SqEntry{xs[u], zs[u], u, static_cast<uint32_t>(__builtin_popcountll(xs[u] & zs[u])) & 3u, 0};
    }
    std::vector<uint64_t> xq(n_sq), zq(n_sq);
    for (size_t c = 0; c < n_sq; ++c)
    {
        StringMasks mk = make_masks(n, sq_codes + c * static_cast<size_t>(n));
        xq[c] = mk.x;
        zq[c] = mk.z;
    }
    uint64_t *d_xs = nullptr, *d_zs = nullptr, *d_xq = nullptr, *d_zq = nullptr;
    SqEntry *d_table = nullptr;
    Cx<T> *d_h = nullptr, *d_out = nullptr;
    std::vector<void *> allocs;
    auto cleanup = [&]() {
        for (void *a : allocs)
            cudaFree(a);
    };
    int rc = upload_vec(&d_xs, xs);
    if (rc == FP_OK) { allocs.push_back(d_xs); rc = upload_vec(&d_zs, zs); }
    if (rc == FP_OK) { allocs.push_back(d_zs); rc = upload_vec(&d_xq, xq); }
    if (rc == FP_OK) { allocs.push_back(d_xq); rc = upload_vec(&d_zq, zq); }
    if (rc == FP_OK) { allocs.push_back(d_zq); rc = upload_vec(&d_table, table); }
    if (rc == FP_OK)
    {
        allocs.push_back(d_table);
        std::vector<Cx<T>> hc(h.size());
        for (size_t i = 0; i < h.size(); ++i)
            hc[i] = Cx<T>{h[i].real(), h[i].imag()};
        rc = upload_vec(&d_h, hc);
    }
    if (rc == FP_OK)
    {
        allocs.push_back(d_h);
        // the (large) result lives in the context's grow-only scratch: no cudaMalloc / cudaFree of hundreds of MB per call
        rc = ctx->work_b.ensure(std::max<size_t>(16, n_sq * K * sizeof(Cx<T>)));
        if (rc == FP_OK)
            d_out = static_cast<Cx<T> *>(ctx->work_b.p);
    }
    if (rc != FP_OK)
    {
        cleanup();
        return rc;
    }
    uint32_t const kmax = kSqMaxKChunks * kSqThreads;
    for (size_t k0 = 0; k0 < K; k0 += kmax)
    {
        uint32_t const kn = static_cast<uint32_t>(std::min<size_t>(kmax, K - k0));
        sop_square_kernel<T><<<static_cast<unsigned>(n_sq), kSqThreads, 0, ctx->stream>>>(
            Su, d_xs, d_zs, d_table, tsize - 1, d_h + k0, static_cast<uint32_t>(K), kn, d_xq, d_zq, d_out + k0);
        ctx->launches++;
    }
    cudaError_t e = cudaMemcpyAsync(coeffs_sq, d_out, n_sq * K * sizeof(Cx<T>), cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        return set_err(FP_CUDA_ERROR, std::string("square: ") + cudaGetErrorString(e));
    }
    return FP_OK;
}
} // namespace

extern "C"
{
    int fp_sop_square(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes, size_t n_operators,
                      const void *coeffs, size_t n_sq, const uint8_t *sq_codes, void *coeffs_sq)
    {
        if (!ctx || !codes || !coeffs || !sq_codes || !coeffs_sq)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        if (n_qubits < 1 || n_qubits > 62)
            return set_err(FP_INVALID_ARGUMENT, "n_qubits must be in [1, 62]");
        if (n_strings == 0 || n_operators == 0 || n_sq == 0)
            return FP_OK;
        if (n_sq > 0x7fffffffull || n_strings > 0x7ffffffeull)
            return set_err(FP_UNSUPPORTED, "too many strings");
        if (is_device_ptr(coeffs))
            return set_err(FP_INVALID_ARGUMENT, "square: coefficients are host data (operator metadata)");
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        try
        {
            if (dtype == FP_C128)
                return run_sop_square<double>(ctx, n_qubits, n_strings, codes, n_operators,
                                              static_cast<std::complex<double> const *>(coeffs), n_sq, sq_codes,
                                              static_cast<std::complex<double> *>(coeffs_sq));
            return run_sop_square<float>(ctx, n_qubits, n_strings, codes, n_operators,
                                         static_cast<std::complex<float> const *>(coeffs), n_sq, sq_codes,
                                         static_cast<std::complex<float> *>(coeffs_sq));
        }
        catch (std::invalid_argument const &e)
        {
            return set_err(FP_INVALID_ARGUMENT, e.what());
        }
        catch (std::bad_alloc const &)
        {
            return set_err(FP_OUT_OF_MEMORY, "host allocation failed");
        }
    }
}
