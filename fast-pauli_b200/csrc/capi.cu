// C ABI of the B200-native fast-pauli hot path (declared in include/fastpauli_b200.h): contexts, memory helpers,
// PauliString and PauliOp entry points.  SummedPauliOp lives in sop_capi.cu, the sharded-state driver in sharded.cpp,
// the coset kernel family's host side in coset_launch.hpp, shared plumbing in capi_internal.hpp.
//
// Host-side responsibilities only: argument checks that mirror the reference's std::invalid_argument sites,
// host/device pointer staging, plan packing (pack.hpp), geometry selection and kernel launches (kernels.cuh).
// No compute happens on the host and there is no CPU fallback: without a CUDA device every compute call fails.
#include "capi_internal.hpp"
#include "coset_launch.hpp"
#include "dcoset.cuh"

namespace
{
thread_local std::string g_err;
// ---------------------------------------------------------------- register-resident coset path (K3c, rcoset.cuh)
template <typename T>
int get_rc_plan(DeviceOp<T> const &op, int n_qubits, int rr, typename DeviceOp<T>::RcPlanDev const **out)
{
    auto it = op.rc_plans.find(rr);
    if (it != op.rc_plans.end())
    {
        *out = &it->second;
        return FP_OK;
    }
    std::vector<CosetPassHost<T>> host = plan_coset<T>(op.host, n_qubits, rr, 0);
    if (host.size() != 1 || host[0].basis.r != rr)
        return set_err(FP_UNSUPPORTED, "register coset plan: the operator does not fit one pass");
    CosetPassHost<T> const &h = host[0];
    uint32_t const rows = 1u << rr;
    // strings ordered by (local gather mask, local z-mask); the planner may have split a large group into
    // several sub-groups with the same gather mask
    std::vector<uint32_t> ustart(rows * rows + 1, 0);
    std::vector<uint64_t> sz;
    std::vector<Cx<T>> sc;
    uint32_t present = 0;
    for (uint32_t xl = 0; xl < rows; ++xl)
        for (uint32_t zl = 0; zl < rows; ++zl)
        {
            for (size_t g = 0; g < h.gxl.size(); ++g)
            {
                if (h.gxl[g] != xl)
                    continue;
                for (uint32_t s = h.gstart[g]; s < h.gstart[g + 1]; ++s)
                {
                    if (h.szl[s] != zl)
                        continue;
                    sz.push_back(h.sz[s]);
                    sc.push_back(Cx<T>{h.sc[s].real(), h.sc[s].imag()});
                    present |= 1u << xl;
                }
            }
            ustart[xl * rows + zl + 1] = static_cast<uint32_t>(sz.size());
        }
    typename DeviceOp<T>::RcPlanDev d;
    d.rr = rr;
    for (int k = 0; k < kRcMaxRank; ++k)
    {
        d.view.basis[k] = k < rr ? h.basis.b[k] : 0;
        d.view.pivot[k] = k < rr ? static_cast<uint32_t>(h.basis.pivot[k]) : 0;
    }
    d.view.present = present;
    uint32_t *d_ustart = nullptr;
    uint64_t *d_sz = nullptr;
    Cx<T> *d_sc = nullptr;
    int rc = upload_vec(&d_ustart, ustart);
    if (rc == FP_OK) { d.allocs.push_back(d_ustart); rc = upload_vec(&d_sz, sz); }
    if (rc == FP_OK) { d.allocs.push_back(d_sz); rc = upload_vec(&d_sc, sc); }
    if (rc == FP_OK) d.allocs.push_back(d_sc);
    if (rc != FP_OK)
    {
        for (void *a : d.allocs)
            cudaFree(a);
        return rc;
    }
    d.view.ustart = d_ustart;
    d.view.sz = d_sz;
    d.view.scoef = d_sc;
    auto ins = op.rc_plans.emplace(rr, std::move(d));
    *out = &ins.first->second;
    return FP_OK;
}

template <typename T, int EPV, int RR, int LOG_NT, int MODE>
int launch_rcoset(fp_ctx *ctx, RcPassView<T> const &view, uint64_t n_cosets, uint64_t rowvecs, uint32_t log2tw,
                  uint32_t log2p, uint32_t nct, uint64_t n_blocks, uint32_t iters, size_t smem, void const *in, void *out, int beta,
                  void *partials, uint32_t Bpad)
{
    static PerDevice configured; // per template instance; try_rcoset never asks for more than 64 KiB
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(rcoset_kernel<T, EPV, RR, LOG_NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   64 * 1024));
        configured.set(ctx->device);
    }
    rcoset_kernel<T, EPV, RR, LOG_NT, MODE><<<static_cast<unsigned>(n_blocks * nct), 1 << LOG_NT, smem, ctx->stream>>>(
        view, n_cosets, rowvecs, log2tw, log2p, nct, iters, static_cast<CVec<T, EPV> const *>(in),
        static_cast<CVec<T, EPV> *>(out), beta, static_cast<Cx<T> *>(partials), Bpad);
    ctx->launches++;
    return FP_OK;
}

// K3d (dcoset.cuh): FP64 tensor-core dense-coset kernel, rank 4 (one warp per coset) or 5 (two); complex128 batches,
// and complex64 batches widened to double in registers (two columns per 16-byte vector)
template <int RR, int WPC, int PFD, int MODE, typename T>
int launch_dcoset(fp_ctx *ctx, RcPassView<T> const &view, uint32_t n_strings, uint64_t n_cosets, uint64_t rowvecs,
                  void const *in, void *out, int beta, uint64_t B)
{
    using Cfg = DcosetCfg<RR, WPC>;
    constexpr int EPV = sizeof(T) == 4 ? 2 : 1;
    constexpr size_t smem = MODE == 1 ? Cfg::smem_expval(EPV) : Cfg::smem;
    static int resident = 0; // CTAs per SM (per template instance): the kernel is persistent
    static PerDevice configured;
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(dcoset_kernel<RR, WPC, PFD, MODE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
        int nb = 0;
        FP_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dcoset_kernel<RR, WPC, PFD, MODE, T>, Cfg::NT, smem));
        resident = std::max(1, nb);
        configured.set(ctx->device);
    }
    uint64_t const sets = (n_cosets + Cfg::CPI - 1) / Cfg::CPI;
    unsigned const ny = MODE == 1 ? static_cast<unsigned>((rowvecs + Cfg::ECOLS - 1) / Cfg::ECOLS) : 1u;
    uint64_t const gx = std::min<uint64_t>(sets, std::max<uint64_t>(1, static_cast<uint64_t>(ctx->sm_count) * resident / ny));
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    if (MODE == 1)
        FP_TRY(ctx->partials.ensure(gx * Bpad * sizeof(Cx<double>)));
    dcoset_kernel<RR, WPC, PFD, MODE, T><<<dim3(static_cast<unsigned>(gx), ny), Cfg::NT, smem, ctx->stream>>>(
        view, n_strings, n_cosets, rowvecs, static_cast<CVec<T, EPV> const *>(in),
        MODE == 1 ? nullptr : static_cast<CVec<T, EPV> *>(out), beta, static_cast<Cx<double> *>(ctx->partials.p), Bpad);
    ctx->launches++;
    if (MODE == 1)
    {
        unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
        finalize_complex_kernel<double, T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
            static_cast<Cx<double> const *>(ctx->partials.p), gx, Bpad, B, static_cast<Cx<T> *>(out), beta);
        ctx->launches++;
    }
    return FP_OK;
}

// Returns FP_OK with *used = false when the operator / batch shape is left to the other kernels.
template <typename T, int MODE>
int try_rcoset(fp_ctx *ctx, DeviceOp<T> const &op, int n_qubits, void *out, void const *in, uint64_t dim, uint64_t B,
               int beta, bool *used)
{
    *used = false;
    constexpr int EPV = sizeof(T) == 4 ? 2 : 1;
    bool const enabled = ctx->rcoset_mode == 2 ||
                         (ctx->rcoset_mode == 1 && ctx->coset_mode == 1 && ctx->coset_log_twc < 0);
    if (!enabled || n_qubits <= 0 || dim != (1ull << n_qubits) || op.host.sz.size() < 2 || op.x_rank > kRcMaxRank)
        return FP_OK;
    if (pick_epv<T>(in, MODE == 1 ? in : out, B) != EPV)
        return FP_OK;
    uint64_t const rowvecs = B / EPV;
    if (rowvecs < 4 && ctx->rcoset_mode != 2)
        return FP_OK; // rows shorter than 64 bytes: coalescing must come from the row index (shared-memory tiles)
    int const rr = std::min(n_qubits, std::max(2, op.x_rank)); // 2-row threads keep too few bytes in flight
    {
        // ranks 4 and 5 are GEMM-shaped per coset (16x16 / 32x32 complex): FP64 tensor cores (complex64 batches are
        // widened in registers: measured at 20 qubits x 64 complex64, all Paulis on 4 / 5 qubits took 0.37 / 1.53 ms on
        // the SIMT kernels, bound by their shared-memory traffic)
        // The tensor-core form costs 2^rr complex FMAs per amplitude however few of the 2^rr coset masks occur; the
        // SIMT forms cost one per occurring mask.  Measured at 20 qubits x 64 columns (scripts/dispatch_sweep.py,
        // profiles/r01s3_dispatch_sweep.txt): rank 4 -- DMMA 0.45 ms apply / 0.53 ms expectation value against
        // 0.37 / 0.43 / 0.49 / 0.57 ms (apply) and 0.34 / 0.42 / 0.50 / 0.58 ms (expectation value) for the register
        // kernel at 4 / 8 / 12 / 16 masks; rank 5 -- DMMA 0.82-0.94 ms against 0.49 / 0.58 / 0.68 / 0.81 / 1.09 ms for
        // the shared-memory coset kernel at 5 / 8 / 12 / 16 / 24 masks.
        size_t const n_masks = op.host.gx.size();
        // complex64 batches: rank 5 only -- at rank 4 the SIMT register kernel is as fast (all 256 Paulis on 4 qubits,
        // 20 q x 64 complex64: 0.37 / 0.37 ms apply / expectation value against 0.39 / 0.48 ms here), at rank 5 the
        // tensor-core form halves the time (all 1024 Paulis on 5 qubits: 1.53 / 1.61 -> 0.77 / 0.92 ms)
        bool const dense_enough = rr == 4 ? (sizeof(T) == 8 && n_masks > (MODE == 1 ? 12u : 8u)) : n_masks > 16u;
        if ((ctx->dcoset == 2 || (ctx->dcoset == 1 && dense_enough)) && (rr == 4 || rr == 5) && rowvecs >= 8)
        {
            typename DeviceOp<T>::RcPlanDev const *dplan = nullptr;
            FP_TRY(get_rc_plan<T>(op, n_qubits, rr, &dplan));
            uint64_t const nc = 1ull << (n_qubits - rr);
            uint32_t const ns = static_cast<uint32_t>(op.host.sz.size());
            if (rr == 4)
                FP_TRY((launch_dcoset<4, 1, 4, MODE, T>(ctx, dplan->view, ns, nc, rowvecs, in, out, beta, B)));
            else
                FP_TRY((launch_dcoset<5, 2, 2, MODE, T>(ctx, dplan->view, ns, nc, rowvecs, in, out, beta, B)));
            *used = true;
            return FP_OK;
        }
    }
    if (rr > kRcMaxSimtRank || rr < 2)
        return FP_OK; // a 1-qubit register has no rank-2 coset: the generic kernel takes it
    int const log_nt = ctx->rcoset_log_nt == 8 ? 8 : 7;
    uint32_t const NT = 1u << log_nt;
    uint32_t log2tw = 0;
    while ((1ull << log2tw) < rowvecs && (1u << log2tw) < NT)
        ++log2tw;
    uint32_t const TW = 1u << log2tw, TY = NT / TW;
    size_t smem = 2 * static_cast<size_t>(TY) * (1u << (2 * rr)) * 2 * sizeof(T); // factor table + its z-mask sums
    if (MODE == 1)
        smem = std::max(smem, static_cast<size_t>(NT) * EPV * 2 * sizeof(T));
    if (smem > 64 * 1024)
        return FP_OK;
    typename DeviceOp<T>::RcPlanDev const *plan = nullptr;
    FP_TRY(get_rc_plan<T>(op, n_qubits, rr, &plan));
    uint64_t const n_cosets = 1ull << (n_qubits - rr);
    uint32_t const nct = static_cast<uint32_t>((rowvecs + TW - 1) / TW);
    uint64_t const n_sets = (n_cosets + TY - 1) / TY;
    uint64_t iters = 1;
    if (MODE == 1)
    {
        uint64_t const target = static_cast<uint64_t>(ctx->sm_count) * ctx->rc_expval_ctas_per_sm;
        uint64_t const want_cb = std::max<uint64_t>(1, target / nct);
        iters = std::min<uint64_t>(std::max<uint64_t>(1, (n_sets + want_cb - 1) / want_cb), 4096);
    }
    while ((n_sets + iters - 1) / iters * nct > 0x7fffffffull)
        iters *= 2;
    uint64_t const n_blocks = (n_sets + iters - 1) / iters;
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    if (MODE == 1)
        FP_TRY(ctx->partials.ensure(n_blocks * Bpad * 2 * sizeof(T)));
    // lanes cooperating on one entry of the z-mask sums: as many as keep every thread busy, at most a warp
    uint32_t log2p = 0;
    while (log2p < 5 && (static_cast<uint64_t>(TY) << (2 * rr + log2p + 1)) <= NT)
        ++log2p;
    bool launched = false;
#define FP_RC_CASE(RRV, LNT)                                                                                           \
    if (rr == RRV && log_nt == LNT)                                                                                    \
    {                                                                                                                  \
        FP_TRY((launch_rcoset<T, EPV, RRV, LNT, MODE>(ctx, plan->view, n_cosets, rowvecs, log2tw, log2p, nct, n_blocks, \
                                                      static_cast<uint32_t>(iters), smem, in, out, beta,               \
                                                      ctx->partials.p, Bpad)));                                        \
        launched = true;                                                                                               \
    }
    FP_RC_CASE(2, 7)
    FP_RC_CASE(3, 7)
    FP_RC_CASE(4, 7)
    FP_RC_CASE(2, 8)
    FP_RC_CASE(3, 8)
    FP_RC_CASE(4, 8)
#undef FP_RC_CASE
    if (!launched)
        return FP_OK; // no instantiation for this shape: never claim a result that was not computed
    if (MODE == 1)
    {
        unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
        if (n_blocks > 2048)
        {
            // thousands of partial rows (many small independent CTAs keep HBM busy the way the apply form does): fold
            // them with a wide first stage -- the single-stage finaliser has only B / 32 CTAs
            uint64_t const per = (n_blocks + 127) / 128;
            unsigned const slices = static_cast<unsigned>((n_blocks + per - 1) / per);
            FP_TRY(ctx->partials2.ensure(static_cast<size_t>(slices) * Bpad * sizeof(Cx<double>)));
            fold_partials_kernel<T><<<dim3(fgrid, slices), dim3(kFinX, kFinY), 0, ctx->stream>>>(
                static_cast<Cx<T> const *>(ctx->partials.p), n_blocks, per, Bpad, B,
                static_cast<Cx<double> *>(ctx->partials2.p));
            finalize_complex_kernel<double, T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
                static_cast<Cx<double> const *>(ctx->partials2.p), slices, Bpad, B, static_cast<Cx<T> *>(out), beta);
            ctx->launches += 2;
        }
        else
        {
            finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
                static_cast<Cx<T> const *>(ctx->partials.p), n_blocks, Bpad, B, static_cast<Cx<T> *>(out), beta);
            ctx->launches++;
        }
    }
    *used = true;
    return FP_OK;
}

// ---------------------------------------------------------------- launchers (all pointers are device pointers here)
template <typename T, int EPV, int MODE, bool INLINE1>
void launch_op_v(fp_ctx *ctx, GeomSel const &gs, OpView<T> const &view, void const *in, void *out, void *partials,
                 int beta, void const *bra = nullptr)
{
    auto const *din = static_cast<CVec<T, EPV> const *>(in);
    auto const *dbra = bra ? static_cast<CVec<T, EPV> const *>(bra) : din;
    auto *dout = static_cast<CVec<T, EPV> *>(out);
    auto *dpart = static_cast<Cx<T> *>(partials);
    dim3 grid(static_cast<unsigned>(gs.grid));
#define FP_LAUNCH_OP(VV, JJ)                                                                                           \
    op_kernel<T, EPV, VV, JJ, MODE, INLINE1><<<grid, kThreads, 0, ctx->stream>>>(view, gs.g, din, dout, dpart, beta, dbra)
    if (gs.J == 4)
    {
        if constexpr (!INLINE1)
        {
            if (gs.V == 2)
                FP_LAUNCH_OP(2, 4);
            else
                FP_LAUNCH_OP(1, 4);
        }
    }
    else if (gs.V == 4)
        FP_LAUNCH_OP(4, 1);
    else
        FP_LAUNCH_OP(1, 1);
#undef FP_LAUNCH_OP
    ctx->launches++;
}

template <typename T>
int run_op_apply(fp_ctx *ctx, DeviceOp<T> const &op, void *out, void const *in, uint64_t dim, uint64_t B, int beta,
                 int n_qubits = 0)
{
    if (dim == 0 || B == 0)
        return FP_OK;
    if (op.host.sz.empty())
    {
        // operator with no strings acts as zero (cannot happen through the checked entry points: dim() == 0)
        if (!beta)
            FP_CU(cudaMemsetAsync(out, 0, dim * B * 2 * sizeof(T), ctx->stream));
        return FP_OK;
    }
    FP_TRY(check_align(in, 2 * sizeof(T), "states"));
    FP_TRY(check_align(out, 2 * sizeof(T), "new_states"));
    if (op.host.sz.size() > 1 && n_qubits > 0)
    {
        bool used = false;
        if (B == 1 && dim == (1ull << n_qubits))
        {
            FP_TRY((try_single_state<T>(ctx, op, n_qubits, out, in, beta, &used)));
            if (used)
                return FP_OK;
        }
        FP_TRY((try_rcoset<T, 0>(ctx, op, n_qubits, out, in, dim, B, beta, &used)));
        if (used)
            return FP_OK;
        FP_TRY((try_coset<T, 0>(ctx, op, n_qubits, out, in, dim, B, beta, nullptr, nullptr, &used)));
        if (used)
            return FP_OK;
    }
    int const epv = pick_epv<T>(in, out, B);
    uint64_t const rowvecs = B / epv;
    bool const single = op.host.sz.size() == 1;
    GeomSel gs = choose_geom(ctx, dim, dim, rowvecs, 2 * sizeof(T) * epv, op.host.gx.size() > 1, false, 1, !single);
    FP_TRY(check_grid(gs.grid));
    OpView<T> view = op.view();
    if constexpr (sizeof(T) == 4)
    {
        if (epv == 2)
        {
            if (single)
                launch_op_v<T, 2, 0, true>(ctx, gs, view, in, out, nullptr, beta);
            else
                launch_op_v<T, 2, 0, false>(ctx, gs, view, in, out, nullptr, beta);
            return FP_OK;
        }
    }
    if (single)
        launch_op_v<T, 1, 0, true>(ctx, gs, view, in, out, nullptr, beta);
    else
        launch_op_v<T, 1, 0, false>(ctx, gs, view, in, out, nullptr, beta);
    return FP_OK;
}

template <typename T>
int run_op_expval(fp_ctx *ctx, DeviceOp<T> const &op, void *out /* B complex, device */, void const *in, uint64_t dim,
                  uint64_t B, int beta, int n_qubits = 0, void const *bra = nullptr)
{
    if (B == 0)
        return FP_OK;
    if (dim == 0 || op.host.sz.empty())
    {
        if (!beta)
            FP_CU(cudaMemsetAsync(out, 0, B * 2 * sizeof(T), ctx->stream));
        return FP_OK;
    }
    FP_TRY(check_align(in, 2 * sizeof(T), "states"));
    if (bra)
        FP_TRY(check_align(bra, 2 * sizeof(T), "bra states"));
    if (op.host.sz.size() > 1 && n_qubits > 0 && (!bra || bra == in))
    {
        bool used = false;
        FP_TRY((try_rcoset<T, 1>(ctx, op, n_qubits, out, in, dim, B, beta, &used)));
        if (used)
            return FP_OK;
        FP_TRY((try_coset<T, 1>(ctx, op, n_qubits, out, in, dim, B, beta, nullptr, nullptr, &used)));
        if (used)
            return FP_OK;
    }
    int const epv = pick_epv<T>(in, bra ? bra : in, B);
    uint64_t const rowvecs = B / epv;
    GeomSel gs = choose_geom(ctx, dim, dim, rowvecs, 2 * sizeof(T) * epv, op.host.gx.size() > 1, true, 1, true);
    FP_TRY(check_grid(gs.grid));
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    if (B > 0xfffffff0ull)
        return set_err(FP_UNSUPPORTED, "n_states too large");
    gs.g.Bpad = Bpad;
    FP_TRY(ctx->partials.ensure(gs.g.nRowBlocks * Bpad * 2 * sizeof(T)));
    OpView<T> view = op.view();
    bool done = false;
    if constexpr (sizeof(T) == 4)
    {
        if (epv == 2)
        {
            launch_op_v<T, 2, 1, false>(ctx, gs, view, in, nullptr, ctx->partials.p, 0, bra);
            done = true;
        }
    }
    if (!done)
        launch_op_v<T, 1, 1, false>(ctx, gs, view, in, nullptr, ctx->partials.p, 0, bra);
    unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
    finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(static_cast<Cx<T> const *>(ctx->partials.p),
                                                                 gs.g.nRowBlocks, Bpad, B, static_cast<Cx<T> *>(out),
                                                                 beta);
    ctx->launches++;
    return FP_OK;
}

// PauliString::expectation_value through the paired kernel: each amplitude is read once.
template <typename T>
int run_string_expval(fp_ctx *ctx, StringMasks const &mk, std::complex<T> coeff, void *out, void const *in,
                      uint64_t dim, uint64_t B, int beta)
{
    if (B == 0)
        return FP_OK;
    if (dim == 0)
    {
        if (!beta)
            FP_CU(cudaMemsetAsync(out, 0, B * 2 * sizeof(T), ctx->stream));
        return FP_OK;
    }
    FP_TRY(check_align(in, 2 * sizeof(T), "states"));
    int const epv = pick_epv<T>(in, in, B);
    uint64_t const rowvecs = B / epv;
    PairChunk ch{};
    ch.x = mk.x;
    ch.s0 = 0;
    ch.count = 1;
    ch.diag = mk.x == 0;
    ch.hbit = mk.x ? 63u - static_cast<uint32_t>(__builtin_clzll(mk.x)) : 0;
    uint64_t const rows = ch.diag ? dim : dim / 2;
    GeomSel gs = choose_geom(ctx, rows, dim, rowvecs, 2 * sizeof(T) * epv, false, true);
    FP_TRY(check_grid(gs.grid));
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    gs.g.Bpad = Bpad;
    uint64_t const slot_stride = gs.g.nRowBlocks * Bpad;
    FP_TRY(ctx->partials.ensure(slot_stride * sizeof(T)));
    T *part = static_cast<T *>(ctx->partials.p);
    bool done = false;
    if constexpr (sizeof(T) == 4)
    {
        if (epv == 2)
        {
            launch_pairs_v<T, 2, 1>(ctx, gs, nullptr, nullptr, nullptr, ch, mk.z, mk.ny & 1u, 1, dim, in, part,
                                    slot_stride);
            done = true;
        }
    }
    if (!done)
        launch_pairs_v<T, 1, 1>(ctx, gs, nullptr, nullptr, nullptr, ch, mk.z, mk.ny & 1u, 1, dim, in, part, slot_stride);
    // factor = coeff * (-i)^nY * (1 | 2 | 2i)
    std::complex<double> f = times_phase(std::complex<double>(coeff.real(), coeff.imag()), mk.ny);
    if (!ch.diag)
        f *= (mk.ny & 1u) ? std::complex<double>(0, 2) : std::complex<double>(2, 0);
    unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
    finalize_pairs_string_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(part, gs.g.nRowBlocks, Bpad, B, f.real(), f.imag(),
                                                                      static_cast<Cx<T> *>(out), beta);
    ctx->launches++;
    return FP_OK;
}

// ---------------------------------------------------------------- typed front-ends over fp_op
// Host-resident PauliString::apply_batch (PS:377-436) as a three-stage pipeline.  An aligned block of 2^m rows of
// the output depends on exactly one such block of the input (block index ^ (x >> m), rows permuted by the low bits
// of x inside it), so the batch streams through the GPU in chunks: copy engine 1 uploads chunk j+1 while the
// kernel permutes chunk j and copy engine 2 downloads chunk j-1 -- both PCIe directions stay busy for the whole
// call instead of upload, kernel and download running back to back.  The sign of the block index is folded into
// the coefficient, so every chunk runs the ordinary single-string kernel K1 on (x_low, z_low).
template <typename T>
int pipelined_string_apply(fp_ctx *ctx, StringMasks const &mk, std::complex<T> coeff, void *out, void const *in,
                           uint64_t dim, uint64_t B)
{
    size_t const rowbytes = B * 2 * sizeof(T);
    int m = 0;
    while ((2ull << m) * rowbytes <= ctx->pipeline_chunk_bytes && (2ull << m) <= dim)
        ++m;
    uint64_t const rows = 1ull << m, n_chunks = dim >> m;
    size_t const cbytes = rows * rowbytes;
    if (!ctx->h2d_stream)
    {
        FP_CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
        FP_CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i)
        {
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_in[i], cudaEventDisableTiming));
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_k[i], cudaEventDisableTiming));
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_out[i], cudaEventDisableTiming));
        }
        FP_CU(cudaEventCreateWithFlags(&ctx->pipe_start, cudaEventDisableTiming));
    }
    FP_TRY(ctx->stage_in.ensure(3 * cbytes));
    FP_TRY(ctx->stage_out.ensure(3 * cbytes));
    // work already queued on the context's stream may still use the scratch buffers
    FP_CU(cudaEventRecord(ctx->pipe_start, ctx->stream));
    FP_CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_start, 0));
    uint64_t const xlo = mk.x & (rows - 1), xhi = mk.x >> m, zlo = mk.z & (rows - 1);
    std::complex<T> const c0 = times_phase(coeff, mk.ny);
    for (uint64_t j = 0; j < n_chunks; ++j)
    {
        int const b = static_cast<int>(j % 3);
        auto *d_in = static_cast<unsigned char *>(ctx->stage_in.p) + static_cast<size_t>(b) * cbytes;
        auto *d_out = static_cast<unsigned char *>(ctx->stage_out.p) + static_cast<size_t>(b) * cbytes;
        uint64_t const jo = j ^ xhi; // output block fed by input block j
        if (j >= 3)
            FP_CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_k[b], 0)); // kernel j-3 has consumed this buffer
        FP_CU(cudaMemcpyAsync(d_in, static_cast<unsigned char const *>(in) + j * cbytes, cbytes,
                              cudaMemcpyHostToDevice, ctx->h2d_stream));
        FP_CU(cudaEventRecord(ctx->pipe_in[b], ctx->h2d_stream));
        FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_in[b], 0));
        if (j >= 3)
            FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_out[b], 0)); // download j-3 has drained this buffer
        DeviceOp<T> op;
        op.host.gx.assign(1, xlo);
        op.host.gstart.assign({0u, 1u});
        op.host.sz.assign(1, zlo);
        op.host.sc.assign(1, (__builtin_popcountll((jo << m) & mk.z) & 1) ? -c0 : c0);
        FP_TRY(run_op_apply<T>(ctx, op, d_out, d_in, rows, B, 0));
        FP_CU(cudaEventRecord(ctx->pipe_k[b], ctx->stream));
        FP_CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->pipe_k[b], 0));
        FP_CU(cudaMemcpyAsync(static_cast<unsigned char *>(out) + jo * cbytes, d_out, cbytes, cudaMemcpyDeviceToHost,
                              ctx->d2h_stream));
        FP_CU(cudaEventRecord(ctx->pipe_out[b], ctx->d2h_stream));
    }
    for (int b = 0; b < 3 && static_cast<uint64_t>(b) < n_chunks; ++b)
        FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_out[b], 0)); // the call's stream completes after every download
    return FP_OK;
}

// Host-resident PauliOp::apply (PO:399-468) as a pipeline over COLUMN blocks: every hot-path formula is independent
// per batch column, so the (dim, B) host batch is cut into blocks of `cols` columns; block j+1 is uploaded (strided 2-D
// copy into a dense device block) while block j runs the ordinary kernels and block j-1 is downloaded, so both PCIe
// directions are busy for the whole call.  Three device blocks per direction.
template <typename T>
int pipelined_op_apply(fp_ctx *ctx, DeviceOp<T> const &op, void *out, void const *in, uint64_t dim, uint64_t B,
                       int n_qubits, bool *used)
{
    *used = false;
    size_t const esize = 2 * sizeof(T);
    // block width: a multiple of 16 vectors (256-byte row segments: the widest coset tile and efficient strided DMA),
    // at most B / 3 so that at least three blocks are in flight, about pipeline_chunk_bytes * 8 per block
    uint64_t const vec_cols = 16 / esize;          // columns per 16-byte vector
    uint64_t cols = 16 * vec_cols;                 // 256 bytes per row (128-byte segments: strided DMA drops to 2/3)
    if (char const *env = getenv("FASTPAULI_PIPE_VECS"))
        cols = std::max<uint64_t>(1, strtoull(env, nullptr, 10)) * vec_cols;
    if (B % cols != 0 || B / cols < 3)
        return FP_OK;
    while (B % (2 * cols) == 0 && B / (2 * cols) >= 4 && dim * (2 * cols) * esize <= (256ull << 20))
        cols *= 2;
    uint64_t const n_blocks = B / cols;
    size_t const bbytes = dim * cols * esize;
    if (!ctx->h2d_stream)
    {
        FP_CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
        FP_CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i)
        {
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_in[i], cudaEventDisableTiming));
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_k[i], cudaEventDisableTiming));
            FP_CU(cudaEventCreateWithFlags(&ctx->pipe_out[i], cudaEventDisableTiming));
        }
        FP_CU(cudaEventCreateWithFlags(&ctx->pipe_start, cudaEventDisableTiming));
    }
    FP_TRY(ctx->stage_in.ensure(3 * bbytes));
    FP_TRY(ctx->stage_out.ensure(3 * bbytes));
    FP_CU(cudaEventRecord(ctx->pipe_start, ctx->stream));
    FP_CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_start, 0));
    size_t const host_pitch = B * esize, dev_pitch = cols * esize;
    for (uint64_t j = 0; j < n_blocks; ++j)
    {
        int const b = static_cast<int>(j % 3);
        auto *d_in = static_cast<unsigned char *>(ctx->stage_in.p) + static_cast<size_t>(b) * bbytes;
        auto *d_out = static_cast<unsigned char *>(ctx->stage_out.p) + static_cast<size_t>(b) * bbytes;
        if (j >= 3)
            FP_CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_k[b], 0)); // kernel j-3 has consumed this block
        FP_CU(cudaMemcpy2DAsync(d_in, dev_pitch, static_cast<unsigned char const *>(in) + j * dev_pitch, host_pitch,
                                dev_pitch, dim, cudaMemcpyHostToDevice, ctx->h2d_stream));
        FP_CU(cudaEventRecord(ctx->pipe_in[b], ctx->h2d_stream));
        FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_in[b], 0));
        if (j >= 3)
            FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_out[b], 0)); // download j-3 has drained this block
        FP_TRY(run_op_apply<T>(ctx, op, d_out, d_in, dim, cols, 0, n_qubits));
        FP_CU(cudaEventRecord(ctx->pipe_k[b], ctx->stream));
        FP_CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->pipe_k[b], 0));
        FP_CU(cudaMemcpy2DAsync(static_cast<unsigned char *>(out) + j * dev_pitch, host_pitch, d_out, dev_pitch, dev_pitch,
                                dim, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        FP_CU(cudaEventRecord(ctx->pipe_out[b], ctx->d2h_stream));
    }
    for (int b = 0; b < 3 && static_cast<uint64_t>(b) < n_blocks; ++b)
        FP_CU(cudaStreamWaitEvent(ctx->stream, ctx->pipe_out[b], 0));
    *used = true;
    return FP_OK;
}

} // namespace

// ================================================================ extern "C"
extern "C"
{

    const char *fp_last_error(void)
    {
        return g_err.c_str();
    }

    // Measured FP64 FMA throughput of the context's GPU (dependent DFMA chains on every SM, best of 3): the
    // denominator bench.py quotes FP64-bound calls against.
    int fp_measure_fp64_tflops(fp_ctx *ctx, double *tflops)
    {
        if (!ctx || !tflops)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        int const ctas = ctx->sm_count * 4, threads = 256, iters = 4096;
        FP_TRY(ctx->work_a.ensure(static_cast<size_t>(ctas) * threads * sizeof(double)));
        cudaEvent_t e0, e1;
        FP_CU(cudaEventCreate(&e0));
        FP_CU(cudaEventCreate(&e1));
        float best = 0;
        for (int rep = 0; rep < 4; ++rep)
        {
            FP_CU(cudaEventRecord(e0, ctx->stream));
            fpk::fp64_peak_kernel<<<ctas, threads, 0, ctx->stream>>>(static_cast<double *>(ctx->work_a.p), iters);
            FP_CU(cudaEventRecord(e1, ctx->stream));
            FP_CU(cudaEventSynchronize(e1));
            float ms = 0;
            FP_CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && (best == 0 || ms < best))
                best = ms;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        double const fma = static_cast<double>(ctas) * threads * iters * 16.0;
        *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
        return FP_OK;
    }

    int fp_internal_set_error(int code, const char *msg)
    {
        g_err = msg ? msg : "";
        return code;
    }

    int fp_version(void)
    {
        return 100; // 0.1.0
    }

    int fp_device_count(int *count)
    {
        if (!count)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        *count = 0;
        cudaError_t e = cudaGetDeviceCount(count);
        if (e != cudaSuccess)
        {
            (void)cudaGetLastError();
            *count = 0;
            return set_err(FP_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        }
        return FP_OK;
    }

    int fp_ctx_create(int device, fp_ctx **out)
    {
        if (!out)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        *out = nullptr;
        int n = 0;
        FP_TRY(fp_device_count(&n));
        if (n == 0)
            return set_err(FP_NO_DEVICE, "no CUDA device visible: fastpauli_b200 has no CPU fallback");
        if (device < 0 || device >= n)
            return set_err(FP_INVALID_ARGUMENT, "device index out of range");
        DeviceGuard guard(device); // the caller's current device (e.g. torch's) is restored on return
        cudaDeviceProp prop;
        FP_CU(cudaGetDeviceProperties(&prop, device));
        std::unique_ptr<fp_ctx> ctx(new fp_ctx);
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        if (prop.l2CacheSize > 0)
            ctx->l2_budget = static_cast<size_t>(prop.l2CacheSize) / 3; // one die's worth of L2 minus headroom
        if (char const *env = getenv("FASTPAULI_L2_BUDGET"))
            ctx->l2_budget = strtoull(env, nullptr, 10);
        if (char const *env = getenv("FASTPAULI_COSET"))
            ctx->coset_mode = atoi(env);
        if (char const *env = getenv("FASTPAULI_COSET_LOG_TWC"))
            ctx->coset_log_twc = atoi(env);
        if (char const *env = getenv("FASTPAULI_COSET_LOG_NT"))
            ctx->coset_log_nt = atoi(env);
        if (char const *env = getenv("FASTPAULI_COSET_VPT"))
            ctx->coset_vpt = atoi(env);
        if (char const *env = getenv("FASTPAULI_COSET_WIDE"))
            ctx->coset_wide_cta = atoi(env) != 0;
        if (char const *env = getenv("FASTPAULI_COSET_FEW"))
        {
            ctx->coset_few = atoi(env) == 5 ? 1 : atoi(env);
            ctx->coset_pair_all = atoi(env) == 5;
        }
        if (char const *env = getenv("FASTPAULI_COSET_FEW_CT"))
            ctx->coset_few_ct = atoi(env);
        if (char const *env = getenv("FASTPAULI_PIPELINE"))
            ctx->pipeline = atoi(env) != 0;
        if (char const *env = getenv("FASTPAULI_PIPELINE_CHUNK"))
            ctx->pipeline_chunk_bytes = std::max<size_t>(1 << 16, strtoull(env, nullptr, 10));
        if (char const *env = getenv("FASTPAULI_ETILE"))
            ctx->etile = atoi(env) != 0;
        if (char const *env = getenv("FASTPAULI_WTILE"))
            ctx->wtile = atoi(env) != 0;
        if (char const *env = getenv("FASTPAULI_RCOSET"))
            ctx->rcoset_mode = atoi(env);
        if (char const *env = getenv("FASTPAULI_DCOSET"))
            ctx->dcoset = atoi(env);
        if (char const *env = getenv("FASTPAULI_RC_EXPVAL_CTAS_PER_SM"))
            ctx->rc_expval_ctas_per_sm = std::max(1, atoi(env));
        if (char const *env = getenv("FASTPAULI_RCOSET_LOG_NT"))
            ctx->rcoset_log_nt = atoi(env);
        if (char const *env = getenv("FASTPAULI_ZERO_COPY"))
            ctx->zero_copy = atoi(env) != 0;
        if (char const *env = getenv("FASTPAULI_TENSOR_CORE"))
            ctx->tensor_core = atoi(env) != 0;
        if (prop.major != 10)
            ctx->tensor_core = false; // tcgen05 exists on sm_100 only
        FP_CU(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        *out = ctx.release();
        return FP_OK;
    }

    int fp_ctx_destroy(fp_ctx *ctx)
    {
        if (!ctx)
            return FP_OK;
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (Scratch *s : {&ctx->stage_in, &ctx->stage_out, &ctx->stage_data, &ctx->partials, &ctx->partials2, &ctx->work_a, &ctx->work_b,
                           &ctx->meta})
            s->release();
        if (ctx->own_stream)
            cudaStreamDestroy(ctx->own_stream);
        if (ctx->h2d_stream)
            cudaStreamDestroy(ctx->h2d_stream);
        if (ctx->d2h_stream)
            cudaStreamDestroy(ctx->d2h_stream);
        for (int i = 0; i < 3; ++i)
            for (cudaEvent_t e : {ctx->pipe_in[i], ctx->pipe_k[i], ctx->pipe_out[i]})
                if (e)
                    cudaEventDestroy(e);
        if (ctx->pipe_start)
            cudaEventDestroy(ctx->pipe_start);
        delete ctx;
        return FP_OK;
    }

    int fp_ctx_device(const fp_ctx *ctx, int *device)
    {
        if (!ctx || !device)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        *device = ctx->device;
        return FP_OK;
    }

    int fp_ctx_pci_bus_id(const fp_ctx *ctx, char *buf, int len)
    {
        if (!ctx || !buf || len < 16)
            return set_err(FP_INVALID_ARGUMENT, "null pointer or buffer shorter than 16 bytes");
        FP_CU(cudaDeviceGetPCIBusId(buf, len, ctx->device));
        return FP_OK;
    }

    int fp_ctx_set_stream(fp_ctx *ctx, void *cuda_stream, int external)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        // external != 0: run on exactly this cudaStream_t -- including 0, the legacy default stream PyTorch uses
        // unless told otherwise; external == 0 restores the context's own stream
        ctx->stream = external ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
        return FP_OK;
    }

    int fp_ctx_set_async(fp_ctx *ctx, int async)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        ctx->async = async != 0;
        return FP_OK;
    }

    int fp_ctx_set_zero_copy(fp_ctx *ctx, int enable)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        ctx->zero_copy = enable != 0;
        return FP_OK;
    }

    int fp_ctx_set_tensor_core(fp_ctx *ctx, int enable)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        ctx->tensor_core = enable != 0;
        return FP_OK;
    }

    int fp_ctx_set_coset(fp_ctx *ctx, int mode, int log_twc, int log_nt)
    {
        if (!ctx || mode < 0 || mode > 2 || log_twc > 4 || !(log_nt == 0 || log_nt == 7 || log_nt == 8))
            return set_err(FP_INVALID_ARGUMENT, "bad coset mode");
        ctx->coset_mode = mode;
        ctx->coset_log_twc = log_twc;
        ctx->coset_log_nt = log_nt;
        return FP_OK;
    }

    int fp_ctx_set_pipeline(fp_ctx *ctx, int enable, size_t min_bytes, size_t chunk_bytes)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        ctx->pipeline = enable != 0;
        if (min_bytes)
            ctx->pipeline_min_bytes = min_bytes;
        if (chunk_bytes)
            ctx->pipeline_chunk_bytes = chunk_bytes;
        return FP_OK;
    }

    int fp_ctx_set_coset_few(fp_ctx *ctx, int mode, int column_tiles_per_cta)
    {
        if (!ctx || mode < 0 || mode > 5 || column_tiles_per_cta < 0)
            return set_err(FP_INVALID_ARGUMENT, "bad few-mask coset mode");
        ctx->coset_few = mode == 5 ? 1 : mode;
        ctx->coset_pair_all = mode == 5;
        ctx->coset_few_ct = column_tiles_per_cta;
        return FP_OK;
    }

    int fp_ctx_coset_kernels_used(fp_ctx *ctx, uint32_t *mask, int reset)
    {
        if (!ctx || !mask)
            return set_err(FP_INVALID_ARGUMENT, "null argument");
        *mask = ctx->coset_kernels;
        if (reset)
            ctx->coset_kernels = 0;
        return FP_OK;
    }

    int fp_ctx_set_rcoset(fp_ctx *ctx, int mode, int log_nt)
    {
        if (!ctx || mode < 0 || mode > 2 || !(log_nt == 0 || log_nt == 7 || log_nt == 8))
            return set_err(FP_INVALID_ARGUMENT, "bad register-coset mode");
        ctx->rcoset_mode = mode;
        ctx->rcoset_log_nt = log_nt == 0 ? 7 : log_nt;
        return FP_OK;
    }

    int fp_ctx_last_gemm_engine(const fp_ctx *ctx, int *engine)
    {
        if (!ctx || !engine)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        *engine = ctx->last_gemm_engine;
        return FP_OK;
    }

    int fp_ctx_set_l2_budget(fp_ctx *ctx, size_t bytes)
    {
        if (!ctx || bytes == 0)
            return set_err(FP_INVALID_ARGUMENT, "null context or zero budget");
        ctx->l2_budget = bytes;
        return FP_OK;
    }

    int fp_ctx_sync(fp_ctx *ctx)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        FP_CU(cudaStreamSynchronize(ctx->stream));
        return FP_OK;
    }

    int fp_ctx_launch_count(const fp_ctx *ctx, uint64_t *count)
    {
        if (!ctx || !count)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        *count = ctx->launches;
        return FP_OK;
    }

    // ------------------------------------------------------------ memory / timing
    int fp_device_malloc(fp_ctx *ctx, size_t bytes, void **ptr)
    {
        if (!ctx || !ptr)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        *ptr = nullptr;
        FP_CU(cudaMalloc(ptr, bytes ? bytes : 16));
        return FP_OK;
    }
    int fp_device_free(fp_ctx *ctx, void *ptr)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        FP_CU(cudaFree(ptr));
        return FP_OK;
    }
    int fp_host_malloc(fp_ctx *ctx, size_t bytes, void **ptr)
    {
        if (!ctx || !ptr)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        *ptr = nullptr;
        FP_CU(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
        return FP_OK;
    }
    int fp_host_free(fp_ctx *ctx, void *ptr)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        FP_CU(cudaFreeHost(ptr));
        return FP_OK;
    }
    int fp_memcpy_async(fp_ctx *ctx, void *dst, const void *src, size_t bytes)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        if (bytes)
            FP_CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
        return FP_OK;
    }
    int fp_memcpy(fp_ctx *ctx, void *dst, const void *src, size_t bytes)
    {
        FP_TRY(fp_memcpy_async(ctx, dst, src, bytes));
        return fp_ctx_sync(ctx);
    }
    int fp_memset(fp_ctx *ctx, void *dst, int value, size_t bytes)
    {
        if (!ctx)
            return set_err(FP_INVALID_ARGUMENT, "null context");
        DeviceGuard g(ctx->device);
        if (bytes)
            FP_CU(cudaMemsetAsync(dst, value, bytes, ctx->stream));
        return FP_OK;
    }
    int fp_device_mem_info(fp_ctx *ctx, size_t *free_bytes, size_t *total_bytes)
    {
        if (!ctx || !free_bytes || !total_bytes)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        FP_CU(cudaMemGetInfo(free_bytes, total_bytes));
        return FP_OK;
    }
    int fp_event_create(fp_event **ev)
    {
        if (!ev)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        std::unique_ptr<fp_event> e(new fp_event);
        FP_CU(cudaEventCreate(&e->ev));
        *ev = e.release();
        return FP_OK;
    }
    int fp_event_destroy(fp_event *ev)
    {
        if (ev)
        {
            cudaEventDestroy(ev->ev);
            delete ev;
        }
        return FP_OK;
    }
    int fp_event_record(fp_ctx *ctx, fp_event *ev)
    {
        if (!ctx || !ev)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        DeviceGuard g(ctx->device);
        FP_CU(cudaEventRecord(ev->ev, ctx->stream));
        return FP_OK;
    }
    int fp_event_elapsed_ms(fp_event *start, fp_event *stop, float *ms)
    {
        if (!start || !stop || !ms)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_CU(cudaEventSynchronize(stop->ev));
        FP_CU(cudaEventElapsedTime(ms, start->ev, stop->ev));
        return FP_OK;
    }

    int fp_fill_uniform(fp_ctx *ctx, int dtype, void *dst, uint64_t n_complex, uint64_t first_complex, uint64_t seed)
    {
        if (!ctx || (!dst && n_complex))
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        if (n_complex == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        Staged s;
        FP_TRY(stage_in(ctx, ctx->stage_out, dst, n_complex * csize(dtype), false, s));
        uint64_t n_real = 2 * n_complex;
        unsigned grid = static_cast<unsigned>(std::min<uint64_t>((n_real + 255) / 256, ctx->sm_count * 32ull));
        if (dtype == FP_C128)
            fill_uniform_kernel<double>
                <<<grid, 256, 0, ctx->stream>>>(static_cast<double *>(s.dev), n_real, 2 * first_complex, seed);
        else
            fill_uniform_kernel<float>
                <<<grid, 256, 0, ctx->stream>>>(static_cast<float *>(s.dev), n_real, 2 * first_complex, seed);
        ctx->launches++;
        FP_TRY(stage_back(ctx, s));
        return finish(ctx, s.staged);
    }

    // ------------------------------------------------------------ PauliString
    int fp_string_apply(fp_ctx *ctx, int dtype, int n_qubits, const uint8_t *codes, const void *coeff, void *out,
                        const void *in, size_t dim, size_t n_states, int accumulate)
    {
        if (!ctx || (!codes && n_qubits > 0) || !coeff)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        if (n_qubits < 0 || n_qubits > 62)
            return set_err(FP_INVALID_ARGUMENT, "n_qubits must be in [0, 62]");
        StringMasks mk;
        try
        {
            mk = make_masks(n_qubits, codes);
        }
        catch (std::invalid_argument const &e)
        {
            return set_err(FP_INVALID_ARGUMENT, e.what());
        }
        if (dim != dim_of(n_qubits)) // PS:275-278, PS:347-353
            return set_err(FP_INVALID_ARGUMENT, "[PauliString] states shape (" + std::to_string(dim) +
                                                    ") must match the dimension of the operators (" +
                                                    std::to_string(dim_of(n_qubits)) + ")");
        if (dim == 0 || n_states == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        size_t const bytes = dim * n_states * csize(dtype);
        if (ctx->pipeline && !accumulate && bytes >= ctx->pipeline_min_bytes && in && out && !is_device_ptr(in) &&
            !is_device_ptr(out))
        {
            auto const *ib = static_cast<unsigned char const *>(in);
            auto const *ob = static_cast<unsigned char const *>(out);
            if (ib + bytes <= ob || ob + bytes <= ib) // disjoint host buffers
            {
                FP_TRY(dtype == FP_C128
                           ? pipelined_string_apply<double>(ctx, mk, *static_cast<std::complex<double> const *>(coeff),
                                                            out, in, dim, n_states)
                           : pipelined_string_apply<float>(ctx, mk, *static_cast<std::complex<float> const *>(coeff),
                                                           out, in, dim, n_states));
                return finish(ctx, true);
            }
        }
        Staged sin, sout;
        FP_TRY(stage_in(ctx, ctx->stage_in, in, bytes, true, sin, true));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, bytes, accumulate != 0, sout, true));
        int rc;
        if (dtype == FP_C128)
        {
            DeviceOp<double> op;
            auto c = *static_cast<std::complex<double> const *>(coeff);
            op.host.gx = {mk.x};
            op.host.gstart = {0, 1};
            op.host.sz = {mk.z};
            op.host.sc = {times_phase(c, mk.ny)};
            rc = run_op_apply<double>(ctx, op, sout.dev, sin.dev, dim, n_states, accumulate);
        }
        else
        {
            DeviceOp<float> op;
            auto c = *static_cast<std::complex<float> const *>(coeff);
            op.host.gx = {mk.x};
            op.host.gstart = {0, 1};
            op.host.sz = {mk.z};
            op.host.sc = {times_phase(c, mk.ny)};
            rc = run_op_apply<float>(ctx, op, sout.dev, sin.dev, dim, n_states, accumulate);
        }
        FP_TRY(rc);
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged || sin.zero_copy || sout.zero_copy);
    }

    int fp_string_expval(fp_ctx *ctx, int dtype, int n_qubits, const uint8_t *codes, const void *coeff, void *out,
                         const void *in, size_t dim, size_t n_states, int accumulate)
    {
        if (!ctx || (!codes && n_qubits > 0) || !coeff)
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        if (n_qubits < 0 || n_qubits > 62)
            return set_err(FP_INVALID_ARGUMENT, "n_qubits must be in [0, 62]");
        StringMasks mk;
        try
        {
            mk = make_masks(n_qubits, codes);
        }
        catch (std::invalid_argument const &e)
        {
            return set_err(FP_INVALID_ARGUMENT, e.what());
        }
        if (dim != dim_of(n_qubits)) // PS:443-446
            return set_err(FP_INVALID_ARGUMENT, "[PauliString] states shape (" + std::to_string(dim) +
                                                    ") must match the dimension of the operators (" +
                                                    std::to_string(dim_of(n_qubits)) + ")");
        if (n_states == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        Staged sin, sout;
        // large pinned batches: the copy engine uploads faster (~55 GB/s) than the kernel's own reads over PCIe
        // (~51 GB/s), so above the pipeline threshold the batch is staged instead of being read in place
        bool const in_place = !(ctx->pipeline && dim * n_states * csize(dtype) >= ctx->pipeline_min_bytes);
        FP_TRY(stage_in(ctx, ctx->stage_in, in, dim * n_states * csize(dtype), true, sin, in_place));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, n_states * csize(dtype), accumulate != 0, sout));
        int rc;
        if (dtype == FP_C128)
            rc = run_string_expval<double>(ctx, mk, *static_cast<std::complex<double> const *>(coeff), sout.dev, sin.dev,
                                           dim, n_states, accumulate);
        else
            rc = run_string_expval<float>(ctx, mk, *static_cast<std::complex<float> const *>(coeff), sout.dev, sin.dev,
                                          dim, n_states, accumulate);
        FP_TRY(rc);
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged || sin.zero_copy);
    }

    // ------------------------------------------------------------ PauliOp
    int fp_op_create(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes, const void *coeffs,
                     fp_op **op)
    {
        if (!ctx || !op || (n_strings && n_qubits > 0 && !codes) || (n_strings && !coeffs))
            return set_err(FP_INVALID_ARGUMENT, "null pointer");
        FP_TRY(check_dtype(dtype));
        DeviceGuard g(ctx->device);
        if (dtype == FP_C128)
            return op_create_t<double>(ctx, dtype, n_qubits, n_strings, codes,
                                       static_cast<std::complex<double> const *>(coeffs), true, op);
        return op_create_t<float>(ctx, dtype, n_qubits, n_strings, codes,
                                  static_cast<std::complex<float> const *>(coeffs), true, op);
    }

    int fp_op_destroy(fp_op *op)
    {
        if (!op)
            return FP_OK;
        DeviceGuard g(op->device);
        op->f.release();
        op->d.release();
        delete op;
        return FP_OK;
    }

    int fp_op_info(const fp_op *op, int *dtype, int *n_qubits, size_t *n_strings, size_t *n_packed, size_t *n_groups)
    {
        if (!op)
            return set_err(FP_INVALID_ARGUMENT, "null operator");
        if (dtype)
            *dtype = op->dtype;
        if (n_qubits)
            *n_qubits = op->n_qubits;
        if (n_strings)
            *n_strings = op->n_strings;
        size_t S = op->dtype == FP_C128 ? op->d.host.sz.size() : op->f.host.sz.size();
        size_t G = op->dtype == FP_C128 ? op->d.host.gx.size() : op->f.host.gx.size();
        if (n_packed)
            *n_packed = S;
        if (n_groups)
            *n_groups = G;
        return FP_OK;
    }

    int fp_op_apply(fp_ctx *ctx, const fp_op *op, void *out, const void *in, size_t dim, size_t n_states, int accumulate)
    {
        FP_TRY(op_check(ctx, op));
        uint64_t const opdim = op->n_strings ? dim_of(op->n_qubits) : 0; // PO:105-115
        if (dim != opdim)                                              // PO:343-346
            return set_err(FP_INVALID_ARGUMENT, "[PauliOp] state size must match the dimension of the operators");
        if (dim == 0 || n_states == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        size_t const bytes = dim * n_states * csize(op->dtype);
        if (ctx->pipeline && !accumulate && bytes >= ctx->pipeline_min_bytes && !is_device_ptr(in) && !is_device_ptr(out))
        {
            auto const *ib = static_cast<unsigned char const *>(in);
            auto const *ob = static_cast<unsigned char const *>(out);
            if (ib + bytes <= ob || ob + bytes <= ib) // disjoint host buffers
            {
                bool used = false;
                if (op->dtype == FP_C128)
                    FP_TRY(pipelined_op_apply<double>(ctx, op->d, out, in, dim, n_states, op->n_qubits, &used));
                else
                    FP_TRY(pipelined_op_apply<float>(ctx, op->f, out, in, dim, n_states, op->n_qubits, &used));
                if (used)
                    return finish(ctx, true);
            }
        }
        Staged sin, sout;
        FP_TRY(stage_in(ctx, ctx->stage_in, in, bytes, true, sin));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, bytes, accumulate != 0, sout));
        if (op->dtype == FP_C128)
            FP_TRY(run_op_apply<double>(ctx, op->d, sout.dev, sin.dev, dim, n_states, accumulate, op->n_qubits));
        else
            FP_TRY(run_op_apply<float>(ctx, op->f, sout.dev, sin.dev, dim, n_states, accumulate, op->n_qubits));
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged);
    }

    int fp_op_expval_bra(fp_ctx *ctx, const fp_op *op, void *out, const void *bra, const void *in, size_t dim,
                         size_t n_states, int accumulate)
    {
        FP_TRY(op_check(ctx, op));
        uint64_t const opdim = op->n_strings ? dim_of(op->n_qubits) : 0;
        if (dim != opdim)
            return set_err(FP_INVALID_ARGUMENT, "[PauliOp] state size must match the dimension of the operators");
        if (n_states == 0)
            return FP_OK;
        if (!is_device_ptr(bra) || !is_device_ptr(in) || !is_device_ptr(out))
            return set_err(FP_INVALID_ARGUMENT, "fp_op_expval_bra takes device pointers only");
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (op->dtype == FP_C128)
            FP_TRY(run_op_expval<double>(ctx, op->d, out, in, dim, n_states, accumulate, op->n_qubits, bra));
        else
            FP_TRY(run_op_expval<float>(ctx, op->f, out, in, dim, n_states, accumulate, op->n_qubits, bra));
        return finish(ctx, false);
    }

    int fp_op_expval(fp_ctx *ctx, const fp_op *op, void *out, const void *in, size_t dim, size_t n_states,
                     int accumulate)
    {
        FP_TRY(op_check(ctx, op));
        uint64_t const opdim = op->n_strings ? dim_of(op->n_qubits) : 0;
        if (dim != opdim) // PO:502-505
            return set_err(FP_INVALID_ARGUMENT, "[PauliOp] state size must match the dimension of the operators");
        if (n_states == 0)
            return FP_OK;
        DeviceGuard g(ctx->device);
        std::lock_guard<std::mutex> lk(ctx->mu);
        Staged sin, sout;
        FP_TRY(stage_in(ctx, ctx->stage_in, in, dim * n_states * csize(op->dtype), true, sin));
        FP_TRY(stage_in(ctx, ctx->stage_out, out, n_states * csize(op->dtype), accumulate != 0, sout));
        if (op->dtype == FP_C128)
            FP_TRY(run_op_expval<double>(ctx, op->d, sout.dev, sin.dev, dim, n_states, accumulate, op->n_qubits));
        else
            FP_TRY(run_op_expval<float>(ctx, op->f, sout.dev, sin.dev, dim, n_states, accumulate, op->n_qubits));
        FP_TRY(stage_back(ctx, sout));
        return finish(ctx, sin.staged || sout.staged);
    }

} // extern "C"
