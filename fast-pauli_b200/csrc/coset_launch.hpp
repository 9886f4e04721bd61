// Host side of the coset-blocked kernel family (K3b / K6 coset.cuh, K3e / K3f / K3g coset2.cuh): plan cache, tile-shape
// cost model, launchers and the per-pass dispatch.  Templates in an anonymous namespace: a translation unit only
// instantiates (and compiles kernels for) the modes it calls -- capi.cu MODE 0 / 1, sop_capi.cu MODE 2.
#pragma once
#include <cuda.h>
#include "capi_internal.hpp"

namespace
{
// ---------------------------------------------------------------- coset-blocked path: plan cache + launch
template <typename T>
int get_coset_plan(DeviceOp<T> const &op, int n_qubits, int rank, int reserve_low_bits,
                   std::vector<typename DeviceOp<T>::CosetPassDev> const **out)
{
    int const key = rank * 8 + reserve_low_bits;
    auto it = op.coset_plans.find(key);
    if (it != op.coset_plans.end())
    {
        *out = &it->second;
        return FP_OK;
    }
    std::vector<CosetPassHost<T>> host = plan_coset<T>(op.host, n_qubits, rank, reserve_low_bits);
    std::vector<typename DeviceOp<T>::CosetPassDev> dev(host.size());
    for (size_t p = 0; p < host.size(); ++p)
    {
        CosetPassHost<T> const &h = host[p];
        auto &d = dev[p];
        for (int k = 0; k < kCosetMaxRank; ++k)
            d.view.basis[k] = k < h.basis.r ? h.basis.b[k] : 0;
        d.view.nonpivot_mask = h.nonpivot_mask;
        d.view.n_chunks = static_cast<uint32_t>(h.chunks.size());
        d.view.n_groups = static_cast<uint32_t>(h.gxl.size());
        if (h.gxl.size() <= 8 && h.sz.size() <= kFewParamStrings && !h.gxl.empty())
        {
            d.few = std::make_shared<FewStrings<T>>();
            std::memset(d.few.get(), 0, sizeof(FewStrings<T>));
            for (size_t i = 0; i < h.sz.size(); ++i)
            {
                d.few->c[i] = Cx<T>{h.sc[i].real(), h.sc[i].imag()};
                d.few->z[i] = h.sz[i];
            }
            for (size_t g = 0; g <= h.gxl.size(); ++g)
                d.few->gs[g] = h.gstart[g];
            for (size_t g = 0; g < h.gxl.size(); ++g)
                d.few->gxl[g] = h.gxl[g];
        }
        if (rank == 8 && n_qubits <= 30 && h.gxl.size() > 8 && h.gxl.size() <= kGenMaxGroups && h.sz.size() <= kGenMaxStrings)
        {
            d.gen = std::make_shared<GenStrings<T>>();
            std::memset(d.gen.get(), 0, sizeof(GenStrings<T>));
            for (size_t i = 0; i < h.sz.size(); ++i)
            {
                d.gen->c[i] = Cx<T>{h.sc[i].real(), h.sc[i].imag()};
                d.gen->z[i] = static_cast<uint32_t>(h.sz[i]);
            }
            for (size_t g = 0; g <= h.gxl.size(); ++g)
                d.gen->gs[g] = static_cast<uint16_t>(h.gstart[g]);
            for (size_t g = 0; g < h.gxl.size(); ++g)
                d.gen->gxl[g] = static_cast<uint8_t>(h.gxl[g]);
        }
        if (rank == 8 && n_qubits <= 30 && !h.gxl.empty() && h.gxl.size() == h.sz.size() &&
            h.gxl.size() <= static_cast<size_t>(kDirMaxMasks))
        {
            // one string per x-mask: the row factor is +-c_g (K3i)
            d.dir = std::make_shared<DirStrings<T>>();
            std::memset(d.dir.get(), 0, sizeof(DirStrings<T>));
            d.dir->n = static_cast<uint32_t>(h.gxl.size());
            for (size_t g = 0; g < h.gxl.size(); ++g)
            {
                d.dir->c[g] = Cx<T>{h.sc[g].real(), h.sc[g].imag()};
                d.dir->z[g] = h.sz[g];
                d.dir->xl[g] = h.gxl[g];
                d.dir->zl[g] = h.szl[g];
            }
        }
        uint64_t nb3[8], nb2[8];
        if (rank == 8 && n_qubits <= 30 && h.sz.size() <= static_cast<size_t>(kPairMaxStrings) && pair_basis<T>(h, nb3, nb2))
        {
            // eight independent x-masks: K3j re-chooses the local coordinates from them (coset_plan.hpp: pair_basis)
            d.pair = std::make_shared<PairStrings<T>>();
            std::memset(d.pair.get(), 0, sizeof(PairStrings<T>));
            for (int k = 0; k < kPairMasks; ++k)
            {
                d.pair->basis[0][k] = nb3[k];
                d.pair->basis[1][k] = nb2[k];
            }
            for (size_t i = 0; i < h.sz.size(); ++i)
            {
                d.pair->c[i] = Cx<T>{h.sc[i].real(), h.sc[i].imag()};
                d.pair->z[i] = static_cast<uint32_t>(h.sz[i]);
            }
            for (size_t g = 0; g <= h.gxl.size(); ++g)
                d.pair->gs[g] = static_cast<uint8_t>(h.gstart[g]);
        }
        CosetChunk *chunks = nullptr;
        uint32_t *gxl = nullptr, *gstart = nullptr, *szl = nullptr, *sidx = nullptr;
        uint64_t *sz = nullptr;
        Cx<T> *sc = nullptr;
        std::vector<Cx<T>> scv(h.sc.size());
        for (size_t i = 0; i < scv.size(); ++i)
            scv[i] = Cx<T>{h.sc[i].real(), h.sc[i].imag()};
        int rc = upload_vec(&chunks, h.chunks);
        if (rc == FP_OK) { d.allocs.push_back(chunks); rc = upload_vec(&gxl, h.gxl); }
        if (rc == FP_OK) { d.allocs.push_back(gxl); rc = upload_vec(&gstart, h.gstart); }
        if (rc == FP_OK) { d.allocs.push_back(gstart); rc = upload_vec(&szl, h.szl); }
        if (rc == FP_OK) { d.allocs.push_back(szl); rc = upload_vec(&sz, h.sz); }
        if (rc == FP_OK) { d.allocs.push_back(sz); rc = upload_vec(&sc, scv); }
        if (rc == FP_OK) { d.allocs.push_back(sc); rc = upload_vec(&sidx, h.sidx); }
        if (rc == FP_OK) d.allocs.push_back(sidx);
        if (rc == FP_OK)
        {
            std::vector<PairChunk> ech;
            std::vector<uint8_t> esodd(h.sidx.size());
            for (size_t i = 0; i < h.sidx.size(); ++i)
                esodd[i] = op.host.sodd[h.sidx[i]];
            for (size_t g = 0; g + 1 < h.gstart.size(); ++g)
            {
                uint32_t const xl = h.gxl[g];
                for (uint32_t s0 = h.gstart[g]; s0 < h.gstart[g + 1]; s0 += kPairMS)
                {
                    PairChunk c;
                    c.x = xl;
                    c.s0 = s0;
                    c.count = std::min<uint32_t>(kPairMS, h.gstart[g + 1] - s0);
                    c.hbit = xl ? 31u - static_cast<uint32_t>(__builtin_clz(xl)) : 0u;
                    c.diag = xl == 0;
                    ech.push_back(c);
                }
            }
            PairChunk *d_ech = nullptr;
            uint8_t *d_esodd = nullptr;
            rc = upload_vec(&d_ech, ech);
            if (rc == FP_OK) { d.allocs.push_back(d_ech); rc = upload_vec(&d_esodd, esodd); }
            if (rc == FP_OK) d.allocs.push_back(d_esodd);
            d.echunks = d_ech;
            d.n_echunks = static_cast<uint32_t>(ech.size());
            d.esodd = d_esodd;
        }
        if (rc != FP_OK)
        {
            for (auto &dd : dev)
                for (void *a : dd.allocs)
                    cudaFree(a);
            return rc;
        }
        d.view.chunks = chunks;
        d.view.gxl = gxl;
        d.view.gstart = gstart;
        d.view.szl = szl;
        d.view.sz = sz;
        d.view.scoef = sc;
        d.view.sidx = sidx;
    }
    auto ins = op.coset_plans.emplace(key, std::move(dev));
    *out = &ins.first->second;
    return FP_OK;
}

// ---------------------------------------------------------------- coset-blocked path: heuristics + launch
struct CosetShape
{
    int log_twc = -1; // TWc = 2^log_twc vectors per row segment
    int log_nt = 8;   // threads per CTA
    int vpt = 16;     // vectors per thread (tile = vpt * NT vectors)
    int rank() const
    {
        return (vpt == 16 ? 4 : 3) + log_nt - log_twc;
    }
    bool ok() const
    {
        return log_twc >= 0;
    }
};

// Pick the tile shape, or an invalid shape for "use the generic gather kernel".
// tma_kernels: the call can use the TMA-fed rank-8 kernels (K3f / K3g: apply on device-resident batches with rows of
// >= 256 bytes), whose passes are cheaper than the general kernel's
template <typename T>
CosetShape choose_coset(fp_ctx const *ctx, DeviceOp<T> const &op, int n_qubits, uint64_t rowvecs, int epv,
                        bool tma_kernels = false, bool pair_expval = false)
{
    CosetShape none;
    if (ctx->coset_mode == 0 || op.host.sz.size() < 2)
        return none;
    if (n_qubits > 12 && op.host.gx.size() > 20000)
        return none; // pass planning is quadratic in the number of x-groups: huge operators use the generic kernel
    if (sizeof(T) == 4 && epv != 2)
        return none;
    auto valid = [&](int v, int lnt) {
        return v >= 0 && v <= 4 && (lnt == 7 || lnt == 8) && (rowvecs % (1ull << v)) == 0 && (4 + lnt - v) <= n_qubits;
    };
    CosetShape pick;
    if (ctx->coset_log_twc >= 0)
    {
        int lnt = ctx->coset_log_nt > 0 ? ctx->coset_log_nt : 8;
        int vpt = ctx->coset_vpt == 8 ? 8 : 16;
        bool ok = vpt == 16 ? valid(ctx->coset_log_twc, lnt)
                            : (lnt == 8 && ctx->coset_log_twc <= 3 && (rowvecs % (1ull << ctx->coset_log_twc)) == 0 &&
                               (3 + lnt - ctx->coset_log_twc) <= n_qubits);
        if (ok)
        {
            pick.log_twc = ctx->coset_log_twc;
            pick.log_nt = lnt;
            pick.vpt = vpt;
        }
    }
    else
    {
        int const lnt_pref = ctx->coset_log_nt > 0 ? ctx->coset_log_nt : 8;
        // the whole state column fits one tile: single pass whatever the operator
        if (n_qubits <= 12 && valid(12 - n_qubits, 8))
        {
            pick.log_twc = 12 - n_qubits;
            pick.log_nt = 8;
        }
        // otherwise: the candidate whose (number of passes) x (relative cost of a pass at that row-segment width)
        // is smallest; pass counts come from the real planner (plans are cached on the operator)
        static double const seg_cost[5] = {3.5, 2.0, 1.45, 1.05, 1.0}; // measured, HBM-bound passes, v = 0..4
        double best = 0;
        for (int v = 4; v >= 0 && !pick.ok(); --v)
        {
            if (!valid(v, lnt_pref))
                continue;
            std::vector<typename DeviceOp<T>::CosetPassDev> const *passes = nullptr;
            int const reserve = std::max(0, 2 - v);
            if (get_coset_plan<T>(op, n_qubits, 4 + lnt_pref - v, reserve, &passes) != FP_OK)
                continue;
            // measured (20 q x 64 chains, 16 q x 1024 config 3): a rank-8 pass of K3e / K3f / K3g costs ~0.7 of a
            // general-kernel pass of the same width
            double cost = static_cast<double>(passes->size()) * seg_cost[v] *
                          ((tma_kernels && v == 4 && lnt_pref == 8) ? 0.7 : 1.0);
            if (pair_expval && v == 4 && lnt_pref == 8)
            {
                // expectation values: only K3j has a TMA-fed reduction mode; its passes cost ~0.6 of a general one
                // (measured at 20 q x 64: 0.38 against 0.6 ms)
                cost = 0;
                for (auto const &pd : *passes)
                    cost += seg_cost[v] * (pd.pair ? 0.6 : 1.0);
            }
            if (best == 0 || cost < best)
            {
                best = cost;
                none.log_twc = v; // remember the best so far in `none` (returned through `pick` below)
                none.log_nt = lnt_pref;
            }
        }
        if (!pick.ok() && none.ok())
            pick = none;
        none = CosetShape{};
    }
    if (!pick.ok())
        return none;
    uint64_t const ctas = (1ull << (n_qubits - pick.rank())) * (rowvecs >> pick.log_twc);
    if (ctx->coset_mode == 1 && ctas < static_cast<uint64_t>(ctx->sm_count))
        return none;
    return pick;
}

template <typename T, int EPV, int LOG_TWC, int LOG_NT, int MODE, int VPT = 16>
int launch_coset_pass(fp_ctx *ctx, CosetPassView<T> const &view, int n_qubits, uint64_t rowvecs, void const *in,
                      void *out, int beta, void *partials, uint32_t Bpad, T const *Wre, T const *Wim, uint64_t B)
{
    using Cfg = CosetCfg<LOG_TWC, LOG_NT, VPT>;
    size_t const smem = coset_smem_bytes<T, LOG_TWC, LOG_NT, VPT>();
    static PerDevice configured; // per template instance
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(coset_kernel<T, EPV, LOG_TWC, LOG_NT, MODE, VPT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.set(ctx->device);
    }
    uint32_t const nct = static_cast<uint32_t>(rowvecs >> LOG_TWC);
    uint64_t const grid = (1ull << (n_qubits - Cfg::R)) * nct;
    FP_TRY(check_grid(grid));
    coset_kernel<T, EPV, LOG_TWC, LOG_NT, MODE, VPT><<<static_cast<unsigned>(grid), Cfg::NT, smem, ctx->stream>>>(
        view, rowvecs, nct, static_cast<CVec<T, EPV> const *>(in), static_cast<CVec<T, EPV> *>(out), beta,
        static_cast<Cx<T> *>(partials), Bpad, Wre, Wim, B);
    ctx->launches++;
    ctx->coset_kernels |= 1u;
    return FP_OK;
}

template <typename T, int EPV, int MODE>
int launch_coset_pass_v(fp_ctx *ctx, CosetShape shape, CosetPassView<T> const &view, int n_qubits, uint64_t rowvecs,
                        void const *in, void *out, int beta, void *partials, uint32_t Bpad, T const *Wre, T const *Wim,
                        uint64_t B)
{
#define FP_COSET_CASE(V, LNT)                                                                                          \
    if (shape.vpt == 16 && shape.log_twc == V && shape.log_nt == LNT)                                                  \
        return launch_coset_pass<T, EPV, V, LNT, MODE>(ctx, view, n_qubits, rowvecs, in, out, beta, partials, Bpad,    \
                                                       Wre, Wim, B);
    if constexpr (MODE == 2)
    {
        // weighted apply on a whole-column tile (rank 12, one vector per row): 512 threads x 8 rows halve the
        // per-thread accumulator + D registers, so 16 warps are resident per SM instead of 8
        if (shape.vpt == 16 && shape.log_twc == 0 && shape.log_nt == 8 && ctx->coset_wide_cta)
            return launch_coset_pass<T, EPV, 0, 9, MODE, 8>(ctx, view, n_qubits, rowvecs, in, out, beta, partials, Bpad,
                                                           Wre, Wim, B);
    }
#define FP_COSET_CASE8(V)                                                                                              \
    if (shape.vpt == 8 && shape.log_twc == V && shape.log_nt == 8)                                                     \
        return launch_coset_pass<T, EPV, V, 8, MODE, 8>(ctx, view, n_qubits, rowvecs, in, out, beta, partials, Bpad,   \
                                                        Wre, Wim, B);
    if constexpr (MODE != 2)
    {
        FP_COSET_CASE8(0)
        FP_COSET_CASE8(1)
        FP_COSET_CASE8(2)
        FP_COSET_CASE8(3)
    }
#undef FP_COSET_CASE8
    FP_COSET_CASE(0, 8)
    FP_COSET_CASE(1, 8)
    FP_COSET_CASE(2, 8)
    FP_COSET_CASE(3, 8)
    FP_COSET_CASE(4, 8)
    FP_COSET_CASE(0, 7)
    FP_COSET_CASE(1, 7)
    FP_COSET_CASE(2, 7)
    FP_COSET_CASE(3, 7)
    FP_COSET_CASE(4, 7)
#undef FP_COSET_CASE
    return set_err(FP_UNSUPPORTED, "unsupported coset tile shape");
}

// ---------------------------------------------------------------- K3e / K3f (coset2.cuh): passes with <= 8 x-masks
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, cuuint64_t const *,
                                      cuuint64_t const *, cuuint32_t const *, cuuint32_t const *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda); nullptr when unavailable
TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
        {
            (void)cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}

// The batch as a 2-D tensor (rows = dim, inner = real scalars of one row) with a box of one 256-byte row segment:
// the shape TMA tile::gather4 wants (four arbitrary rows per operation).
template <typename T> bool make_row_tensor_map(CUtensorMap *tm, void const *base, uint64_t dim, uint64_t rowvecs)
{
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc)
        return false;
    cuuint64_t dims[2] = {rowvecs * (16 / sizeof(T)), dim};
    cuuint64_t strides[1] = {rowvecs * 16};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(256 / sizeof(T)), 1};
    cuuint32_t es[2] = {1, 1};
    return enc(tm, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
               const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int EPV, int LOG_TWC, int NBUF, bool PSTR, int MODE = 0, bool RMWPF = false>
int launch_coset_few_v(fp_ctx *ctx, CosetPassView<T> const &view, FewStrings<T> const &strs, int n_qubits,
                       uint64_t rowvecs, void const *in, void *out, int beta, void *partials = nullptr,
                       uint32_t Bpad = 0)
{
    using Cfg = FewCfg<LOG_TWC>;
    constexpr int GMAX = 8;
    if constexpr (MODE == 0 && !RMWPF)
    {
        // accumulating passes prefetch the old output rows through one more shared-memory buffer
        if (beta)
            return launch_coset_few_v<T, EPV, LOG_TWC, NBUF, PSTR, MODE, true>(ctx, view, strs, n_qubits, rowvecs, in, out,
                                                                               beta, partials, Bpad);
    }
    constexpr size_t smem = (NBUF + (RMWPF ? 1 : 0)) * Cfg::TILE_BYTES;
    static PerDevice configured; // per template instance
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(coset_few_kernel<T, EPV, LOG_TWC, GMAX, NBUF, 2, PSTR, RMWPF, MODE>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.set(ctx->device);
    }
    uint32_t const nct = static_cast<uint32_t>(rowvecs >> LOG_TWC);
    uint64_t const n_cosets = 1ull << (n_qubits - Cfg::R);
    // column tiles per CTA: as many as possible (the row factors are formed once per CTA) while >= 4 waves remain
    uint32_t per = nct;
    if (ctx->coset_few_ct > 0)
        per = std::min<uint32_t>(nct, static_cast<uint32_t>(ctx->coset_few_ct));
    else
        while (per > 1 && n_cosets * ((nct + per - 1) / per) < 8ull * static_cast<uint64_t>(ctx->sm_count))
            per = (per + 1) / 2;
    uint32_t const groups = (nct + per - 1) / per;
    uint64_t const grid = n_cosets * groups;
    FP_TRY(check_grid(grid));
    coset_few_kernel<T, EPV, LOG_TWC, GMAX, NBUF, 2, PSTR, RMWPF, MODE>
        <<<static_cast<unsigned>(grid), Cfg::NT, smem, ctx->stream>>>(
            view, rowvecs, nct, per, groups, static_cast<CVec<T, EPV> const *>(in), static_cast<CVec<T, EPV> *>(out), beta,
            strs, static_cast<Cx<T> *>(partials), Bpad);
    ctx->launches++;
    ctx->coset_kernels |= 2u;
    return FP_OK;
}

// K3f: persistent TMA-fed kernel, overwrite or accumulate, 12..30 qubits, rows of >= 256 bytes
template <typename T, int EPV>
int launch_coset_few_tma(fp_ctx *ctx, CosetPassView<T> const &view, FewStrings<T> const &strs, int n_qubits,
                         uint64_t rowvecs, void const *in, void *out, int beta, bool *launched)
{
    *launched = false;
    CUtensorMap tm;
    if (!make_row_tensor_map<T>(&tm, in, 1ull << n_qubits, rowvecs))
        return FP_OK;
    constexpr size_t smem = kFewTmaBufs * kFewTmaTile;
    static PerDevice configured;
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(coset_few_tma_kernel<T, EPV, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
        configured.set(ctx->device);
    }
    uint64_t const n_pairs = 1ull << (n_qubits - 9);
    unsigned const grid = static_cast<unsigned>(std::min<uint64_t>(n_pairs, static_cast<uint64_t>(ctx->sm_count)));
    coset_few_tma_kernel<T, EPV, 8><<<grid, kFewTmaThreads, smem, ctx->stream>>>(
        view, rowvecs, static_cast<uint32_t>(rowvecs >> 4), n_pairs, static_cast<CVec<T, EPV> *>(out), beta, strs, tm);
    ctx->launches++;
    ctx->coset_kernels |= 4u;
    *launched = true;
    return FP_OK;
}

// K3g: persistent TMA-fed kernel for passes with more than 8 x-masks (row factors per tile from the constant bank)
template <typename T, int EPV>
int launch_coset_gen_tma(fp_ctx *ctx, CosetPassView<T> const &view, GenStrings<T> const &gstr, int n_qubits,
                         uint64_t rowvecs, void const *in, void *out, int beta, bool *launched)
{
    *launched = false;
    CUtensorMap tm;
    if (!make_row_tensor_map<T>(&tm, in, 1ull << n_qubits, rowvecs))
        return FP_OK;
    constexpr size_t smem = kFewTmaBufs * kFewTmaTile + kGenMetaBytes;
    static PerDevice configured;
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(coset_gen_tma_kernel<T, EPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
        configured.set(ctx->device);
    }
    uint64_t const n_pairs = 1ull << (n_qubits - 9);
    uint32_t const nct = static_cast<uint32_t>(rowvecs >> 4);
    // work items = coset pairs x chunks of column tiles: enough of them to balance 148 persistent CTAs
    uint32_t chunk = nct;
    while (chunk > 1 && chunk % 2 == 0 && n_pairs * (nct / chunk) < 6ull * static_cast<uint64_t>(ctx->sm_count))
        chunk /= 2;
    uint64_t const items = n_pairs * (nct / chunk);
    unsigned const grid = static_cast<unsigned>(std::min<uint64_t>(items, static_cast<uint64_t>(ctx->sm_count)));
    coset_gen_tma_kernel<T, EPV, true><<<grid, kFewTmaThreads, smem, ctx->stream>>>(
        view, rowvecs, nct, n_pairs, chunk, static_cast<CVec<T, EPV> *>(out), beta, tm, gstr);
    ctx->launches++;
    ctx->coset_kernels |= 8u;
    *launched = true;
    return FP_OK;
}

// K3i: persistent TMA-fed kernel with direct stores for passes whose x-masks carry one string each (coset3.cuh)
template <typename T, int EPV>
int launch_coset_dir_tma(fp_ctx *ctx, CosetPassView<T> const &view, DirStrings<T> const &strs, int n_qubits,
                         uint64_t rowvecs, void const *in, void *out, int beta, bool *launched)
{
    *launched = false;
    CUtensorMap tm;
    if (!make_row_tensor_map<T>(&tm, in, 1ull << n_qubits, rowvecs))
        return FP_OK;
    constexpr size_t smem = kFewTmaBufs * kFewTmaTile;
    uint32_t const nct = static_cast<uint32_t>(rowvecs >> 4);
    uint64_t const n_tiles = (1ull << (n_qubits - 8)) * nct;
    unsigned const grid = static_cast<unsigned>(std::min<uint64_t>(n_tiles, static_cast<uint64_t>(ctx->sm_count)));
#define FP_LAUNCH_DIR(NCH)                                                                                             \
    {                                                                                                                  \
        static PerDevice configured;                                                                                   \
        if (!configured.done(ctx->device))                                                                             \
        {                                                                                                              \
            FP_CU(cudaFuncSetAttribute(coset_dir_tma_kernel<T, EPV, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                       static_cast<int>(smem)));                                                       \
            configured.set(ctx->device);                                                                               \
        }                                                                                                              \
        coset_dir_tma_kernel<T, EPV, NCH><<<grid, kDirThreads, smem, ctx->stream>>>(                                   \
            view, rowvecs, nct, n_tiles, static_cast<CVec<T, EPV> *>(out), beta, strs, tm);                            \
    }
    if (strs.n <= 8)
        FP_LAUNCH_DIR(1)
    else if (strs.n <= 16)
        FP_LAUNCH_DIR(2)
    else
        FP_LAUNCH_DIR(4)
#undef FP_LAUNCH_DIR
    ctx->launches++;
    ctx->coset_kernels |= 16u;
    *launched = true;
    return FP_OK;
}

// K3j: persistent TMA-fed kernel with direct stores, paired masks and a row-factor table, for passes of eight
// independent x-masks with any number of strings each (coset4.cuh)
// MODE 1: expectation-value partials, one row per (CTA, consumer warp), zeroed here and folded by the caller
template <typename T, int EPV, int MODE = 0>
int launch_coset_pair_tma(fp_ctx *ctx, CosetPassView<T> const &view, PairStrings<T> const &strs, int n_qubits,
                          uint64_t rowvecs, void const *in, void *out, int beta, bool *launched, void *partials = nullptr,
                          uint32_t Bpad = 0, uint64_t *n_partial_rows = nullptr)
{
    *launched = false;
    CUtensorMap tm;
    if (!make_row_tensor_map<T>(&tm, in, 1ull << n_qubits, rowvecs))
        return FP_OK;
    constexpr size_t smem = kFewTmaBufs * kFewTmaTile + pair_table_bytes<T>();
    uint32_t const nct = static_cast<uint32_t>(rowvecs >> 4);
    uint64_t const n_tiles = (1ull << (n_qubits - 8)) * nct;
    unsigned const grid = static_cast<unsigned>(std::min<uint64_t>(n_tiles, static_cast<uint64_t>(ctx->sm_count)));
    if (MODE == 1)
    {
        uint64_t const rows = static_cast<uint64_t>(grid) * kPairConsumerWarps;
        FP_CU(cudaMemsetAsync(partials, 0, rows * Bpad * sizeof(Cx<T>), ctx->stream));
        *n_partial_rows = rows;
    }
#define FP_LAUNCH_PAIR(RB)                                                                                             \
    {                                                                                                                  \
        static PerDevice configured;                                                                                   \
        if (!configured.done(ctx->device))                                                                             \
        {                                                                                                              \
            FP_CU(cudaFuncSetAttribute(coset_pair_tma_kernel<T, EPV, RB, MODE>,                                        \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));          \
            configured.set(ctx->device);                                                                               \
        }                                                                                                              \
        coset_pair_tma_kernel<T, EPV, RB, MODE><<<grid, kPairThreads, smem, ctx->stream>>>(                            \
            view.nonpivot_mask, rowvecs, nct, n_tiles, MODE == 1 ? nullptr : static_cast<CVec<T, EPV> *>(out), beta,   \
            strs, tm, static_cast<Cx<T> *>(partials), Bpad);                                                           \
    }
    static int const rb = getenv("FASTPAULI_PAIR_RB") ? atoi(getenv("FASTPAULI_PAIR_RB")) : 2;
    if (rb == 3)
        FP_LAUNCH_PAIR(3)
    else
        FP_LAUNCH_PAIR(2)
#undef FP_LAUNCH_PAIR
    ctx->launches++;
    ctx->coset_kernels |= 32u;
    *launched = true;
    return FP_OK;
}

// PauliOp::apply on ONE state (PO:362-383 through PO:399-468 with one column; the local piece of the sharded 34-qubit
// state, BASELINE config 5).  A row of a one-column batch is 16 bytes, so the coset kernels gather single vectors and
// form a row factor per vector (measured: 2.1 TB/s algorithmic at 28 qubits).  Here the state is VIEWED as 2^(n-4) rows x
// 16 columns -- the column is the 4 lowest index bits j -- so rows are 256-byte segments again; a string then also
// permutes the columns (j -> j ^ (x & 15)) and signs them ((-1)^popc(j & z & 15)), which K3i takes as one more XOR in the
// gather address and one more parity in the lane's sign word.  The operator is re-planned on the upper n-4 bits (strings
// that differ only in the low bits share a row offset; K3i treats every string as its own mask).
template <typename T>
int try_single_state(fp_ctx *ctx, DeviceOp<T> const &op, int n_qubits, void *out, void const *in, int beta, bool *used)
{
    *used = false;
    if constexpr (sizeof(T) != 8)
        return FP_OK;
    else
    {
        int const nr = n_qubits - 4;
        if (ctx->coset_few != 1 || ctx->coset_mode == 0 || n_qubits < 22 || nr > 30 || op.single_state_tried < 0 ||
            !is_device_ptr(in) || tensor_map_encoder() == nullptr || op.host.sz.size() < 2)
            return FP_OK;
        if (op.single_state_tried == 0)
        {
            op.single_state_tried = -1;
            // the operator on the upper bits with the low nibbles kept per string (coset_plan.hpp: single_state_reshape)
            SingleStateOp<T> const ss = single_state_reshape<T>(op.host, n_qubits);
            PackedOp<T> const &r = ss.r;
            std::vector<CosetPassHost<T>> host = plan_coset<T>(r, nr, 8, 0);
            std::vector<typename DeviceOp<T>::SingleStatePass> plan(host.size());
            // every pass re-streams the state: only worth it while the passes stay few (the general path needs one
            // pass per 10..12 independent masks at a third of the speed)
            bool ok = !host.empty() && host.size() <= 8;
            for (size_t p = 0; ok && p < host.size(); ++p)
            {
                CosetPassHost<T> const &h = host[p];
                if (h.sz.size() > static_cast<size_t>(kDirMaxMasks) || h.sz.empty())
                {
                    ok = false;
                    break;
                }
                auto &d = plan[p];
                for (int k = 0; k < kCosetMaxRank; ++k)
                    d.view.basis[k] = k < h.basis.r ? h.basis.b[k] : 0;
                d.view.nonpivot_mask = h.nonpivot_mask;
                std::memset(&d.strs, 0, sizeof(d.strs));
                d.strs.n = static_cast<uint32_t>(h.sz.size());
                for (size_t g = 0; g < h.gxl.size(); ++g)
                    for (uint32_t t = h.gstart[g]; t < h.gstart[g + 1]; ++t)
                    {
                        d.strs.c[t] = Cx<T>{h.sc[t].real(), h.sc[t].imag()};
                        d.strs.z[t] = h.sz[t];
                        d.strs.xl[t] = h.gxl[g];
                        d.strs.zl[t] = h.szl[t];
                        d.strs.xlo[t] = ss.xlo[h.sidx[t]];
                        d.strs.zlo[t] = ss.zlo[h.sidx[t]];
                    }
            }
            if (ok)
            {
                op.single_state_plan = std::move(plan);
                op.single_state_tried = 1;
            }
        }
        if (op.single_state_tried != 1)
            return FP_OK;
        for (size_t p = 0; p < op.single_state_plan.size(); ++p)
        {
            bool launched = false;
            auto const &d = op.single_state_plan[p];
            FP_TRY((launch_coset_dir_tma<T, 1>(ctx, d.view, d.strs, nr, 16, in, out, p == 0 ? beta : 1, &launched)));
            if (!launched)
            {
                if (p == 0)
                    return FP_OK; // no tensor map: the caller's general path
                return set_err(FP_CUDA_ERROR, "single-state PauliOp.apply: tensor map creation failed mid-plan");
            }
        }
        *used = true;
        return FP_OK;
    }
}

// Picks the variant for one pass; *launched = false when the pass has to go through coset_kernel (K3b).
template <typename T, int EPV, int MODE = 0>
int launch_coset_few(fp_ctx *ctx, typename DeviceOp<T>::CosetPassDev const &pd, int n_qubits, uint64_t rowvecs,
                     void const *in, void *out, int beta, bool *launched, void *partials = nullptr, uint32_t Bpad = 0,
                     uint64_t *n_partial_rows = nullptr)
{
    *launched = false;
    CosetPassView<T> const &view = pd.view;
    bool const tma_mode = ctx->coset_few == 1 || ctx->coset_few == 3 || ctx->coset_few == 4;
    // eight independent x-masks: direct stores + paired masks + row-factor table (measured at 20 qubits x 64: 64 strings
    // over 8 masks 0.512 -> 0.416 ms, 8 single-string masks 0.417 -> 0.382 ms); read-modify-write passes of single-string
    // masks stay on K3i, whose early loads of the old output rows suit them slightly better (64 random strings: 4.25
    // against 4.28 ms with K3j on all eight passes)
    // (expectation values -- MODE 1 -- take it for every such pass: 8-mask operator 0.50 -> 0.39 ms, 64 random strings 3.96 -> 3.01 ms)
    if ((MODE == 0 || MODE == 1) && ctx->coset_few == 1 && pd.pair &&
        (MODE == 1 ? n_partial_rows != nullptr : (!pd.dir || beta == 0 || ctx->coset_pair_all)) && n_qubits >= 12 &&
        n_qubits <= 30 && rowvecs % 16 == 0 && is_device_ptr(in) &&
        (rowvecs >> 4) << (n_qubits - 8) >= 4ull * static_cast<uint64_t>(ctx->sm_count))
    {
        if constexpr (MODE == 0 || MODE == 1)
            FP_TRY((launch_coset_pair_tma<T, EPV, MODE>(ctx, view, *pd.pair, n_qubits, rowvecs, in, out, beta, launched,
                                                        partials, Bpad, n_partial_rows)));
        if (*launched)
            return FP_OK;
    }
    // one string per x-mask (random strings): direct-store kernel, overwrite and read-modify-write passes alike
    // (measured at 20 qubits x 64: 8 masks 0.49 -> 0.43 ms, the 8 passes of 64 random strings 4.55 -> 4.16 ms)
    if (MODE == 0 && (ctx->coset_few == 1 || ctx->coset_few == 4) && pd.dir && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
        is_device_ptr(in) && (rowvecs >> 4) << (n_qubits - 8) >= 4ull * static_cast<uint64_t>(ctx->sm_count))
    {
        FP_TRY((launch_coset_dir_tma<T, EPV>(ctx, view, *pd.dir, n_qubits, rowvecs, in, out, beta, launched)));
        if (*launched)
            return FP_OK;
    }
    if (MODE == 0 && tma_mode && pd.gen && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
        is_device_ptr(in))
        return launch_coset_gen_tma<T, EPV>(ctx, view, *pd.gen, n_qubits, rowvecs, in, out, beta, launched);
    if (!ctx->coset_few || view.n_groups == 0 || view.n_groups > 8 || n_qubits < 8)
        return FP_OK;
    static FewStrings<T> const no_strings{};
    bool const pstr = pd.few != nullptr;
    FewStrings<T> const &strs = pstr ? *pd.few : no_strings;
    // K3f wins on overwrite passes of large registers (measured at 20 qubits: 4 masks 0.42 -> 0.38 ms, 256 columns
    // 1.95 -> 1.85 ms; 8 masks equal); read-modify-write passes and small registers stay on the resident-CTA kernel
    if (MODE == 0 && tma_mode && pstr && beta == 0 && n_qubits >= 16 && n_qubits <= 30 && rowvecs % 16 == 0 &&
        is_device_ptr(in))
    {
        FP_TRY((launch_coset_few_tma<T, EPV>(ctx, view, strs, n_qubits, rowvecs, in, out, beta, launched)));
        if (*launched)
            return FP_OK;
    }
    if (rowvecs % 8 == 0 && pstr)
        FP_TRY((launch_coset_few_v<T, EPV, 3, 2, true, MODE>(ctx, view, strs, n_qubits, rowvecs, in, out, beta, partials,
                                                             Bpad)));
    else if (rowvecs % 8 == 0)
        FP_TRY((launch_coset_few_v<T, EPV, 3, 2, false, MODE>(ctx, view, strs, n_qubits, rowvecs, in, out, beta, partials,
                                                              Bpad)));
    else
        return FP_OK;
    *launched = true;
    return FP_OK;
}

// Runs all passes.  Returns FP_OK with *used = false when the generic kernel should be used instead.
template <typename T, int MODE>
int try_coset(fp_ctx *ctx, DeviceOp<T> const &op, int n_qubits, void *out, void const *in, uint64_t dim, uint64_t B,
              int beta, T const *Wre, T const *Wim, bool *used)
{
    *used = false;
    constexpr int EPV = sizeof(T) == 4 ? 2 : 1;
    if (n_qubits <= 0 || dim != (1ull << n_qubits))
        return FP_OK;
    int const epv = pick_epv<T>(in, MODE == 1 ? in : out, B);
    if (epv != EPV)
        return FP_OK;
    uint64_t const rowvecs = B / EPV;
    bool const tma_ok = MODE == 0 && (ctx->coset_few == 1 || ctx->coset_few == 3 || ctx->coset_few == 4) && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
                        is_device_ptr(in) && tensor_map_encoder() != nullptr;
    bool const pair_ok = MODE == 1 && ctx->coset_few == 1 && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
                         is_device_ptr(in) && tensor_map_encoder() != nullptr &&
                         (rowvecs >> 4) << (n_qubits - 8) >= 4ull * static_cast<uint64_t>(ctx->sm_count);
    CosetShape const shape = choose_coset<T>(ctx, op, n_qubits, rowvecs, epv, tma_ok, pair_ok);
    if (!shape.ok())
        return FP_OK;
    std::vector<typename DeviceOp<T>::CosetPassDev> const *passes = nullptr;
    // narrow row segments (16 / 32 bytes): force the lowest row bits into the tile so it is made of >= 64-byte runs
    int const reserve = std::max(0, 2 - shape.log_twc);
    FP_TRY(get_coset_plan<T>(op, n_qubits, shape.rank(), reserve, &passes));
    // every pass re-streams the batch (read in, read-modify-write out): only worth it while passes << groups
    if (ctx->coset_mode == 1 && passes->size() * 3 > op.host.gx.size() && passes->size() > 1)
        return FP_OK;
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    uint64_t const n_cosets = 1ull << (n_qubits - shape.rank());
    if (MODE == 1) // one row per coset, or (K3j) one per consumer warp of a persistent grid
        FP_TRY(ctx->partials.ensure(std::max<uint64_t>(n_cosets, static_cast<uint64_t>(ctx->sm_count) * kPairConsumerWarps) *
                                    Bpad * 2 * sizeof(T)));
    for (size_t p = 0; p < passes->size(); ++p)
    {
        int const b = (p == 0) ? beta : 1;
        if constexpr (MODE == 0 || MODE == 1)
        {
            if (shape.rank() == 8 && shape.log_nt == 8)
            {
                bool launched = false;
                uint64_t partial_rows = n_cosets; // K3j writes one partial row per consumer warp instead of one per coset
                FP_TRY((launch_coset_few<T, EPV, MODE>(ctx, (*passes)[p], n_qubits, rowvecs, in, out, b, &launched,
                                                       ctx->partials.p, Bpad, &partial_rows)));
                if (launched)
                {
                    if (MODE == 1)
                    {
                        unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
                        finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
                            static_cast<Cx<T> const *>(ctx->partials.p), partial_rows, Bpad, B, static_cast<Cx<T> *>(out), b);
                        ctx->launches++;
                    }
                    continue;
                }
            }
        }
        FP_TRY((launch_coset_pass_v<T, EPV, MODE>(ctx, shape, (*passes)[p].view, n_qubits, rowvecs, in, out, b,
                                                   ctx->partials.p, Bpad, Wre, Wim, B)));
        if (MODE == 1)
        {
            unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
            finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
                static_cast<Cx<T> const *>(ctx->partials.p), n_cosets, Bpad, B, static_cast<Cx<T> *>(out), b);
            ctx->launches++;
        }
    }
    *used = true;
    return FP_OK;
}

} // namespace
