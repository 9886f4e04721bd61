// C ABI of the B200-native fast-pauli hot path (declared in include/fastpauli_b200.h).
//
// Host-side responsibilities only: argument checks that mirror the reference's std::invalid_argument sites,
// host/device pointer staging, plan packing (pack.hpp), geometry selection and kernel launches (kernels.cuh).
// No compute happens on the host and there is no CPU fallback: without a CUDA device every compute call fails.
#include "../../include/fastpauli_b200.h"
#include "internal.h"

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "coset.cuh"
#include "coset2.cuh"
#include "rcoset.cuh"
#include "dcoset.cuh"
#include "wtile.cuh"
#include "etile.cuh"
#include "square.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "pack.hpp"

using namespace fpk;

// ================================================================ errors
namespace
{
thread_local std::string g_err;

int set_err(int code, std::string msg)
{
    g_err = std::move(msg);
    return code;
}

#define FP_CU(call)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? FP_OUT_OF_MEMORY                                           \
                        : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? FP_NO_DEVICE                \
                                                                                         : FP_CUDA_ERROR;              \
            return set_err(code_, std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
        }                                                                                                              \
    } while (0)

#define FP_TRY(expr)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc_ = (expr);                                                                                              \
        if (rc_ != FP_OK)                                                                                              \
            return rc_;                                                                                                \
    } while (0)

// cudaFuncSetAttribute is per device: remember per template instance (one static PerDevice each) which devices of
// this process have been configured (a host may hold one context per GPU)
struct PerDevice
{
    uint64_t mask = 0;
    bool done(int device) const
    {
        return (mask >> (device & 63)) & 1ull;
    }
    void set(int device)
    {
        mask |= 1ull << (device & 63);
    }
};

struct Scratch
{
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return FP_OK;
        if (p)
        {
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + (bytes >> 3); // 12.5 % headroom so slowly growing calls do not reallocate every time
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess)
        {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess)
        {
            (void)cudaGetLastError();
            p = nullptr;
            return set_err(FP_OUT_OF_MEMORY, "device scratch allocation of " + std::to_string(bytes) + " bytes failed");
        }
        cap = want;
        return FP_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
} // namespace

// ================================================================ opaque types
struct fp_ctx
{
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    bool async = false;
    bool tensor_core = true;
    bool zero_copy = true; // single-pass kernels read/write pinned host buffers in place
    uint64_t launches = 0;
    int last_gemm_engine = -1; // 0 = SIMT, 1 = tcgen05 (diagnostics)
    size_t l2_budget = 40ull << 20;
    int coset_mode = 1;       // 0: never use the coset-blocked kernels, 1: heuristic, 2: whenever applicable
    int coset_log_twc = -1;   // >= 0 forces the row-segment width of the tile (TWc = 1 << v vectors)
    int coset_log_nt = 0;     // 7 or 8 forces the CTA size (128 / 256 threads); 0 = default (256)
    int coset_vpt = 16;       // vectors per thread when the shape is forced (8 or 16)
    bool coset_wide_cta = true; // 512-thread CTAs for the rank-12 weighted-apply tile
    int coset_few = 1;          // K3e / K3f (coset2.cuh) for passes with <= 8 x-masks: 0 off, 1 auto, 2 never the TMA
                                // kernel (K3f)
    int coset_few_ct = 0;       // column tiles per CTA of K3e (0 = all of them while the grid still fills the chip)
    bool pipeline = true;       // chunked H2D / kernel / D2H pipeline for large host-resident single-string applies
    size_t pipeline_min_bytes = 128ull << 20, pipeline_chunk_bytes = 32ull << 20;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t pipe_in[3] = {}, pipe_k[3] = {}, pipe_out[3] = {}, pipe_start = nullptr;
    bool etile = true;          // packed-FP32 per-string expectation kernel (complex64, 9-12 qubits)
    bool wtile = true;          // dedicated whole-column weighted-apply kernel (complex64, 11-12 qubits)
    int rcoset_mode = 1;        // register-resident coset kernel (x-mask rank <= 4): 0 never, 1 auto, 2 whenever applicable
    int rc_expval_ctas_per_sm = 16; // MODE 1 grid target: CTAs per SM (each walks n_sets / grid coset sets in turn)
    int dcoset = 1;             // FP64 tensor-core dense-coset kernel (complex128, x-mask rank 4 or 5): 0 never,
                                // 1 when the cost model below prefers it, 2 whenever applicable
    int rcoset_log_nt = 7;      // its CTA size (128 / 256 threads)
    Scratch stage_in, stage_out, stage_data, partials, work_a, work_b, meta;
    std::mutex mu;
};

struct fp_event
{
    cudaEvent_t ev = nullptr;
};

template <typename T> struct DeviceOp
{
    PackedOp<T> host;
    uint64_t *gx = nullptr;
    uint32_t *gstart = nullptr;
    uint64_t *sz = nullptr;
    Cx<T> *sc = nullptr;
    uint8_t *sodd = nullptr;
    PairChunk *chunks = nullptr; // strings of each group in chunks of <= kPairMS (paired expectation kernel)
    uint32_t n_chunks = 0;
    bool any_diag = false;
    int x_rank = 0; // GF(2) rank of the x-masks (capped at kCosetMaxRank + 1)

    // coset-blocked plans, built lazily per tile rank (key: rank = log2 rows per tile)
    struct CosetPassDev
    {
        CosetPassView<T> view{};
        // the pass' strings as pair chunks in pass-local coordinates (coset-tiled expectation values, etile.cuh)
        PairChunk const *echunks = nullptr;
        uint32_t n_echunks = 0;
        uint8_t const *esodd = nullptr;
        // the pass' strings as a kernel-parameter block (K3e / K3f): passes with <= 8 groups and <= 128 strings
        std::shared_ptr<FewStrings<T>> few;
        // ... and for passes with any number of groups (K3g): <= 768 strings, <= 256 groups, <= 30 qubits
        std::shared_ptr<GenStrings<T>> gen;
        std::vector<void *> allocs;
    };
    mutable std::map<int, std::vector<CosetPassDev>> coset_plans;

    // register-resident coset plan (x-mask rank <= kRcMaxRank), built lazily
    struct RcPlanDev
    {
        RcPassView<T> view{};
        int rr = 0;
        std::vector<void *> allocs;
    };
    mutable std::map<int, RcPlanDev> rc_plans;

    OpView<T> view() const
    {
        OpView<T> v{};
        v.gx = gx;
        v.gstart = gstart;
        v.sz = sz;
        v.scoef = sc;
        v.G = static_cast<uint32_t>(host.gx.size());
        if (host.sz.size() == 1)
        {
            v.x0 = host.gx[0];
            v.z0 = host.sz[0];
            v.c0 = Cx<T>{host.sc[0].real(), host.sc[0].imag()};
        }
        return v;
    }
    void release()
    {
        cudaFree(gx);
        cudaFree(gstart);
        cudaFree(sz);
        cudaFree(sc);
        cudaFree(sodd);
        cudaFree(chunks);
        for (auto &kv : coset_plans)
            for (auto &pd : kv.second)
                for (void *a : pd.allocs)
                    cudaFree(a);
        coset_plans.clear();
        for (auto &kv : rc_plans)
            for (void *a : kv.second.allocs)
                cudaFree(a);
        rc_plans.clear();
        gx = nullptr;
        gstart = nullptr;
        sz = nullptr;
        sc = nullptr;
        sodd = nullptr;
        chunks = nullptr;
    }
};

constexpr int kPairMS = 4;

struct fp_op
{
    int dtype = FP_C128;
    int device = 0;
    int n_qubits = 0;
    size_t n_strings = 0;
    DeviceOp<float> f;
    DeviceOp<double> d;
};

struct fp_sop
{
    int dtype = FP_C128;
    int device = 0;
    int n_qubits = 0;
    size_t n_strings = 0, n_ops = 0;
    fp_op *summed = nullptr; // PauliOp with c_j = sum_k coeffs(j,k)  (SummedPauliOp::apply)
    fp_op *strings = nullptr; // unmerged packed strings (unit coefficients) for apply_weighted / expectation_value
    void *A_w = nullptr;      // [2S x K] planar (-i)^nY coeffs, rows in packed order        (W = A_w * data)
    void *A_e = nullptr;      // [2K x S] planar coeffs * (-i)^nY * pair factor, transposed  (out = A_e * E)
};

// ================================================================ helpers
namespace
{
template <typename T> int upload_vec(T **dst, std::vector<T> const &v)
{
    void *p = nullptr;
    size_t bytes = v.size() * sizeof(T);
    if (bytes == 0)
    {
        FP_CU(cudaMalloc(&p, 16));
    }
    else
    {
        FP_CU(cudaMalloc(&p, bytes));
        FP_CU(cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice));
    }
    *dst = static_cast<T *>(p);
    return FP_OK;
}

template <typename T> int upload_op(DeviceOp<T> &d)
{
    FP_TRY(upload_vec(&d.gx, d.host.gx));
    FP_TRY(upload_vec(&d.gstart, d.host.gstart));
    FP_TRY(upload_vec(&d.sz, d.host.sz));
    {
        std::vector<Cx<T>> sc(d.host.sc.size());
        for (size_t i = 0; i < sc.size(); ++i)
            sc[i] = Cx<T>{d.host.sc[i].real(), d.host.sc[i].imag()};
        FP_TRY(upload_vec(&d.sc, sc));
    }
    FP_TRY(upload_vec(&d.sodd, d.host.sodd));
    std::vector<PairChunk> chunks;
    d.any_diag = false;
    for (size_t g = 0; g + 1 < d.host.gstart.size(); ++g)
    {
        uint64_t x = d.host.gx[g];
        uint32_t hbit = 0;
        if (x)
            hbit = 63u - static_cast<uint32_t>(__builtin_clzll(x));
        else
            d.any_diag = true;
        for (uint32_t s = d.host.gstart[g]; s < d.host.gstart[g + 1]; s += kPairMS)
        {
            PairChunk c;
            c.x = x;
            c.s0 = s;
            c.count = std::min<uint32_t>(kPairMS, d.host.gstart[g + 1] - s);
            c.hbit = hbit;
            c.diag = x == 0;
            chunks.push_back(c);
        }
    }
    d.n_chunks = static_cast<uint32_t>(chunks.size());
    FP_TRY(upload_vec(&d.chunks, chunks));
    {
        Gf2Basis bb;
        d.x_rank = 0;
        for (uint64_t x : d.host.gx)
            if (!bb.insert(x, kCosetMaxRank))
            {
                d.x_rank = kCosetMaxRank + 1;
                break;
            }
        if (d.x_rank == 0)
            d.x_rank = bb.r;
    }
    return FP_OK;
}

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

bool is_device_ptr(void const *p)
{
    if (!p)
        return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// A caller buffer, used in place when it lives on the device, staged through scratch when it lives on the host.
struct Staged
{
    void *dev = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    bool staged = false;
    bool zero_copy = false; // pinned host memory used in place by the kernel: the call must still synchronise
};

// Pinned (page-locked / registered) host memory is mapped into the device address space: single-pass streaming
// kernels can read and write it in place over PCIe, which overlaps the two directions inside one launch instead of
// H2D copy -> kernel -> D2H copy back to back.
void *pinned_device_alias(void const *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer)
        return nullptr;
    return a.devicePointer;
}

int stage_in(fp_ctx *ctx, Scratch &scratch, void const *p, size_t bytes, bool copy, Staged &s,
             bool allow_zero_copy = false)
{
    s.bytes = bytes;
    if (bytes == 0)
    {
        s.dev = const_cast<void *>(p);
        return FP_OK;
    }
    if (!p)
        return set_err(FP_INVALID_ARGUMENT, "null data pointer");
    if (is_device_ptr(p))
    {
        s.dev = const_cast<void *>(p);
        return FP_OK;
    }
    if (allow_zero_copy && ctx->zero_copy)
    {
        if (void *alias = pinned_device_alias(p))
        {
            s.dev = alias;
            s.host = const_cast<void *>(p);
            s.zero_copy = true;
            return FP_OK;
        }
    }
    FP_TRY(scratch.ensure(bytes));
    s.dev = scratch.p;
    s.host = const_cast<void *>(p);
    s.staged = true;
    if (copy)
        FP_CU(cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FP_OK;
}

int stage_back(fp_ctx *ctx, Staged &s)
{
    if (s.staged && s.bytes)
        FP_CU(cudaMemcpyAsync(s.host, s.dev, s.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FP_OK;
}

int finish(fp_ctx *ctx, bool any_staged)
{
    FP_CU(cudaGetLastError());
    if (any_staged || !ctx->async)
        FP_CU(cudaStreamSynchronize(ctx->stream));
    return FP_OK;
}

struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

size_t csize(int dtype)
{
    return dtype == FP_C128 ? 16 : 8;
}

int check_dtype(int dtype)
{
    if (dtype != FP_C64 && dtype != FP_C128)
        return set_err(FP_INVALID_ARGUMENT, "dtype must be FP_C64 or FP_C128");
    return FP_OK;
}

uint64_t dim_of(int n)
{
    return n > 0 ? (1ull << n) : 0; // PS:266-269: an empty string has dim 0
}

// ---------------------------------------------------------------- geometry
struct GeomSel
{
    Geom g{};
    int V = 4; // rows per thread
    int J = 1; // vectors per thread along the row (strided by TW)
    uint64_t grid = 0;
};

// rows: rows the kernel iterates over (dim, or dim/2 pair-rows); dim: rows of the state (L2 working set).
// wantJ: let each thread own 4 vectors of its row so row-only factors are amortised (multi-string operators).
GeomSel choose_geom(fp_ctx const *ctx, uint64_t rows, uint64_t dim, uint64_t rowvecs, size_t vec_bytes, bool multi_group,
                    bool reduce, uint64_t n_chunks = 1, bool wantJ = false)
{
    GeomSel s;
    // tile width in vectors (power of two)
    uint32_t const wcap = wantJ ? 1024u : static_cast<uint32_t>(kThreads);
    uint32_t w = 1;
    while (w < rowvecs && w < wcap)
        w <<= 1;
    if (multi_group)
    {
        // batch-tile the sweep so dim x tile stays L2-resident while all x-groups gather from it;
        // never go below one 64-byte DRAM granule per row
        uint32_t floor_w = static_cast<uint32_t>(std::max<size_t>(1, 64 / vec_bytes));
        while (w > floor_w && dim * w * vec_bytes > ctx->l2_budget)
            w >>= 1;
    }
    int J = (wantJ && w >= 8) ? 4 : 1;
    uint32_t tw = std::min<uint32_t>(w / J, kThreads);
    uint32_t log2tw = 0;
    while ((1u << log2tw) < tw)
        ++log2tw;
    uint32_t const TY = kThreads / tw;
    uint32_t const nct = static_cast<uint32_t>((rowvecs + static_cast<uint64_t>(tw) * J - 1) / (static_cast<uint64_t>(tw) * J));
    int V = (J == 4) ? 2 : 4;
    {
        uint64_t blocksV = ((rows + static_cast<uint64_t>(TY) * V - 1) / (static_cast<uint64_t>(TY) * V)) * nct * n_chunks;
        if (rows < static_cast<uint64_t>(TY) * V || blocksV < static_cast<uint64_t>(ctx->sm_count) * 2)
            V = 1;
    }
    uint64_t const rows_per_iter = static_cast<uint64_t>(TY) * V;
    uint64_t const n_row_iters = (rows + rows_per_iter - 1) / rows_per_iter;
    s.V = V;
    s.J = J;
    s.g.N = rows;
    s.g.rowvecs = rowvecs;
    s.g.nColTiles = nct;
    s.g.log2TW = log2tw;
    if (!reduce)
    {
        s.g.iters = 1;
        s.g.nRowBlocks = n_row_iters;
    }
    else
    {
        // about 8 CTAs per SM, never more (iters rounds UP): with 4 resident CTAs per SM that is two full waves and
        // no straggler third wave
        uint64_t const target = static_cast<uint64_t>(ctx->sm_count) * 8;
        uint64_t const fixed = static_cast<uint64_t>(nct) * n_chunks;
        uint64_t want_rb = std::max<uint64_t>(1, target / fixed);
        uint64_t iters = std::max<uint64_t>(1, (n_row_iters + want_rb - 1) / want_rb);
        iters = std::min<uint64_t>(iters, 1024);
        s.g.iters = static_cast<uint32_t>(iters);
        s.g.nRowBlocks = (n_row_iters + iters - 1) / iters;
    }
    s.grid = s.g.nRowBlocks * s.g.nColTiles * n_chunks;
    return s;
}

template <typename T> int pick_epv(void const *a, void const *b, uint64_t B)
{
    if (sizeof(T) == 8)
        return 1;
    bool aligned = (reinterpret_cast<uintptr_t>(a) % 16 == 0) && (reinterpret_cast<uintptr_t>(b) % 16 == 0);
    return (B % 2 == 0 && aligned) ? 2 : 1;
}

int check_align(void const *p, size_t align, char const *what)
{
    if (reinterpret_cast<uintptr_t>(p) % align)
        return set_err(FP_INVALID_ARGUMENT, std::string(what) + " must be " + std::to_string(align) + "-byte aligned");
    return FP_OK;
}

int check_grid(uint64_t grid)
{
    if (grid == 0 || grid > 0x7fffffffull)
        return set_err(FP_UNSUPPORTED, "problem too large for a single launch (grid " + std::to_string(grid) + ")");
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
                        bool tma_kernels = false)
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
            double const cost = static_cast<double>(passes->size()) * seg_cost[v] *
                                ((tma_kernels && v == 4 && lnt_pref == 8) ? 0.7 : 1.0);
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

template <typename T, int EPV, int LOG_TWC, int NBUF, bool PSTR, int MODE = 0>
int launch_coset_few_v(fp_ctx *ctx, CosetPassView<T> const &view, FewStrings<T> const &strs, int n_qubits,
                       uint64_t rowvecs, void const *in, void *out, int beta, void *partials = nullptr,
                       uint32_t Bpad = 0)
{
    using Cfg = FewCfg<LOG_TWC>;
    constexpr int GMAX = 8;
    constexpr size_t smem = NBUF * Cfg::TILE_BYTES;
    static PerDevice configured; // per template instance
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(coset_few_kernel<T, EPV, LOG_TWC, GMAX, NBUF, 2, PSTR, false, MODE>,
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
    coset_few_kernel<T, EPV, LOG_TWC, GMAX, NBUF, 2, PSTR, false, MODE>
        <<<static_cast<unsigned>(grid), Cfg::NT, smem, ctx->stream>>>(
            view, rowvecs, nct, per, groups, static_cast<CVec<T, EPV> const *>(in), static_cast<CVec<T, EPV> *>(out), beta,
            strs, static_cast<Cx<T> *>(partials), Bpad);
    ctx->launches++;
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
    *launched = true;
    return FP_OK;
}

// Picks the variant for one pass; *launched = false when the pass has to go through coset_kernel (K3b).
template <typename T, int EPV, int MODE = 0>
int launch_coset_few(fp_ctx *ctx, typename DeviceOp<T>::CosetPassDev const &pd, int n_qubits, uint64_t rowvecs,
                     void const *in, void *out, int beta, bool *launched, void *partials = nullptr, uint32_t Bpad = 0)
{
    *launched = false;
    CosetPassView<T> const &view = pd.view;
    if (MODE == 0 && ctx->coset_few == 1 && pd.gen && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
        is_device_ptr(in))
        return launch_coset_gen_tma<T, EPV>(ctx, view, *pd.gen, n_qubits, rowvecs, in, out, beta, launched);
    if (!ctx->coset_few || view.n_groups == 0 || view.n_groups > 8 || n_qubits < 8)
        return FP_OK;
    static FewStrings<T> const no_strings{};
    bool const pstr = pd.few != nullptr;
    FewStrings<T> const &strs = pstr ? *pd.few : no_strings;
    // K3f wins on overwrite passes of large registers (measured at 20 qubits: 4 masks 0.42 -> 0.38 ms, 256 columns
    // 1.95 -> 1.85 ms; 8 masks equal); read-modify-write passes and small registers stay on the resident-CTA kernel
    if (MODE == 0 && ctx->coset_few == 1 && pstr && beta == 0 && n_qubits >= 16 && n_qubits <= 30 && rowvecs % 16 == 0 &&
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
    bool const tma_ok = MODE == 0 && ctx->coset_few == 1 && n_qubits >= 12 && n_qubits <= 30 && rowvecs % 16 == 0 &&
                        is_device_ptr(in) && tensor_map_encoder() != nullptr;
    CosetShape const shape = choose_coset<T>(ctx, op, n_qubits, rowvecs, epv, tma_ok);
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
    if (MODE == 1)
        FP_TRY(ctx->partials.ensure(n_cosets * Bpad * 2 * sizeof(T)));
    for (size_t p = 0; p < passes->size(); ++p)
    {
        int const b = (p == 0) ? beta : 1;
        if constexpr (MODE == 0 || MODE == 1)
        {
            if (shape.rank() == 8 && shape.log_nt == 8)
            {
                bool launched = false;
                FP_TRY((launch_coset_few<T, EPV, MODE>(ctx, (*passes)[p], n_qubits, rowvecs, in, out, b, &launched,
                                                       ctx->partials.p, Bpad)));
                if (launched)
                {
                    if (MODE == 1)
                    {
                        unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
                        finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
                            static_cast<Cx<T> const *>(ctx->partials.p), n_cosets, Bpad, B, static_cast<Cx<T> *>(out), b);
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

// K3d (dcoset.cuh): FP64 tensor-core dense-coset kernel, complex128 apply, rank 4 (one warp per coset) or 5 (two)
template <int RR, int WPC, int PFD, int MODE>
int launch_dcoset(fp_ctx *ctx, RcPassView<double> const &view, uint32_t n_strings, uint64_t n_cosets, uint64_t rowvecs,
                  void const *in, void *out, int beta, uint64_t B)
{
    using Cfg = DcosetCfg<RR, WPC>;
    constexpr size_t smem = MODE == 1 ? Cfg::smem_expval : Cfg::smem;
    static int resident = 0; // CTAs per SM (per template instance): the kernel is persistent
    static PerDevice configured;
    if (!configured.done(ctx->device))
    {
        FP_CU(cudaFuncSetAttribute(dcoset_kernel<RR, WPC, PFD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
        int nb = 0;
        FP_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dcoset_kernel<RR, WPC, PFD, MODE>, Cfg::NT, smem));
        resident = std::max(1, nb);
        configured.set(ctx->device);
    }
    uint64_t const sets = (n_cosets + Cfg::CPI - 1) / Cfg::CPI;
    unsigned const ny = MODE == 1 ? static_cast<unsigned>((rowvecs + Cfg::ECOLS - 1) / Cfg::ECOLS) : 1u;
    uint64_t const gx = std::min<uint64_t>(sets, std::max<uint64_t>(1, static_cast<uint64_t>(ctx->sm_count) * resident / ny));
    uint32_t const Bpad = static_cast<uint32_t>((B + 3) & ~3ull);
    if (MODE == 1)
        FP_TRY(ctx->partials.ensure(gx * Bpad * sizeof(Cx<double>)));
    dcoset_kernel<RR, WPC, PFD, MODE><<<dim3(static_cast<unsigned>(gx), ny), Cfg::NT, smem, ctx->stream>>>(
        view, n_strings, n_cosets, rowvecs, static_cast<CVec<double, 1> const *>(in),
        MODE == 1 ? nullptr : static_cast<CVec<double, 1> *>(out), beta, static_cast<Cx<double> *>(ctx->partials.p), Bpad);
    ctx->launches++;
    if (MODE == 1)
    {
        unsigned fgrid = static_cast<unsigned>((B + kFinX - 1) / kFinX);
        finalize_complex_kernel<double><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
            static_cast<Cx<double> const *>(ctx->partials.p), gx, Bpad, B, static_cast<Cx<double> *>(out), beta);
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
    if constexpr (sizeof(T) == 8)
    {
        // ranks 4 and 5 are GEMM-shaped per coset (16x16 / 32x32 complex): FP64 tensor cores
        // The tensor-core form costs 2^rr complex FMAs per amplitude however few of the 2^rr coset masks occur; the
        // SIMT forms cost one per occurring mask.  Measured at 20 qubits x 64 columns (scripts/dispatch_sweep.py,
        // profiles/r01s3_dispatch_sweep.txt): rank 4 -- DMMA 0.45 ms apply / 0.53 ms expectation value against
        // 0.37 / 0.43 / 0.49 / 0.57 ms (apply) and 0.34 / 0.42 / 0.50 / 0.58 ms (expectation value) for the register
        // kernel at 4 / 8 / 12 / 16 masks; rank 5 -- DMMA 0.82-0.94 ms against 0.49 / 0.58 / 0.68 / 0.81 / 1.09 ms for
        // the shared-memory coset kernel at 5 / 8 / 12 / 16 / 24 masks.
        size_t const n_masks = op.host.gx.size();
        bool const dense_enough = rr == 4 ? n_masks > (MODE == 1 ? 12u : 8u) : n_masks > 16u;
        if ((ctx->dcoset == 2 || (ctx->dcoset == 1 && dense_enough)) && (rr == 4 || rr == 5) && rowvecs >= 8)
        {
            typename DeviceOp<T>::RcPlanDev const *dplan = nullptr;
            FP_TRY(get_rc_plan<T>(op, n_qubits, rr, &dplan));
            uint64_t const nc = 1ull << (n_qubits - rr);
            uint32_t const ns = static_cast<uint32_t>(op.host.sz.size());
            if (rr == 4)
                FP_TRY((launch_dcoset<4, 1, 4, MODE>(ctx, dplan->view, ns, nc, rowvecs, in, out, beta, B)));
            else
                FP_TRY((launch_dcoset<5, 2, 2, MODE>(ctx, dplan->view, ns, nc, rowvecs, in, out, beta, B)));
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
        finalize_complex_kernel<T><<<fgrid, dim3(kFinX, kFinY), 0, ctx->stream>>>(
            static_cast<Cx<T> const *>(ctx->partials.p), n_blocks, Bpad, B, static_cast<Cx<T> *>(out), beta);
        ctx->launches++;
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

template <typename T, int EPV, int MS>
void launch_pairs_v(fp_ctx *ctx, GeomSel const &gs, PairChunk const *chunks, uint64_t const *sz, uint8_t const *sodd,
                    PairChunk inl, uint64_t inl_z, uint32_t inl_odd, int use_inline, uint64_t dim, void const *in,
                    T *partials, uint64_t slot_stride)
{
    auto const *din = static_cast<CVec<T, EPV> const *>(in);
    dim3 grid(static_cast<unsigned>(gs.grid));
    if (gs.V == 4)
        expval_pairs_kernel<T, EPV, 4, MS><<<grid, kThreads, 0, ctx->stream>>>(
            chunks, sz, sodd, inl, inl_z, inl_odd, use_inline, gs.g, dim, din, partials, slot_stride);
    else
        expval_pairs_kernel<T, EPV, 1, MS><<<grid, kThreads, 0, ctx->stream>>>(
            chunks, sz, sodd, inl, inl_z, inl_odd, use_inline, gs.g, dim, din, partials, slot_stride);
    ctx->launches++;
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
    uint64_t cols = 16 * vec_cols;                 // 256 bytes per row
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

template <typename T> DeviceOp<T> &dop(fp_op *op);
template <> DeviceOp<float> &dop<float>(fp_op *op)
{
    return op->f;
}
template <> DeviceOp<double> &dop<double>(fp_op *op)
{
    return op->d;
}
template <typename T> DeviceOp<T> const &dop(fp_op const *op)
{
    return dop<T>(const_cast<fp_op *>(op));
}

template <typename T>
int op_create_t(fp_ctx *ctx, int dtype, int n, size_t S, uint8_t const *codes, std::complex<T> const *coeffs,
                bool merge, fp_op **out)
{
    std::unique_ptr<fp_op> op(new fp_op);
    op->dtype = dtype;
    op->device = ctx->device;
    op->n_qubits = n;
    op->n_strings = S;
    try
    {
        dop<T>(op.get()).host = pack_op<T>(n, S, codes, coeffs, merge);
    }
    catch (std::invalid_argument const &e)
    {
        return set_err(FP_INVALID_ARGUMENT, e.what());
    }
    int rc = upload_op(dop<T>(op.get()));
    if (rc != FP_OK)
    {
        dop<T>(op.get()).release();
        return rc;
    }
    *out = op.release();
    return FP_OK;
}

int op_check(fp_ctx *ctx, fp_op const *op)
{
    if (!ctx || !op)
        return set_err(FP_INVALID_ARGUMENT, "null context or operator");
    if (op->device != ctx->device)
        return set_err(FP_INVALID_ARGUMENT, "operator plan was created on a different device than the context");
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
        return set_err(code, msg ? msg : "");
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
            ctx->coset_few = atoi(env);
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
        for (Scratch *s : {&ctx->stage_in, &ctx->stage_out, &ctx->stage_data, &ctx->partials, &ctx->work_a, &ctx->work_b,
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
        if (!ctx || mode < 0 || mode > 2 || column_tiles_per_cta < 0)
            return set_err(FP_INVALID_ARGUMENT, "bad few-mask coset mode");
        ctx->coset_few = mode;
        ctx->coset_few_ct = column_tiles_per_cta;
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
