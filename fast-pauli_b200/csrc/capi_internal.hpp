// Shared internals of libfastpauli_b200.so's translation units (capi.cu: contexts, memory, PauliString / PauliOp;
// sop_capi.cu: SummedPauliOp, square, one-shot entry points).  Everything here is header-only and lives in an
// anonymous namespace or is a template: each translation unit gets its own copy, nothing is exported.
#pragma once
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
#include "coset3.cuh"
#include "coset4.cuh"
#include "rcoset.cuh"
#include "kernels.cuh"
#include "pack.hpp"

using namespace fpk;

// ================================================================ errors
namespace
{
// the message store (thread_local) lives in capi.cu behind fp_internal_set_error
inline int set_err(int code, std::string msg)
{
    return fp_internal_set_error(code, msg.c_str());
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
    uint32_t coset_kernels = 0; // coset-family kernels launched since the last reset: 1 K3b, 2 K3e, 4 K3f, 8 K3g, 16 K3i
    int last_gemm_engine = -1; // 0 = SIMT, 1 = tcgen05 (diagnostics)
    size_t l2_budget = 40ull << 20;
    int coset_mode = 1;       // 0: never use the coset-blocked kernels, 1: heuristic, 2: whenever applicable
    int coset_log_twc = -1;   // >= 0 forces the row-segment width of the tile (TWc = 1 << v vectors)
    int coset_log_nt = 0;     // 7 or 8 forces the CTA size (128 / 256 threads); 0 = default (256)
    int coset_vpt = 16;       // vectors per thread when the shape is forced (8 or 16)
    bool coset_wide_cta = true; // 512-thread CTAs for the rank-12 weighted-apply tile
    int coset_few = 1;          // K3e / K3f / K3i / K3j for passes with few x-masks: 0 off, 1 auto, 2 never the TMA-fed
                                // kernels, 3 auto without K3i / K3j, 4 auto without K3j
    bool coset_pair_all = false; // K3j also for passes whose masks carry one string each (set by mode 5 = mode 1 + this)
    int coset_few_ct = 0;       // column tiles per CTA of K3e (0 = all of them while the grid still fills the chip)
    bool pipeline = true;       // chunked H2D / kernel / D2H pipeline for large host-resident single-string applies
    size_t pipeline_min_bytes = 128ull << 20, pipeline_chunk_bytes = 32ull << 20;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t pipe_in[3] = {}, pipe_k[3] = {}, pipe_out[3] = {}, pipe_start = nullptr;
    bool etile = true;          // packed-FP32 per-string expectation kernel (complex64, 9-12 qubits)
    bool wtile = true;          // dedicated whole-column weighted-apply kernel (complex64, 11-12 qubits)
    int rcoset_mode = 1;        // register-resident coset kernel (x-mask rank <= 4): 0 never, 1 auto, 2 whenever applicable
    int rc_expval_ctas_per_sm = 64; // MODE 1 grid target: CTAs per SM (each walks n_sets / grid coset sets in turn)
    int dcoset = 1;             // FP64 tensor-core dense-coset kernel (complex128, x-mask rank 4 or 5): 0 never,
                                // 1 when the cost model below prefers it, 2 whenever applicable
    int rcoset_log_nt = 7;      // its CTA size (128 / 256 threads)
    Scratch stage_in, stage_out, stage_data, partials, partials2, work_a, work_b, meta;
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
        // ... and for passes whose x-masks carry one string each (K3i, coset3.cuh): <= 32 masks, <= 30 qubits
        std::shared_ptr<DirStrings<T>> dir;
        // ... and for passes of eight independent x-masks with <= 64 strings (K3j, coset4.cuh): paired-mask basis
        std::shared_ptr<PairStrings<T>> pair;
        std::vector<void *> allocs;
    };
    mutable std::map<int, std::vector<CosetPassDev>> coset_plans;

    // single-state form (batch of ONE complex128 state of >= 22 qubits, K3i): the state is viewed as 2^(n-4) rows x 16
    // columns (the 4 lowest index bits), the operator re-planned on the upper n-4 bits; built lazily, see coset_launch.hpp
    struct SingleStatePass
    {
        CosetPassView<T> view{}; // basis + nonpivot_mask only
        DirStrings<T> strs{};
    };
    mutable std::vector<SingleStatePass> single_state_plan;
    mutable int single_state_tried = 0; // 0 not yet, 1 plan usable, -1 not applicable

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

// ---------------------------------------------------------------- fp_op helpers
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


// ---------------------------------------------------------------- paired expectation kernel launch (K2 / K4)
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

} // namespace
