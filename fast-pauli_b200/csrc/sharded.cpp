// K7: PauliOp::apply / expectation_value on ONE state vector sharded by its high index bits over the GPUs of a box
// (BASELINE config 5: 34 qubits = 256 GiB over 8 B200).  One process per GPU; this translation unit is host code
// only: it composes the library's own single-GPU entry points (fp_op_*, fp_string_apply) with NCCL point-to-point
// transfers over NVLink / NVSwitch.  No torch, no MPI: the caller ships the 128-byte ncclUniqueId to the ranks
// however it likes (a file, a socket, torchrun's store).
//
// Reference semantics (absent in the reference, which cannot go past 30 qubits: __pauli_string.hpp:57): with the
// global row index i = (r << n_local) | i_lo and a string's masks split the same way,
//     out_r[i_lo] += h (-i)^nY (-1)^popc(r & z_hi) (-1)^popc(i_lo & z_lo) psi_{r ^ x_hi}[i_lo ^ x_lo]      (PO:362-383)
// so strings are grouped by x_hi.  The x_hi = 0 class is an ordinary local PauliOp (all fused kernels apply).  Every
// other class needs the shard of peer r ^ x_hi.  It is streamed in CHUNKS of 2^m rows (default 256 MiB) through two
// receive buffers: ncclGroupStart / ncclSend(my chunk k) / ncclRecv(peer's chunk k) / ncclGroupEnd on a
// communication stream, while the compute stream applies the class' strings to the previous chunk -- source chunk k
// of string s lands in output block k ^ (x_lo >> m) with the in-block permutation x_lo & (2^m - 1), the block sign
// (-1)^popc(block & z_lo >> m) folded into the coefficient, i.e. one streaming single-string kernel (K1) per string
// and chunk.  Memory: state + output + 2 chunks (no shard-sized receive buffers).
#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h> // types and prototypes only: the library is dlopen'ed on first use (no link-time dependency)

#include "../../include/fastpauli_b200.h"
#include "internal.h"

namespace
{
enum
{
    OK = 0,
    INVALID = 1,
    CUDA_ERR = 2,
    NO_DEVICE = 3
};

int fail(int code, std::string const &msg)
{
    return fp_internal_set_error(code, msg.c_str());
}

#define SH_CU(call)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(CUDA_ERR, std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
    } while (0)
#define SH_TRY(call)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc_ = (call);                                                                                              \
        if (rc_ != 0)                                                                                                  \
            return rc_;                                                                                                \
    } while (0)

// ---- NCCL, loaded lazily (torch's bundled copy is picked up when it is already mapped into the process)
struct Nccl
{
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

Nccl &nccl()
{
    static Nccl n = []() {
        Nccl t;
        char const *env = getenv("FASTPAULI_NCCL_LIB");
        char const *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (char const *name : names)
        {
            if (!name)
                continue;
            t.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (t.handle)
                break;
        }
        if (!t.handle)
        {
            t.error = "NCCL library not found (libnccl.so.2; set FASTPAULI_NCCL_LIB)";
            return t;
        }
#define SH_SYM(name)                                                                                                   \
    t.name = reinterpret_cast<decltype(t.name)>(dlsym(t.handle, "nccl" #name));                                        \
    if (!t.name)                                                                                                       \
        t.error = "NCCL symbol nccl" #name " missing";
        SH_SYM(GetUniqueId)
        SH_SYM(CommInitRank)
        SH_SYM(CommDestroy)
        SH_SYM(GroupStart)
        SH_SYM(GroupEnd)
        SH_SYM(Send)
        SH_SYM(Recv)
        SH_SYM(AllReduce)
        SH_SYM(GetErrorString)
        SH_SYM(GetVersion)
#undef SH_SYM
        return t;
    }();
    return n;
}

#define SH_NCCL(call)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
            return fail(CUDA_ERR, std::string(#call) + ": " + nccl().GetErrorString(r_));                              \
    } while (0)

int log2_exact(uint64_t v)
{
    if (v == 0 || (v & (v - 1)))
        return -1;
    return 63 - __builtin_clzll(v);
}

struct DeviceScope
{
    int prev = -1;
    explicit DeviceScope(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
    }
    ~DeviceScope()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};
} // namespace

struct fp_comm
{
    fp_ctx *user_ctx = nullptr; // the caller's context: synchronised at entry (its stream may still produce `in`)
    fp_ctx *ctx = nullptr;      // private context of the sharded calls: own stream, asynchronous
    int device = 0, world = 1, rank = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t compute = nullptr, comm_stream = nullptr;
    cudaEvent_t ready[2] = {}, done[2] = {};
    void *bufs[2] = {};
    size_t buf_bytes = 0;
    double *scratch = nullptr; // 64 doubles for small reductions
    bool emulated = false;     // no NCCL: one process plays rank `rank`, peer chunks are copied from a buffer that
                               // holds every shard (tests on a single GPU)
};

struct ShardedString
{
    std::vector<uint8_t> low; // n_local codes
    uint64_t x_lo = 0, z_lo = 0;
    std::complex<double> c;   // h (-i)^nY_hi (-1)^popc(rank & z_hi)
};

struct ShardedClass
{
    uint64_t x_hi = 0;
    std::vector<ShardedString> strings;
    fp_op *op = nullptr; // the class as one fused PauliOp on n_local qubits (whole-shard mode)
};

struct fp_sharded_op
{
    fp_comm *comm = nullptr;
    int dtype = FP_C128;
    int n_qubits = 0, n_local = 0;
    fp_op *local_op = nullptr;           // the x_hi = 0 class (may be absent)
    std::vector<ShardedClass> remote;    // sorted by x_hi
    uint64_t chunk_rows = 0;             // 0 = pick from chunk_bytes
    size_t chunk_bytes = 256ull << 20;
    int mode = 0;                        // 0 auto, 1 chunked streaming, 2 whole-shard receive + fused class operator
    int last_mode = 0;
    // statistics of the last call
    uint64_t last_bytes_sent = 0, last_chunks = 0, last_kernels = 0;
    float last_ms = 0; // device time of the last apply (CUDA events on the compute stream, which waits for the exchange)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};

extern "C"
{
    int fp_comm_unique_id(unsigned char *id)
    {
        if (!id)
            return fail(INVALID, "null pointer");
        Nccl &n = nccl();
        if (!n.error.empty())
            return fail(NO_DEVICE, n.error);
        static_assert(sizeof(ncclUniqueId) == FP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        ncclUniqueId uid;
        SH_NCCL(n.GetUniqueId(&uid));
        std::memcpy(id, &uid, sizeof uid);
        return OK;
    }

    int fp_comm_create(fp_ctx *ctx, const unsigned char *id, int world, int rank, fp_comm **out)
    {
        if (!ctx || !id || !out)
            return fail(INVALID, "null pointer");
        *out = nullptr;
        if (world <= 0 || rank < 0 || rank >= world || log2_exact(static_cast<uint64_t>(world)) < 0)
            return fail(INVALID, "world must be a power of two and 0 <= rank < world");
        Nccl &n = nccl();
        if (!n.error.empty())
            return fail(NO_DEVICE, n.error);
        std::unique_ptr<fp_comm> c(new fp_comm);
        c->user_ctx = ctx;
        c->world = world;
        c->rank = rank;
        SH_TRY(fp_ctx_device(ctx, &c->device));
        DeviceScope scope(c->device);
        SH_TRY(fp_ctx_create(c->device, &c->ctx));
        SH_CU(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        SH_CU(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
        SH_TRY(fp_ctx_set_stream(c->ctx, c->compute, 1));
        SH_TRY(fp_ctx_set_async(c->ctx, 1));
        for (int i = 0; i < 2; ++i)
        {
            SH_CU(cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming));
            SH_CU(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
        }
        SH_CU(cudaMalloc(reinterpret_cast<void **>(&c->scratch), 64 * sizeof(double)));
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof uid);
        SH_NCCL(n.CommInitRank(&c->comm, world, uid, rank));
        *out = c.release();
        return OK;
    }

    int fp_comm_create_emulated(fp_ctx *ctx, int world, int rank, fp_comm **out)
    {
        if (!ctx || !out)
            return fail(INVALID, "null pointer");
        *out = nullptr;
        if (world <= 0 || rank < 0 || rank >= world || log2_exact(static_cast<uint64_t>(world)) < 0)
            return fail(INVALID, "world must be a power of two and 0 <= rank < world");
        std::unique_ptr<fp_comm> c(new fp_comm);
        c->user_ctx = ctx;
        c->world = world;
        c->rank = rank;
        c->emulated = true;
        SH_TRY(fp_ctx_device(ctx, &c->device));
        DeviceScope scope(c->device);
        SH_TRY(fp_ctx_create(c->device, &c->ctx));
        SH_CU(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        SH_CU(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
        SH_TRY(fp_ctx_set_stream(c->ctx, c->compute, 1));
        SH_TRY(fp_ctx_set_async(c->ctx, 1));
        for (int i = 0; i < 2; ++i)
        {
            SH_CU(cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming));
            SH_CU(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
        }
        *out = c.release();
        return OK;
    }

    int fp_comm_destroy(fp_comm *c)
    {
        if (!c)
            return OK;
        DeviceScope scope(c->device);
        cudaStreamSynchronize(c->compute);
        cudaStreamSynchronize(c->comm_stream);
        if (c->comm && !c->emulated)
            nccl().CommDestroy(c->comm);
        for (int i = 0; i < 2; ++i)
        {
            cudaFree(c->bufs[i]);
            cudaEventDestroy(c->ready[i]);
            cudaEventDestroy(c->done[i]);
        }
        cudaFree(c->scratch);
        fp_ctx_destroy(c->ctx);
        cudaStreamDestroy(c->compute);
        cudaStreamDestroy(c->comm_stream);
        delete c;
        return OK;
    }

    int fp_comm_info(const fp_comm *c, int *world, int *rank, int *nccl_version)
    {
        if (!c)
            return fail(INVALID, "null pointer");
        if (world)
            *world = c->world;
        if (rank)
            *rank = c->rank;
        if (nccl_version)
            nccl().GetVersion(nccl_version);
        return OK;
    }

    // sum (op = 0), max (op = 1) or min (op = 2) of n <= 64 host doubles over the ranks, in place; also a barrier
    int fp_comm_allreduce_f64(fp_comm *c, double *values, size_t n, int op)
    {
        if (!c || (!values && n))
            return fail(INVALID, "null pointer");
        if (n > 64)
            return fail(INVALID, "at most 64 values");
        if (c->emulated)
            return fail(INVALID, "collectives need a real communicator (fp_comm_create)");
        DeviceScope scope(c->device);
        double zero = 0;
        if (n == 0)
        {
            values = &zero;
            n = 1;
        }
        SH_CU(cudaMemcpyAsync(c->scratch, values, n * sizeof(double), cudaMemcpyHostToDevice, c->comm_stream));
        SH_NCCL(nccl().AllReduce(c->scratch, c->scratch, n, ncclDouble, op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum),
                                 c->comm, c->comm_stream));
        SH_CU(cudaMemcpyAsync(values, c->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, c->comm_stream));
        SH_CU(cudaStreamSynchronize(c->comm_stream));
        return OK;
    }

    int fp_comm_barrier(fp_comm *c)
    {
        return fp_comm_allreduce_f64(c, nullptr, 0, 0);
    }

    int fp_sharded_op_create(fp_comm *c, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes,
                             const void *coeffs, fp_sharded_op **out)
    {
        if (!c || !out || (n_strings && (!codes || !coeffs)))
            return fail(INVALID, "null pointer");
        *out = nullptr;
        if (dtype != FP_C64 && dtype != FP_C128)
            return fail(INVALID, "dtype must be FP_C64 or FP_C128");
        int const n_hi = log2_exact(static_cast<uint64_t>(c->world));
        if (n_qubits < n_hi || n_qubits > 62)
            return fail(INVALID, "n_qubits must be in [log2(world), 62]");
        if (n_strings == 0)
            return fail(INVALID, "a sharded operator needs at least one string");
        int const n_local = n_qubits - n_hi;
        std::unique_ptr<fp_sharded_op> op(new fp_sharded_op);
        op->comm = c;
        op->dtype = dtype;
        op->n_qubits = n_qubits;
        op->n_local = n_local;
        if (char const *env = getenv("FASTPAULI_SHARD_CHUNK_BYTES"))
            op->chunk_bytes = std::max<size_t>(1024, strtoull(env, nullptr, 10));

        static std::complex<double> const phase[4] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}};
        std::map<uint64_t, ShardedClass> classes;
        for (size_t s = 0; s < n_strings; ++s)
        {
            uint8_t const *cs = codes + s * static_cast<size_t>(n_qubits);
            uint64_t x_hi = 0, z_hi = 0;
            int ny_hi = 0;
            for (int q = 0; q < n_qubits; ++q)
                if (cs[q] > 3)
                    return fail(INVALID, "Pauli code must be 0, 1, 2, or 3");
            for (int q = 0; q < n_hi; ++q) // left-most character = most significant qubit (PS:52-54)
            {
                uint64_t const bit = 1ull << (n_hi - 1 - q);
                if (cs[q] == 1 || cs[q] == 2)
                    x_hi |= bit;
                if (cs[q] == 2 || cs[q] == 3)
                    z_hi |= bit;
                ny_hi += cs[q] == 2;
            }
            ShardedString st;
            st.low.assign(cs + n_hi, cs + n_qubits);
            for (int q = 0; q < n_local; ++q)
            {
                uint64_t const bit = 1ull << (n_local - 1 - q);
                if (st.low[q] == 1 || st.low[q] == 2)
                    st.x_lo |= bit;
                if (st.low[q] == 2 || st.low[q] == 3)
                    st.z_lo |= bit;
            }
            std::complex<double> h = dtype == FP_C128
                                         ? static_cast<std::complex<double> const *>(coeffs)[s]
                                         : std::complex<double>(static_cast<std::complex<float> const *>(coeffs)[s]);
            double const sign = (__builtin_popcountll(static_cast<uint64_t>(c->rank) & z_hi) & 1) ? -1.0 : 1.0;
            st.c = h * phase[ny_hi & 3] * sign;
            ShardedClass &cl = classes[x_hi];
            cl.x_hi = x_hi;
            cl.strings.push_back(std::move(st));
        }
        DeviceScope scope(c->device);
        // a class as an ordinary PauliOp on n_local qubits with the rank-dependent signs in the coefficients
        auto make_class_op = [&](ShardedClass const &cl, fp_op **dst) -> int {
            size_t const S0 = cl.strings.size();
            std::vector<uint8_t> low(S0 * static_cast<size_t>(std::max(n_local, 1)));
            std::vector<std::complex<double>> cd(S0);
            std::vector<std::complex<float>> cf(S0);
            for (size_t s = 0; s < S0; ++s)
            {
                if (n_local)
                    std::memcpy(&low[s * n_local], cl.strings[s].low.data(), n_local);
                cd[s] = cl.strings[s].c;
                cf[s] = std::complex<float>(cl.strings[s].c);
            }
            return fp_op_create(c->ctx, dtype, n_local, S0, low.data(),
                                dtype == FP_C128 ? static_cast<void const *>(cd.data()) : static_cast<void const *>(cf.data()),
                                dst);
        };
        for (auto &kv : classes)
        {
            if (kv.first != 0)
            {
                op->remote.push_back(std::move(kv.second));
                SH_TRY(make_class_op(op->remote.back(), &op->remote.back().op));
                continue;
            }
            SH_TRY(make_class_op(kv.second, &op->local_op));
        }
        *out = op.release();
        return OK;
    }

    int fp_sharded_op_destroy(fp_sharded_op *op)
    {
        if (!op)
            return OK;
        fp_op_destroy(op->local_op);
        for (auto &cl : op->remote)
            fp_op_destroy(cl.op);
        if (op->t0)
            cudaEventDestroy(op->t0);
        if (op->t1)
            cudaEventDestroy(op->t1);
        delete op;
        return OK;
    }

    int fp_sharded_op_set_mode(fp_sharded_op *op, int mode)
    {
        if (!op || mode < 0 || mode > 2)
            return fail(INVALID, "mode must be 0 (auto), 1 (chunked) or 2 (whole shard)");
        op->mode = mode;
        return OK;
    }

    int fp_sharded_op_set_chunk_bytes(fp_sharded_op *op, size_t bytes)
    {
        if (!op)
            return fail(INVALID, "null pointer");
        op->chunk_bytes = bytes ? bytes : (256ull << 20);
        return OK;
    }

    int fp_sharded_op_last_mode(const fp_sharded_op *op, int *mode)
    {
        if (!op || !mode)
            return fail(INVALID, "null pointer");
        *mode = op->last_mode;
        return OK;
    }

    int fp_sharded_op_last_ms(const fp_sharded_op *op, float *ms)
    {
        if (!op || !ms)
            return fail(INVALID, "null pointer");
        *ms = op->last_ms;
        return OK;
    }

    // Pairwise exchange probe: every rank swaps `bytes` with rank ^ 1 through ncclSend / ncclRecv (the same calls the
    // sharded apply uses) `iters` times; returns GB/s per direction per GPU (device time, slowest rank's own clock).
    int fp_comm_measure_p2p(fp_comm *c, size_t bytes, int iters, double *gbps)
    {
        if (!c || !gbps)
            return fail(INVALID, "null pointer");
        if (c->world < 2)
            return fail(INVALID, "needs at least two ranks");
        DeviceScope scope(c->device);
        void *a = nullptr, *b = nullptr;
        SH_CU(cudaMalloc(&a, bytes));
        SH_CU(cudaMalloc(&b, bytes));
        SH_CU(cudaMemsetAsync(a, 1, bytes, c->comm_stream));
        cudaEvent_t e0, e1;
        SH_CU(cudaEventCreate(&e0));
        SH_CU(cudaEventCreate(&e1));
        int const peer = c->rank ^ 1;
        Nccl &n = nccl();
        for (int it = -1; it < iters; ++it)
        {
            if (it == 0)
                SH_CU(cudaEventRecord(e0, c->comm_stream));
            SH_NCCL(n.GroupStart());
            SH_NCCL(n.Send(a, bytes, ncclChar, peer, c->comm, c->comm_stream));
            SH_NCCL(n.Recv(b, bytes, ncclChar, peer, c->comm, c->comm_stream));
            SH_NCCL(n.GroupEnd());
        }
        SH_CU(cudaEventRecord(e1, c->comm_stream));
        SH_CU(cudaEventSynchronize(e1));
        float ms = 0;
        SH_CU(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(a);
        cudaFree(b);
        double t = ms;
        SH_TRY(fp_comm_allreduce_f64(c, &t, 1, 1));
        *gbps = static_cast<double>(bytes) * iters / (t * 1e-3) / 1e9;
        return OK;
    }

    int fp_sharded_op_info(const fp_sharded_op *op, size_t *n_remote_classes, uint64_t *peer_offsets /* >= world */,
                           uint64_t *bytes_sent_last, uint64_t *chunks_last, uint64_t *kernels_last)
    {
        if (!op)
            return fail(INVALID, "null pointer");
        if (n_remote_classes)
            *n_remote_classes = op->remote.size();
        if (peer_offsets)
            for (size_t i = 0; i < op->remote.size(); ++i)
                peer_offsets[i] = op->remote[i].x_hi;
        if (bytes_sent_last)
            *bytes_sent_last = op->last_bytes_sent;
        if (chunks_last)
            *chunks_last = op->last_chunks;
        if (kernels_last)
            *kernels_last = op->last_kernels;
        return OK;
    }

    static int sharded_apply(fp_sharded_op *op, void *out, const void *in, const void *all_shards, size_t local_dim,
                             size_t n_states, int accumulate)
    {
        if (!op || !out || !in)
            return fail(INVALID, "null pointer");
        fp_comm *c = op->comm;
        if (c->emulated != (all_shards != nullptr))
            return fail(INVALID, c->emulated ? "an emulated communicator needs fp_sharded_op_apply_emulated"
                                             : "fp_sharded_op_apply_emulated needs an emulated communicator");
        if (local_dim != (1ull << op->n_local)) // PO:343-346 on the shard
            return fail(INVALID, "[PauliOp] state size must match the dimension of the operators (local shard: 2^" +
                                     std::to_string(op->n_local) + " rows)");
        if (n_states == 0)
            return OK;
        cudaPointerAttributes ai{}, ao{};
        if (cudaPointerGetAttributes(&ai, in) != cudaSuccess || cudaPointerGetAttributes(&ao, out) != cudaSuccess ||
            ai.type != cudaMemoryTypeDevice || ao.type != cudaMemoryTypeDevice)
        {
            (void)cudaGetLastError();
            return fail(INVALID, "fp_sharded_op_apply takes device pointers (a shard lives on its GPU)");
        }
        DeviceScope scope(c->device);
        SH_TRY(fp_ctx_sync(c->user_ctx)); // whatever produced `in` on the caller's context has finished
        size_t const esize = op->dtype == FP_C128 ? 16 : 8;
        size_t const row_bytes = n_states * esize;
        op->last_bytes_sent = op->last_chunks = op->last_kernels = 0;
        if (!op->t0)
        {
            SH_CU(cudaEventCreate(&op->t0));
            SH_CU(cudaEventCreate(&op->t1));
        }
        SH_CU(cudaEventRecord(op->t0, c->compute));

        // ---- the local class first: it overlaps with the first exchanges
        if (op->local_op)
            SH_TRY(fp_op_apply(c->ctx, op->local_op, out, in, local_dim, n_states, accumulate));
        else if (!accumulate)
            SH_CU(cudaMemsetAsync(out, 0, local_dim * row_bytes, c->compute)); // every remote class accumulates

        // ---- whole-shard mode: the peer's shard is received into a shard-sized buffer (two when memory allows, so
        // the next class' exchange overlaps this class' kernels) and the class runs as ONE fused PauliOp -- measured
        // 3.6x faster than per-string streaming at 8 strings per class (31 local qubits); costs 1-2 shards of memory
        size_t const shard_bytes = local_dim * row_bytes;
        bool whole = false;
        int n_whole_bufs = 0;
        if (!op->remote.empty() && op->mode != 1)
        {
            if (c->buf_bytes >= shard_bytes)
                n_whole_bufs = 2;
            else
            {
                size_t free_b = 0, total_b = 0;
                SH_CU(cudaMemGetInfo(&free_b, &total_b));
                free_b += 2 * c->buf_bytes;
                size_t const margin = 2ull << 30; // the fused kernels' own scratch
                n_whole_bufs = free_b >= 2 * shard_bytes + margin ? 2 : (free_b >= shard_bytes + margin ? 1 : 0);
                if (op->mode == 0 && op->remote.size() == 1 && n_whole_bufs == 2)
                    n_whole_bufs = 1;
            }
            // the exchange schedule is collective: every rank must pick the same mode, so the ranks agree on the
            // smallest buffer count any of them can afford
            if (!c->emulated && c->world > 1)
            {
                double v = n_whole_bufs;
                SH_TRY(fp_comm_allreduce_f64(c, &v, 1, 2));
                n_whole_bufs = static_cast<int>(v);
            }
            whole = n_whole_bufs > 0 && (op->mode == 2 || shard_bytes > op->chunk_bytes);
            if (op->mode == 2 && n_whole_bufs == 0)
                return fail(4, "whole-shard mode: not enough device memory for a shard-sized receive buffer");
        }
        op->last_mode = whole ? 2 : 1;
        if (whole)
        {
            if (c->buf_bytes < shard_bytes)
            {
                SH_CU(cudaStreamSynchronize(c->compute));
                for (int i = 0; i < 2; ++i)
                {
                    cudaFree(c->bufs[i]);
                    c->bufs[i] = nullptr;
                }
                c->buf_bytes = 0;
                for (int i = 0; i < n_whole_bufs; ++i)
                    SH_CU(cudaMalloc(&c->bufs[i], shard_bytes));
                if (n_whole_bufs == 2)
                    c->buf_bytes = shard_bytes;
            }
            int const nb = c->bufs[1] ? 2 : 1;
            size_t const count = local_dim * n_states * 2;
            ncclDataType_t const ndt = op->dtype == FP_C128 ? ncclDouble : ncclFloat;
            static Nccl unused_w;
            Nccl &n = c->emulated ? unused_w : nccl();
            uint64_t t = 0;
            for (ShardedClass const &cl : op->remote)
            {
                int const peer = c->rank ^ static_cast<int>(cl.x_hi);
                int const b = static_cast<int>(t % nb);
                if (t >= static_cast<uint64_t>(nb))
                    SH_CU(cudaStreamWaitEvent(c->comm_stream, c->done[b], 0));
                if (c->emulated)
                    SH_CU(cudaMemcpyAsync(c->bufs[b], static_cast<unsigned char const *>(all_shards) + static_cast<size_t>(peer) * shard_bytes,
                                          shard_bytes, cudaMemcpyDeviceToDevice, c->comm_stream));
                else
                {
                    SH_NCCL(n.GroupStart());
                    SH_NCCL(n.Send(in, count, ndt, peer, c->comm, c->comm_stream));
                    SH_NCCL(n.Recv(c->bufs[b], count, ndt, peer, c->comm, c->comm_stream));
                    SH_NCCL(n.GroupEnd());
                }
                SH_CU(cudaEventRecord(c->ready[b], c->comm_stream));
                SH_CU(cudaStreamWaitEvent(c->compute, c->ready[b], 0));
                SH_TRY(fp_op_apply(c->ctx, cl.op, out, c->bufs[b], local_dim, n_states, 1));
                SH_CU(cudaEventRecord(c->done[b], c->compute));
                op->last_bytes_sent += shard_bytes;
                op->last_chunks++;
                op->last_kernels++;
                ++t;
            }
            if (nb == 1 && c->buf_bytes == 0)
            {
                // a single odd-sized buffer is not kept between calls
                SH_CU(cudaStreamSynchronize(c->compute));
                cudaFree(c->bufs[0]);
                c->bufs[0] = nullptr;
            }
        }
        else if (!op->remote.empty())
        {
            // chunk = 2^m rows, about chunk_bytes
            int m = op->n_local;
            while (m > 0 && (row_bytes << m) > op->chunk_bytes)
                --m;
            uint64_t const C = 1ull << m;
            uint64_t const n_chunks = local_dim >> m;
            size_t const cbytes = C * row_bytes;
            if (c->buf_bytes < cbytes)
            {
                SH_CU(cudaStreamSynchronize(c->compute));
                for (int i = 0; i < 2; ++i)
                {
                    cudaFree(c->bufs[i]);
                    c->bufs[i] = nullptr;
                    SH_CU(cudaMalloc(&c->bufs[i], cbytes));
                }
                c->buf_bytes = cbytes;
            }
            size_t const count = C * n_states * 2; // real scalars per chunk
            ncclDataType_t const ndt = op->dtype == FP_C128 ? ncclDouble : ncclFloat;
            static Nccl unused;
            Nccl &n = c->emulated ? unused : nccl();
            static std::complex<double> const phase[4] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}};
            uint64_t t = 0; // global chunk counter: buffer t & 1
            for (ShardedClass const &cl : op->remote)
            {
                int const peer = c->rank ^ static_cast<int>(cl.x_hi);
                for (uint64_t k = 0; k < n_chunks; ++k, ++t)
                {
                    int const b = static_cast<int>(t & 1);
                    if (t >= 2)
                        SH_CU(cudaStreamWaitEvent(c->comm_stream, c->done[b], 0)); // kernels on chunk t-2 are done
                    auto const *src = static_cast<unsigned char const *>(in) + k * cbytes;
                    if (c->emulated)
                    {
                        auto const *peer_chunk = static_cast<unsigned char const *>(all_shards) +
                                                 static_cast<size_t>(peer) * local_dim * row_bytes + k * cbytes;
                        SH_CU(cudaMemcpyAsync(c->bufs[b], peer_chunk, cbytes, cudaMemcpyDeviceToDevice, c->comm_stream));
                    }
                    else
                    {
                        SH_NCCL(n.GroupStart());
                        SH_NCCL(n.Send(src, count, ndt, peer, c->comm, c->comm_stream));
                        SH_NCCL(n.Recv(c->bufs[b], count, ndt, peer, c->comm, c->comm_stream));
                        SH_NCCL(n.GroupEnd());
                    }
                    SH_CU(cudaEventRecord(c->ready[b], c->comm_stream));
                    SH_CU(cudaStreamWaitEvent(c->compute, c->ready[b], 0));
                    op->last_bytes_sent += cbytes;
                    op->last_chunks++;
                    for (ShardedString const &st : cl.strings)
                    {
                        uint64_t const xh = st.x_lo >> m, zh = st.z_lo >> m;
                        uint64_t const j = k ^ xh; // output block fed by source chunk k
                        // phase of the characters above the chunk (the kernel adds the phase of the low m characters)
                        int ny_mid = 0;
                        for (int q = 0; q < op->n_local - m; ++q)
                            ny_mid += st.low[q] == 2;
                        double const sign = (__builtin_popcountll(j & zh) & 1) ? -1.0 : 1.0;
                        std::complex<double> const cd = st.c * phase[ny_mid & 3] * sign;
                        std::complex<float> const cf(cd);
                        auto *dst = static_cast<unsigned char *>(out) + j * cbytes;
                        SH_TRY(fp_string_apply(c->ctx, op->dtype, m, st.low.data() + (op->n_local - m),
                                               op->dtype == FP_C128 ? static_cast<void const *>(&cd) : static_cast<void const *>(&cf),
                                               dst, c->bufs[b], C, n_states, 1));
                        op->last_kernels++;
                    }
                    SH_CU(cudaEventRecord(c->done[b], c->compute));
                }
            }
        }
        SH_CU(cudaStreamSynchronize(c->comm_stream));
        SH_CU(cudaEventRecord(op->t1, c->compute));
        SH_CU(cudaStreamSynchronize(c->compute));
        SH_CU(cudaEventElapsedTime(&op->last_ms, op->t0, op->t1));
        return OK;
    }

    int fp_sharded_op_apply(fp_sharded_op *op, void *out, const void *in, size_t local_dim, size_t n_states,
                            int accumulate)
    {
        return sharded_apply(op, out, in, nullptr, local_dim, n_states, accumulate);
    }

    int fp_sharded_op_apply_emulated(fp_sharded_op *op, void *out, const void *all_shards, size_t local_dim,
                                     size_t n_states, int accumulate)
    {
        if (!op || !all_shards)
            return fail(INVALID, "null pointer");
        size_t const esize = op->dtype == FP_C128 ? 16 : 8;
        auto const *in = static_cast<unsigned char const *>(all_shards) +
                         static_cast<size_t>(op->comm->rank) * local_dim * n_states * esize;
        return sharded_apply(op, out, in, all_shards, local_dim, n_states, accumulate);
    }

    int fp_sharded_op_expval(fp_sharded_op *op, void *out_host, const void *in, void *work, size_t local_dim,
                             size_t n_states)
    {
        // <psi|A|psi> per column = sum over ranks of sum_i conj(psi_r[i]) (A psi)_r[i]: apply into `work` (a shard-sized
        // device buffer the caller provides), a local fused dot product per rank, one all-reduce.
        if (!op || !out_host || !in || !work)
            return fail(INVALID, "null pointer");
        if (n_states > 32)
            return fail(INVALID, "fp_sharded_op_expval: at most 32 columns");
        SH_TRY(fp_sharded_op_apply(op, work, in, local_dim, n_states, 0));
        fp_comm *c = op->comm;
        DeviceScope scope(c->device);
        // the identity-string PauliOp with a foreign bra: e[t] = sum_i conj(in(i,t)) * work(i,t)
        std::vector<uint8_t> ident(static_cast<size_t>(std::max(op->n_local, 1)), 0);
        std::complex<double> const one_d(1, 0);
        std::complex<float> const one_f(1, 0);
        fp_op *id_op = nullptr;
        SH_TRY(fp_op_create(c->ctx, op->dtype, op->n_local, 1, ident.data(),
                            op->dtype == FP_C128 ? static_cast<void const *>(&one_d) : static_cast<void const *>(&one_f), &id_op));
        void *dev_e = nullptr;
        size_t const esize = op->dtype == FP_C128 ? 16 : 8;
        SH_CU(cudaMalloc(&dev_e, n_states * esize));
        int rc = fp_op_expval_bra(c->ctx, id_op, dev_e, in, work, local_dim, n_states, 0);
        std::vector<double> vals(2 * n_states);
        if (rc == 0)
        {
            cudaStreamSynchronize(c->compute);
            if (op->dtype == FP_C128)
                cudaMemcpy(vals.data(), dev_e, n_states * 16, cudaMemcpyDeviceToHost);
            else
            {
                std::vector<float> f(2 * n_states);
                cudaMemcpy(f.data(), dev_e, n_states * 8, cudaMemcpyDeviceToHost);
                for (size_t i = 0; i < f.size(); ++i)
                    vals[i] = f[i];
            }
        }
        cudaFree(dev_e);
        fp_op_destroy(id_op);
        SH_TRY(rc);
        SH_TRY(fp_comm_allreduce_f64(c, vals.data(), vals.size(), 0));
        if (op->dtype == FP_C128)
            std::memcpy(out_host, vals.data(), n_states * 16);
        else
            for (size_t i = 0; i < vals.size(); ++i)
                static_cast<float *>(out_host)[i] = static_cast<float>(vals[i]);
        return OK;
    }
} // extern "C"
