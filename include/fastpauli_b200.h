/*
 * fastpauli_b200 -- C ABI of the B200-native fast-pauli hot path.
 *
 * This is the drop-in boundary: the reference (qognitive/fast-pauli) has no
 * FFI of its own -- its hot path is nine C++ template methods -- so each entry
 * point below names the reference method whose body it replaces
 * (paths relative to fast_pauli/cpp/include/ in the reference tree):
 *
 *   PS  = __pauli_string.hpp   PO = __pauli_op.hpp   SPO = __summed_pauli_op.hpp
 *
 * Conventions (identical to the reference's mdspan arguments):
 *   - state batches are row-major (dim, n_states) with the batch axis
 *     contiguous ("transposed" layout, PS:361-363), interleaved (re, im);
 *     dtype FP_C128 = std::complex<double>, FP_C64 = std::complex<float>;
 *   - Pauli strings travel as uint8 codes, n_strings x n_qubits, 0:I 1:X 2:Y
 *     3:Z; codes[s*n + 0] is the left-most character = most significant qubit
 *     (PS:52-54);
 *   - every data pointer may be a HOST pointer (pageable or pinned: staged
 *     through device scratch inside the call) or a DEVICE pointer on the
 *     context's GPU (used in place); the library never keeps caller pointers;
 *   - `accumulate` != 0 reproduces the C++ methods' `+=` into the caller's
 *     buffer (PS:419,432,523,534); 0 overwrites (what the reference's Python
 *     bindings observe, because they hand in zeroed outputs, NB:149-172);
 *   - calls are synchronous: they return after the work on the context's
 *     stream has completed, unless fp_ctx_set_async(ctx, 1) was called, in
 *     which case calls whose pointers are all device pointers only enqueue;
 *   - return value: FP_OK or an fp_status code; fp_last_error() holds the
 *     message (thread-local).  FP_INVALID_ARGUMENT is raised exactly where
 *     the reference throws std::invalid_argument.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with FP_NO_DEVICE.
 */
#ifndef FASTPAULI_B200_H
#define FASTPAULI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef enum fp_status
    {
        FP_OK = 0,
        FP_INVALID_ARGUMENT = 1, /* the reference would throw std::invalid_argument */
        FP_CUDA_ERROR = 2,       /* a CUDA runtime call failed */
        FP_NO_DEVICE = 3,        /* no usable CUDA device */
        FP_OUT_OF_MEMORY = 4,
        FP_UNSUPPORTED = 5
    } fp_status;

    typedef enum fp_dtype
    {
        FP_C64 = 0, /* std::complex<float>  */
        FP_C128 = 1 /* std::complex<double> */
    } fp_dtype;

    typedef struct fp_ctx fp_ctx;     /* one GPU + stream + scratch */
    typedef struct fp_op fp_op;       /* device-resident packed PauliOp (strings grouped by x-mask) */
    typedef struct fp_sop fp_sop;     /* device-resident packed SummedPauliOp */
    typedef struct fp_event fp_event; /* CUDA event for timing on the context's stream */

    /* ---- library / context ------------------------------------------------------------------ */
    const char *fp_last_error(void);
    int fp_version(void);
    int fp_device_count(int *count);
    int fp_ctx_create(int device, fp_ctx **ctx);
    int fp_ctx_destroy(fp_ctx *ctx);
    int fp_ctx_device(const fp_ctx *ctx, int *device);
    /* external != 0: run on exactly this caller-owned cudaStream_t (e.g. PyTorch's current stream; 0 / NULL is the
     * legacy default stream); external == 0 restores the context's own non-blocking stream. */
    /* PCI bus id of the context's GPU ("0000:1b:00.0"; buf >= 16 bytes): lets a multi-GPU host place each process and
     * its pinned buffers on the NUMA node its GPU hangs off (/sys/bus/pci/devices/<id>/numa_node). */
    int fp_ctx_pci_bus_id(const fp_ctx *ctx, char *buf, int len);
    int fp_ctx_set_stream(fp_ctx *ctx, void *cuda_stream, int external);
    int fp_ctx_set_async(fp_ctx *ctx, int async);
    int fp_ctx_sync(fp_ctx *ctx);
    /* Measured FP64 FMA throughput (TFLOP/s) of the context's GPU: a short DFMA microkernel on every SM, best of 3.
     * Measurement aid for the roofline of compute-bound operators (no reference counterpart). */
    int fp_measure_fp64_tflops(fp_ctx *ctx, double *tflops);
    /* Number of kernels this context has launched so far (bench.py's gpu_launches). */
    int fp_ctx_launch_count(const fp_ctx *ctx, uint64_t *count);
    /* Pinned (page-locked) HOST buffers passed to the single-string entry points are read and written in place by
     * the kernel over PCIe (both directions overlap inside one launch) instead of being staged; 0 disables. */
    int fp_ctx_set_zero_copy(fp_ctx *ctx, int enable);
    /* Coset-blocked (state tile in shared memory) kernels for multi-x-mask operators: mode 0 = never, 1 = heuristic
     * (default), 2 = whenever applicable.  log_twc >= 0 forces the row-segment width of the tile (2^log_twc 16-byte
     * vectors, 0..4; -1 = automatic); log_nt = 7 or 8 forces 128- or 256-thread CTAs (0 = default).  A tile holds
     * 16 * 2^log_nt vectors, i.e. 2^(4 + log_nt - log_twc) rows. */
    int fp_ctx_set_coset(fp_ctx *ctx, int mode, int log_twc, int log_nt);
    /* fp_string_apply with HOST in/out buffers of at least min_bytes (default 128 MiB; accumulate = 0) streams the
     * batch through the GPU in aligned row blocks of about chunk_bytes (default 32 MiB): upload of block j+1, the
     * kernel on block j and download of block j-1 overlap, so both PCIe directions are busy for the whole call.
     * enable = 0 restores the single-shot path; 0 for a size keeps the current value. */
    int fp_ctx_set_pipeline(fp_ctx *ctx, int enable, size_t min_bytes, size_t chunk_bytes);
    /* Register-resident coset kernels (operators whose x-masks span a GF(2) subspace of rank <= 4: each thread holds
     * the <= 16 rows of one coset for one 16-byte vector, no shared-memory staging): mode 0 = never, 1 = automatic
     * (default; stands aside when fp_ctx_set_coset forces a mode or shape), 2 = whenever applicable.  log_nt = 7 or 8
     * picks 128- or 256-thread CTAs (0 = default, 128). */
    int fp_ctx_set_rcoset(fp_ctx *ctx, int mode, int log_nt);
    /* Passes of a coset plan with at most 8 x-masks (K3e / K3f, csrc/coset2.cuh: row factors in registers, strings in
     * the constant bank, TMA-fed persistent variant), passes whose x-masks carry one string each (K3i,
     * csrc/coset3.cuh: TMA-fed, direct stores) and passes of eight independent x-masks (K3j, csrc/coset4.cuh: direct
     * stores, paired masks, row-factor table): mode 0 = never (such passes run on the general coset kernel),
     * 1 = automatic (default), 2 = never the TMA-fed kernels, 3 = automatic without K3i and K3j, 4 = automatic
     * without K3j, 5 = automatic with K3j also on single-string masks (by default those stay on K3i).  column_tiles_per_cta > 0
     * forces how many column tiles one CTA of K3e walks (0 = automatic). */
    int fp_ctx_set_coset_few(fp_ctx *ctx, int mode, int column_tiles_per_cta);
    /* Which kernels of the coset family ran on this context since the last reset (bit mask: 1 = K3b coset_kernel,
     * 2 = K3e coset_few_kernel, 4 = K3f coset_few_tma_kernel, 8 = K3g coset_gen_tma_kernel, 16 = K3i
     * coset_dir_tma_kernel, 32 = K3j coset_pair_tma_kernel); reset != 0 clears it.  For tests and benchmarks that assert the launch path. */
    int fp_ctx_coset_kernels_used(fp_ctx *ctx, uint32_t *mask, int reset);
    /* Override the L2 working-set budget (bytes) used to pick the batch-tile width of multi-group kernels. */
    int fp_ctx_set_l2_budget(fp_ctx *ctx, size_t bytes);

    /* ---- memory / timing helpers (thin wrappers so hosts need not link libcudart) ------------- */
    int fp_device_malloc(fp_ctx *ctx, size_t bytes, void **ptr);
    int fp_device_free(fp_ctx *ctx, void *ptr);
    int fp_host_malloc(fp_ctx *ctx, size_t bytes, void **ptr); /* pinned */
    int fp_host_free(fp_ctx *ctx, void *ptr);
    int fp_memcpy(fp_ctx *ctx, void *dst, const void *src, size_t bytes);       /* any direction, synchronous */
    int fp_memcpy_async(fp_ctx *ctx, void *dst, const void *src, size_t bytes); /* enqueued on the ctx stream */
    int fp_memset(fp_ctx *ctx, void *dst, int value, size_t bytes);
    int fp_device_mem_info(fp_ctx *ctx, size_t *free_bytes, size_t *total_bytes);
    int fp_event_create(fp_event **ev);
    int fp_event_destroy(fp_event *ev);
    int fp_event_record(fp_ctx *ctx, fp_event *ev);
    int fp_event_elapsed_ms(fp_event *start, fp_event *stop, float *ms); /* synchronises on `stop` */
    /* Fill a device or host buffer with the bench's counter-based U[0,1) amplitudes:
     * element e (a real scalar; 2 per complex) = splitmix64(seed, first_elem + e) -> [0,1).  On-device generation. */
    int fp_fill_uniform(fp_ctx *ctx, int dtype, void *dst, uint64_t n_complex, uint64_t first_complex, uint64_t seed);

    /* ---- PauliString (one-shot: the string is three words, no plan needed) -------------------- */
    /* PauliString::apply, 1-D (PS:296-341) when n_states == 1 and PauliString::apply_batch (PS:377-436):
     *   out(i,t) (+)= coeff * m[i] * in(i ^ x, t).  Errors: PS:271-283, PS:343-359. */
    int fp_string_apply(fp_ctx *ctx, int dtype, int n_qubits, const uint8_t *codes, const void *coeff /* 1 complex */,
                        void *out, const void *in, size_t dim, size_t n_states, int accumulate);
    /* PauliString::expectation_value (PS:470-538): out[t] (+)= sum_i conj(in(i,t)) coeff m[i] in(i^x,t).
     * Errors: PS:438-450. */
    int fp_string_expval(fp_ctx *ctx, int dtype, int n_qubits, const uint8_t *codes, const void *coeff,
                         void *out /* n_states complex */, const void *in, size_t dim, size_t n_states,
                         int accumulate);

    /* ---- PauliOp (PO:38-97): sum_s h_s P_s ----------------------------------------------------- */
    /* Pack (x,z,phase) per string, merge duplicates, group by x-mask, upload.  coeffs: n_strings complex of `dtype`.
     * Errors: unequal string sizes cannot occur in this encoding; codes > 3 -> P:58-59. */
    int fp_op_create(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes,
                     const void *coeffs, fp_op **op);
    int fp_op_destroy(fp_op *op);
    int fp_op_info(const fp_op *op, int *dtype, int *n_qubits, size_t *n_strings, size_t *n_packed_strings,
                   size_t *n_groups);
    /* PauliOp::apply 1-D (PO:362-383, n_states == 1) and 2-D (PO:399-468). Errors: PO:340-351. */
    int fp_op_apply(fp_ctx *ctx, const fp_op *op, void *out, const void *in, size_t dim, size_t n_states,
                    int accumulate);
    /* PauliOp::expectation_value (PO:482-549). Errors: PO:502-505. */
    int fp_op_expval(fp_ctx *ctx, const fp_op *op, void *out /* n_states complex */, const void *in, size_t dim,
                     size_t n_states, int accumulate);

    /* Sharded-state building block (no reference counterpart): out[t] (+)= sum_i conj(bra(i,t)) (A in)(i,t) with
     * bra != in, i.e. the local bra shard against a peer's ket shard.  Device pointers only. */
    int fp_op_expval_bra(fp_ctx *ctx, const fp_op *op, void *out, const void *bra, const void *in, size_t dim,
                         size_t n_states, int accumulate);

    /* ---- SummedPauliOp (SPO:37-145): K operators over one string set, coeffs (n_strings, n_operators) ---- */
    int fp_sop_create(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes,
                      size_t n_operators, const void *coeffs /* row-major (n_strings, n_operators) complex */,
                      fp_sop **sop);
    int fp_sop_destroy(fp_sop *sop);
    /* SummedPauliOp::apply (SPO:277-349). Errors: SPO:290-293 (+ the dim check the reference omits). */
    int fp_sop_apply(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, size_t dim, size_t n_states,
                     int accumulate);
    /* SummedPauliOp::apply_weighted (SPO:364-503): data is real (n_operators, n_states), float or double
     * independently of the state dtype.  Errors: SPO:384-399. */
    int fp_sop_apply_weighted(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, const void *data,
                              int data_is_f64, size_t dim, size_t n_states, int accumulate);
    /* SummedPauliOp::expectation_value (SPO:520-614): out is (n_operators, n_states) complex. Errors: SPO:539-558. */
    int fp_sop_expval(fp_ctx *ctx, const fp_sop *sop, void *out, const void *in, size_t dim, size_t n_states,
                      int accumulate);
    /* SummedPauliOp::square() (SPO:197-268), the coefficient contraction of A_k -> A_k^2 on the device:
     * coeffs_sq(c, k) = sum over ordered pairs (a, b) with P_a P_b = phase P_c of phase * coeffs(a, k) * coeffs(b, k)
     * for every output string c of sq_codes (n_sq x n_qubits codes; the caller enumerates the reference's output set,
     * calculate_pauli_strings_max_weight(n, min(n, 2 * max weight)), in its own language).  coeffs is host memory
     * (n_strings, n_operators) row-major; coeffs_sq (n_sq, n_operators) may be host or device memory and is
     * overwritten.  Duplicate input strings are merged first (same operators). */
    int fp_sop_square(fp_ctx *ctx, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes, size_t n_operators,
                      const void *coeffs, size_t n_sq, const uint8_t *sq_codes, void *coeffs_sq);
    /* Select the coefficient-contraction engine of apply_weighted / expectation_value for FP_C64 plans:
     * 0 = FP32 SIMT, 1 = tcgen05 3xTF32 tensor-core path (default when available). */
    int fp_ctx_set_tensor_core(fp_ctx *ctx, int enable);

    /* ---- peer memory: one process per GPU, kernels read a peer GPU's shard directly over NVLink -----------------
     * fp_ipc_export wraps cudaIpcGetMemHandle for a buffer obtained from fp_device_malloc (the handle is
     * FP_IPC_HANDLE_BYTES opaque bytes to be shipped to the peer process, e.g. with torch.distributed);
     * fp_ipc_open maps the peer's buffer into this process (peer access is enabled lazily) and returns a device
     * pointer that every entry point accepts as `in`; fp_ipc_close unmaps it. */
#define FP_IPC_HANDLE_BYTES 64
    int fp_ipc_export(fp_ctx *ctx, const void *dev_ptr, unsigned char *handle);
    int fp_ipc_open(fp_ctx *ctx, const unsigned char *handle, void **peer_ptr);
    int fp_ipc_close(fp_ctx *ctx, void *peer_ptr);

    /* ---- sharded state (BASELINE config 5): one state vector split by its high index bits over the GPUs of a box ---
     * One process per GPU.  NCCL is used by the library itself (dlopen of libnccl.so.2, no link-time dependency, no
     * torch / MPI): rank 0 calls fp_comm_unique_id and ships the FP_COMM_ID_BYTES bytes to every rank by any means;
     * every rank then calls fp_comm_create (collective).  world must be a power of two; rank r owns the rows whose
     * top log2(world) index bits equal r, i.e. a contiguous shard of dim / world rows.
     * fp_sharded_op_apply: out_shard (+)= (A psi)_shard (PauliOp::apply 1-D / 2-D, PO:362-468, on the whole state):
     * strings whose x-mask leaves the high bits alone run as one fused local PauliOp; every other group of strings
     * (one per peer offset x_hi) streams the peer's shard in chunks (default 256 MiB, two receive buffers) with
     * ncclSend / ncclRecv inside ncclGroupStart / End on a communication stream while the previous chunk is applied
     * by the streaming single-string kernel.  Device pointers only; synchronous; collective (every rank calls it).
     * fp_sharded_op_expval: <psi|A|psi> per column (PO:482-549) = apply into `work` (a shard-sized device buffer of
     * the caller) + local <psi|work> + one all-reduce; out_host receives n_states complex on EVERY rank. */
#define FP_COMM_ID_BYTES 128
    typedef struct fp_comm fp_comm;
    typedef struct fp_sharded_op fp_sharded_op;
    int fp_comm_unique_id(unsigned char *id /* FP_COMM_ID_BYTES */);
    int fp_comm_create(fp_ctx *ctx, const unsigned char *id, int world, int rank, fp_comm **comm);
    /* Single-process stand-in for rank `rank` of `world` (tests on one GPU): no NCCL; the sharded apply then takes a
     * device buffer holding ALL shards back to back (fp_sharded_op_apply_emulated) and copies the peer chunks from it,
     * every other step -- classes, chunk schedule, block signs, kernels -- being the production code. */
    int fp_comm_create_emulated(fp_ctx *ctx, int world, int rank, fp_comm **comm);
    int fp_comm_destroy(fp_comm *comm);
    int fp_comm_info(const fp_comm *comm, int *world, int *rank, int *nccl_version);
    int fp_comm_barrier(fp_comm *comm);
    /* in-place sum (op = 0), max (op = 1) or min (op = 2) of n <= 64 host doubles over the ranks (timing, small
     * reductions) */
    int fp_comm_allreduce_f64(fp_comm *comm, double *values, size_t n, int op);
    /* every rank swaps `bytes` with rank ^ 1 through ncclSend / ncclRecv `iters` times: GB/s per direction per GPU
     * (device time of the slowest rank) -- the measured NVLink denominator of the sharded apply */
    int fp_comm_measure_p2p(fp_comm *comm, size_t bytes, int iters, double *gbps);
    int fp_sharded_op_create(fp_comm *comm, int dtype, int n_qubits, size_t n_strings, const uint8_t *codes,
                             const void *coeffs, fp_sharded_op **op);
    int fp_sharded_op_destroy(fp_sharded_op *op);
    int fp_sharded_op_set_chunk_bytes(fp_sharded_op *op, size_t bytes); /* 0 = default (256 MiB) */
    /* How a peer's shard reaches the kernels: 1 = chunked streaming (two chunk buffers; one streaming single-string
     * kernel per string and chunk), 2 = whole shard (one or two shard-sized receive buffers; each group of strings
     * runs as ONE fused PauliOp on the received shard while the next group's exchange is in flight), 0 = automatic:
     * whole shard when the device has room for it (it is 2-4x faster with several strings per peer offset). */
    int fp_sharded_op_set_mode(fp_sharded_op *op, int mode);
    int fp_sharded_op_apply(fp_sharded_op *op, void *out_shard, const void *in_shard, size_t local_dim, size_t n_states,
                            int accumulate);
    int fp_sharded_op_expval(fp_sharded_op *op, void *out_host, const void *in_shard, void *work_shard, size_t local_dim,
                             size_t n_states);
    int fp_sharded_op_apply_emulated(fp_sharded_op *op, void *out_shard, const void *all_shards, size_t local_dim,
                                     size_t n_states, int accumulate);
    /* remote classes (= peer offsets in use), and of the last apply: bytes sent per rank, chunks, kernels launched */
    int fp_sharded_op_info(const fp_sharded_op *op, size_t *n_remote_classes, uint64_t *peer_offsets,
                           uint64_t *bytes_sent_last, uint64_t *chunks_last, uint64_t *kernels_last);
    /* device time (ms) of this rank's last fp_sharded_op_apply: CUDA events around the call's compute stream */
    int fp_sharded_op_last_ms(const fp_sharded_op *op, float *ms);
    int fp_sharded_op_last_mode(const fp_sharded_op *op, int *mode); /* 1 chunked, 2 whole shard */

    /* ---- diagnostics ----------------------------------------------------------------------------
     * The contraction engine on its own: C[split_k][M x N] = A[M x Kd] * B[Kd x N] (row-major fp32, split-K planes
     * left unreduced), engine 0 = FP32 SIMT, 1 = tcgen05 3xTF32 (falls back to SIMT when the shape is unsupported;
     * fp_ctx_last_gemm_engine tells which one ran). */
    int fp_debug_gemm_f32(fp_ctx *ctx, int engine, const float *A, const float *B, float *C, uint32_t M, uint64_t N,
                          uint32_t Kd, uint32_t split_k);
    int fp_ctx_last_gemm_engine(const fp_ctx *ctx, int *engine);

    /* ---- one-shot entry points with the oracle's raw signature ---------------------------------
     * Same argument lists as orc_* (oracle/pauli_oracle.c) and ref_* (oracle/ref_wrapper.cpp) so one
     * harness drives the reference, the port and the GPU.  They run on a process-wide default context
     * (device from FASTPAULI_DEVICE, default 0), ACCUMULATE into `out` like the C++ methods, and accept
     * `par` only for signature compatibility: both execution policies route to the GPU. */
#define FP_DECLARE_ONESHOT(SFX, T)                                                                                     \
    int fp_string_apply1d_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, int par);    \
    int fp_string_apply_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, size_t B,      \
                              int par);                                                                                \
    int fp_string_expval_##SFX(int n, const uint8_t *codes, const T *c, T *out, const T *in, size_t dim, size_t B,     \
                               int par);                                                                               \
    int fp_op_apply1d_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,   \
                            int par);                                                                                  \
    int fp_op_apply_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,     \
                          size_t B, int par);                                                                          \
    int fp_op_expval_##SFX(int n, size_t S, const uint8_t *codes, const T *coeffs, T *out, const T *in, size_t dim,    \
                           size_t B, int par);                                                                         \
    int fp_sop_apply_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out, const T *in,      \
                           size_t dim, size_t B, int par);                                                             \
    int fp_sop_apply_weighted_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out,          \
                                    const T *in, const void *data, int data_is_f64, size_t dim, size_t B, int par);    \
    int fp_sop_expval_##SFX(int n, size_t S, const uint8_t *codes, size_t K, const T *coeffs, T *out, const T *in,     \
                            size_t dim, size_t B, int par);
    FP_DECLARE_ONESHOT(c128, double)
    FP_DECLARE_ONESHOT(c64, float)
#undef FP_DECLARE_ONESHOT
    /* The default context used by the one-shot entry points (created on first use). */
    int fp_default_ctx(fp_ctx **ctx);

#ifdef __cplusplus
}
#endif
#endif /* FASTPAULI_B200_H */
