// C++ front-end of the sharded-state path (BASELINE config 5): one state vector split by its high index bits over
// the GPUs of a box, one process per GPU.  Thin RAII over the C ABI (fp_comm_* / fp_sharded_op_*, csrc/sharded.cpp):
// NCCL lives inside the library, the host needs neither torch nor MPI -- only a way to hand the 128-byte id of rank 0
// to the other ranks.  Nothing like this exists in the reference (its dim is an int shift, __pauli_string.hpp:57);
// the arithmetic being distributed is PauliOp::apply / expectation_value (__pauli_op.hpp:362-549).
//
//     std::array<unsigned char, FP_COMM_ID_BYTES> id = rank == 0 ? fast_pauli::ShardedComm::unique_id() : recv_id();
//     fast_pauli::ShardedComm comm(id, world, rank);                 // collective
//     fast_pauli::ShardedPauliOp<double> op(comm, pauli_op);         // any fast_pauli::PauliOp<double>
//     op.apply(out_shard_device, in_shard_device);                   // collective; device pointers to dim/world rows
#pragma once
#include <array>
#include <complex>
#include <vector>

#include "detail.hpp"
#include "pauli_op.hpp"

namespace fast_pauli
{

class ShardedComm
{
  public:
    static std::array<unsigned char, FP_COMM_ID_BYTES> unique_id()
    {
        std::array<unsigned char, FP_COMM_ID_BYTES> id{};
        gpu::check(fp_comm_unique_id(id.data()));
        return id;
    }
    ShardedComm(std::array<unsigned char, FP_COMM_ID_BYTES> const &id, int world, int rank, fp_ctx *ctx = nullptr)
    {
        gpu::check(fp_comm_create(ctx ? ctx : gpu::context(), id.data(), world, rank, &comm_));
        world_ = world;
        rank_ = rank;
    }
    ShardedComm(ShardedComm const &) = delete;
    ShardedComm &operator=(ShardedComm const &) = delete;
    ~ShardedComm()
    {
        fp_comm_destroy(comm_);
    }
    int world() const
    {
        return world_;
    }
    int rank() const
    {
        return rank_;
    }
    void barrier()
    {
        gpu::check(fp_comm_barrier(comm_));
    }
    fp_comm *handle() const
    {
        return comm_;
    }

  private:
    fp_comm *comm_ = nullptr;
    int world_ = 1, rank_ = 0;
};

template <std::floating_point T> class ShardedPauliOp
{
  public:
    ShardedPauliOp(ShardedComm &comm, PauliOp<T, std::complex<T>> const &op) : n_qubits_(op.n_qubits())
    {
        size_t const S = op.n_pauli_strings(), n = op.n_qubits();
        std::vector<uint8_t> codes(S * n);
        for (size_t s = 0; s < S; ++s)
            for (size_t q = 0; q < n; ++q)
                codes[s * n + q] = op.pauli_strings[s].paulis[q].code;
        gpu::check(fp_sharded_op_create(comm.handle(), gpu::dtype_of<T>(), static_cast<int>(n), S, codes.data(),
                                        op.coeffs.data(), &op_));
        int lg = 0;
        while ((1 << lg) < comm.world())
            ++lg;
        local_dim_ = size_t(1) << (n - lg);
    }
    ShardedPauliOp(ShardedPauliOp const &) = delete;
    ShardedPauliOp &operator=(ShardedPauliOp const &) = delete;
    ~ShardedPauliOp()
    {
        fp_sharded_op_destroy(op_);
    }
    size_t local_dim() const
    {
        return local_dim_;
    }
    // out_shard (+)= (A psi)_shard; DEVICE pointers to local_dim() x n_states row-major complex<T>
    void apply(std::complex<T> *out_shard, std::complex<T> const *in_shard, size_t n_states = 1, bool accumulate = false)
    {
        gpu::check(fp_sharded_op_apply(op_, out_shard, in_shard, local_dim_, n_states, accumulate ? 1 : 0));
    }
    // <psi|A|psi> per column, the same values on every rank; `work` is a shard-sized device scratch buffer
    std::vector<std::complex<T>> expectation_value(std::complex<T> const *in_shard, std::complex<T> *work,
                                                   size_t n_states = 1)
    {
        std::vector<std::complex<T>> out(n_states);
        gpu::check(fp_sharded_op_expval(op_, out.data(), in_shard, work, local_dim_, n_states));
        return out;
    }
    float last_device_ms() const
    {
        float ms = 0;
        gpu::check(fp_sharded_op_last_ms(op_, &ms));
        return ms;
    }

  private:
    fp_sharded_op *op_ = nullptr;
    size_t n_qubits_ = 0, local_dim_ = 0;
};

} // namespace fast_pauli
