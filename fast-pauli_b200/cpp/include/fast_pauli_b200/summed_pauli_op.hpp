// fast_pauli::SummedPauliOp<T> -- K operators A_k = sum_i h_ik P_i over one string set.
// Reference API being mirrored: __summed_pauli_op.hpp:37-666 (same members, overloads and exceptions).
// apply / apply_weighted / expectation_value / square run on the GPU; split / to_tensor are small host code.
// Unlike the reference (whose implicit copy leaves `coeffs` pointing into the source object, SPO:43-45) copies
// re-point the mdspan at their own buffer.
#pragma once
#include <cstring>
#include <unordered_map>

#include "helpers.hpp"
#include "pauli_op.hpp"

namespace fast_pauli
{

template <std::floating_point T> struct SummedPauliOp
{
    template <size_t N> using Tensor = std::mdspan<std::complex<T>, std::dextents<size_t, N>>;

    std::vector<PauliString> pauli_strings;
    std::vector<std::complex<T>> coeffs_raw;
    Tensor<2> coeffs; // (n_pauli_strings, n_operators)

    SummedPauliOp() noexcept = default;
    SummedPauliOp(std::vector<PauliString> const &strings, std::vector<std::complex<T>> const &flat_coeffs)
        : pauli_strings(strings), coeffs_raw(flat_coeffs)
    {
        if (pauli_strings.empty())
            throw std::invalid_argument("SummedPauliOp needs at least one PauliString");
        size_t const n_ops = coeffs_raw.size() / pauli_strings.size();
        coeffs = Tensor<2>(coeffs_raw.data(), pauli_strings.size(), n_ops);
        check_ctor(coeffs.extent(0));
    }
    SummedPauliOp(std::vector<PauliString> const &strings, Tensor<2> const coeffs_in) : pauli_strings(strings)
    {
        adopt(coeffs_in);
    }
    SummedPauliOp(std::vector<std::string> const &strings, Tensor<2> const coeffs_in)
    {
        pauli_strings.reserve(strings.size());
        for (auto const &s : strings)
            pauli_strings.emplace_back(s);
        adopt(coeffs_in);
    }
    SummedPauliOp(SummedPauliOp const &o) : pauli_strings(o.pauli_strings), coeffs_raw(o.coeffs_raw)
    {
        coeffs = Tensor<2>(coeffs_raw.data(), o.coeffs.extent(0), o.coeffs.extent(1));
    }
    SummedPauliOp &operator=(SummedPauliOp const &o)
    {
        if (this != &o)
        {
            pauli_strings = o.pauli_strings;
            coeffs_raw = o.coeffs_raw;
            coeffs = Tensor<2>(coeffs_raw.data(), o.coeffs.extent(0), o.coeffs.extent(1));
            cache_.reset();
        }
        return *this;
    }
    ~SummedPauliOp()
    {
        cache_.reset();
    }

    size_t dim() const noexcept
    {
        return pauli_strings.empty() ? 0 : pauli_strings.front().dim();
    }
    size_t n_qubits() const noexcept
    {
        return pauli_strings.empty() ? 0 : pauli_strings.front().n_qubits();
    }
    size_t n_operators() const noexcept
    {
        return coeffs.extent(1);
    }
    size_t n_pauli_strings() const noexcept
    {
        return pauli_strings.size();
    }

    // ---- apply (reference: SPO:277-349): new_states += sum_j (sum_k h_jk) P_j states
    void apply(Tensor<2> new_states, Tensor<2> states) const
    {
        apply(std::execution::seq, new_states, states);
    }
    template <execution_policy ExecutionPolicy> void apply(ExecutionPolicy &&, Tensor<2> new_states, Tensor<2> states) const
    {
        if (states.extent(0) != new_states.extent(0) || states.extent(1) != new_states.extent(1))
            throw std::invalid_argument("new_states must have the same dimensions as states");
        gpu::check(fp_sop_apply(gpu::context(), device_plan(), new_states.data_handle(), states.data_handle(),
                                states.extent(0), states.extent(1), /*accumulate=*/1));
    }

    // ---- apply_weighted (reference: SPO:364-503): new_states(l,t) += sum_j [sum_k h_jk x_kt] (P_j psi_t)(l)
    template <std::floating_point data_dtype>
    void apply_weighted(Tensor<2> new_states, Tensor<2> states,
                        std::mdspan<data_dtype, std::dextents<size_t, 2>> data) const
    {
        apply_weighted(std::execution::seq, new_states, states, data);
    }
    template <execution_policy ExecutionPolicy, std::floating_point data_dtype>
    void apply_weighted(ExecutionPolicy &&, Tensor<2> new_states, Tensor<2> states,
                        std::mdspan<data_dtype, std::dextents<size_t, 2>> data) const
    {
        static_assert(std::is_same_v<data_dtype, float> || std::is_same_v<data_dtype, double>);
        if (states.extent(0) != new_states.extent(0) || states.extent(1) != new_states.extent(1))
            throw std::invalid_argument("new_states must have the same dimensions as states");
        if (data.extent(0) != n_operators() || data.extent(1) != states.extent(1))
            throw std::invalid_argument("data(k,t) must have the same number of operators as the SummedPauliOp "
                                        "and the same number of states as the input states");
        if (states.extent(0) != dim())
            throw std::invalid_argument("state size must match the dimension of the operators");
        gpu::check(fp_sop_apply_weighted(gpu::context(), device_plan(), new_states.data_handle(), states.data_handle(),
                                         data.data_handle(), std::is_same_v<data_dtype, double>, states.extent(0),
                                         states.extent(1), /*accumulate=*/1));
    }

    // ---- expectation_value (reference: SPO:520-614): out(k,t) += sum_j h_jk <psi_t|P_j|psi_t>
    void expectation_value(Tensor<2> expectation_vals_out, Tensor<2> states) const
    {
        expectation_value(std::execution::seq, expectation_vals_out, states);
    }
    template <execution_policy ExecutionPolicy>
    void expectation_value(ExecutionPolicy &&, Tensor<2> expectation_vals_out, Tensor<2> states) const
    {
        if (expectation_vals_out.extent(0) != n_operators())
            throw std::invalid_argument("expectation_vals_out must have the same number of operators (" +
                                        std::to_string(expectation_vals_out.extent(0)) + ") as the SummedPauliOp (" +
                                        std::to_string(n_operators()) + ")");
        if (states.extent(0) != dim())
            throw std::invalid_argument("states must have the same dimension (" + std::to_string(states.extent(0)) +
                                        ") as the SummedPauliOp (" + std::to_string(dim()) + ")");
        if (expectation_vals_out.extent(1) != states.extent(1))
            throw std::invalid_argument("expectation_vals_out must have the same number of states (" +
                                        std::to_string(expectation_vals_out.extent(1)) + ") as the input states (" +
                                        std::to_string(states.extent(1)) + ")");
        gpu::check(fp_sop_expval(gpu::context(), device_plan(), expectation_vals_out.data_handle(),
                                 states.data_handle(), states.extent(0), states.extent(1), /*accumulate=*/1));
    }

    // ---- host-side helpers (not on the data-parallel path)
    // A_k -> A_k^2: coefficient of string c in operator k is sum_{a,b: P_a P_b ~ P_c} phase(a,b) h_ak h_bk
    // (reference: SPO:197-268).  The output string set is every string up to weight min(n, 2 * max weight),
    // in calculate_pauli_strings_max_weight order.
    SummedPauliOp<T> square() const
    {
        // output string set as in the reference (SPO:205-214); the T_aij contraction (SPO:216-265) runs on the GPU
        size_t max_w = 0;
        for (auto const &ps : pauli_strings)
            max_w = std::max<size_t>(max_w, ps.weight);
        std::vector<PauliString> sq = calculate_pauli_strings_max_weight(n_qubits(), std::min(n_qubits(), 2 * max_w));
        size_t const K = n_operators(), S = n_pauli_strings(), n = n_qubits();
        std::vector<uint8_t> codes(S * n), sq_codes(sq.size() * n);
        for (size_t s = 0; s < S; ++s)
            for (size_t q = 0; q < n; ++q)
                codes[s * n + q] = pauli_strings[s].paulis[q].code;
        for (size_t c = 0; c < sq.size(); ++c)
            for (size_t q = 0; q < n; ++q)
                sq_codes[c * n + q] = sq[c].paulis[q].code;
        std::vector<std::complex<T>> out(sq.size() * K);
        gpu::check(fp_sop_square(gpu::context(), gpu::dtype_of<T>(), static_cast<int>(n), S, codes.data(), K,
                                 coeffs.data_handle(), sq.size(), sq_codes.data(), out.data()));
        return SummedPauliOp<T>(sq, out);
    }
    std::vector<PauliOp<T>> split() const
    {
        std::vector<PauliOp<T>> ops;
        ops.reserve(n_operators());
        for (size_t k = 0; k < n_operators(); ++k)
        {
            std::vector<std::complex<T>> c(n_pauli_strings());
            for (size_t i = 0; i < c.size(); ++i)
                c[i] = coeffs(i, k);
            ops.emplace_back(std::move(c), pauli_strings);
        }
        return ops;
    }
    void to_tensor(Tensor<3> A_k_out) const
    {
        for (size_t i = 0; i < pauli_strings.size(); ++i)
        {
            auto [cols, vals] = get_sparse_repr<T>(pauli_strings[i].paulis);
            for (size_t k = 0; k < n_operators(); ++k)
                for (size_t j = 0; j < dim(); ++j)
                    A_k_out(k, j, cols[j]) += coeffs(i, k) * vals[j];
        }
    }

  private:
    // the plan is reused only while a byte-exact snapshot of (codes, coeffs) still compares equal (memcmp, no hash)
    struct PlanCache
    {
        fp_sop *plan = nullptr;
        std::vector<unsigned char> snap_codes, snap_coeffs;
        std::mutex mu;
        bool matches(std::vector<uint8_t> const &codes, void const *c, size_t bytes) const
        {
            return plan && snap_codes.size() == codes.size() && snap_coeffs.size() == bytes &&
                   (codes.empty() || std::memcmp(snap_codes.data(), codes.data(), codes.size()) == 0) &&
                   (bytes == 0 || std::memcmp(snap_coeffs.data(), c, bytes) == 0);
        }
        void remember(std::vector<uint8_t> const &codes, void const *c, size_t bytes)
        {
            snap_codes.assign(codes.begin(), codes.end());
            auto const *p = static_cast<unsigned char const *>(c);
            snap_coeffs.assign(p, p + bytes);
        }
        void reset()
        {
            if (plan)
                fp_sop_destroy(plan);
            plan = nullptr;
            snap_codes.clear();
            snap_coeffs.clear();
        }
    };
    mutable PlanCache cache_;

    void check_ctor(size_t coeff_rows) const
    {
        for (auto const &ps : pauli_strings)
            if (ps.n_qubits() != pauli_strings.front().n_qubits())
                throw std::invalid_argument("All PauliStrings must have the same size");
        if (coeff_rows != pauli_strings.size())
            throw std::invalid_argument("The number of PauliStrings must match the number of rows in the coeffs matrix");
    }
    void adopt(Tensor<2> const coeffs_in)
    {
        if (pauli_strings.empty())
            throw std::invalid_argument("SummedPauliOp needs at least one PauliString");
        check_ctor(coeffs_in.extent(0));
        coeffs_raw.assign(coeffs_in.data_handle(), coeffs_in.data_handle() + coeffs_in.size());
        coeffs = Tensor<2>(coeffs_raw.data(), coeffs_in.extent(0), coeffs_in.extent(1));
    }
    fp_sop *device_plan() const
    {
        size_t const n = n_qubits(), S = n_pauli_strings();
        std::vector<uint8_t> codes(S * n);
        for (size_t s = 0; s < S; ++s)
            for (size_t q = 0; q < n; ++q)
                codes[s * n + q] = pauli_strings[s].paulis[q].code;
        size_t const bytes = coeffs.size() * sizeof(std::complex<T>);
        std::lock_guard<std::mutex> lk(cache_.mu);
        if (!cache_.matches(codes, coeffs.data_handle(), bytes))
        {
            cache_.reset();
            gpu::check(fp_sop_create(gpu::context(), gpu::dtype_of<T>(), static_cast<int>(n), S, codes.data(),
                                     n_operators(), coeffs.data_handle(), &cache_.plan));
            cache_.remember(codes, coeffs.data_handle(), bytes);
        }
        return cache_.plan;
    }
};

} // namespace fast_pauli
