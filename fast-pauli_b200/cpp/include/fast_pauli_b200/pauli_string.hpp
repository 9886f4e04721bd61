// fast_pauli::PauliString -- tensor product of Pauli matrices; apply / apply_batch / expectation_value on the GPU.
// Reference API being mirrored: __pauli_string.hpp:126-561 (same member names, overloads and exceptions).
#pragma once
#include <algorithm>
#include <functional>
#include <ostream>
#include <span>
#include <string>
#include <tuple>

#include "pauli.hpp"

namespace fast_pauli
{

namespace gpu
{
// uint8 code row of a string, left-most character first (the C ABI's string encoding)
inline std::vector<uint8_t> codes_of(std::vector<Pauli> const &paulis)
{
    std::vector<uint8_t> c(paulis.size());
    std::transform(paulis.begin(), paulis.end(), c.begin(), [](Pauli const &p) { return p.code; });
    return c;
}
} // namespace gpu

// (k, m) with  P[i, k[i]] = m[i]  -- host-side closed form of the PauliComposer table, used only by the dense
// debug exports.  k[i] = i ^ x,  m[i] = (-i)^nY (-1)^popcount(i & z); bit (n-1-q) belongs to paulis[q].
template <std::floating_point T>
std::tuple<std::vector<size_t>, std::vector<std::complex<T>>> get_sparse_repr(std::vector<Pauli> const &paulis)
{
    size_t const n = paulis.size();
    if (n == 0)
        return {};
    uint64_t x = 0, z = 0;
    unsigned ny = 0;
    for (size_t q = 0; q < n; ++q)
    {
        uint64_t const bit = uint64_t(1) << (n - 1 - q);
        uint8_t const c = paulis[q].code;
        x |= (c == 1 || c == 2) ? bit : 0;
        z |= (c == 2 || c == 3) ? bit : 0;
        ny += c == 2;
    }
    std::complex<T> const phases[4] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}};
    std::complex<T> const base = phases[ny & 3u];
    size_t const dim = size_t(1) << n;
    std::vector<size_t> k(dim);
    std::vector<std::complex<T>> m(dim);
    for (size_t i = 0; i < dim; ++i)
    {
        k[i] = i ^ x;
        m[i] = (__builtin_popcountll(i & z) & 1) ? -base : base;
    }
    return {std::move(k), std::move(m)};
}

struct PauliString
{
    uint8_t weight = 0;
    std::vector<Pauli> paulis;

    PauliString() noexcept = default;
    PauliString(std::vector<Pauli> p) : paulis(std::move(p))
    {
        count_weight();
    }
    PauliString(std::span<fast_pauli::Pauli> p) : paulis(p.begin(), p.end())
    {
        count_weight();
    }
    PauliString(std::string const &str)
    {
        paulis.reserve(str.size());
        for (char ch : str)
        {
            if (ch != 'I' && ch != 'X' && ch != 'Y' && ch != 'Z')
                throw std::invalid_argument(std::string("Invalid Pauli character ") + ch);
            paulis.emplace_back(ch);
        }
        count_weight();
    }
    PauliString(char const *str) : PauliString(std::string(str))
    {
    }
    PauliString(PauliString const &) = default;
    PauliString &operator=(PauliString const &) = default;
    friend auto operator<=>(PauliString const &, PauliString const &) = default;

    // (phase, string) of the matrix product of two strings
    friend std::pair<std::complex<double>, PauliString> operator*(PauliString const &lhs, PauliString const &rhs)
    {
        if (lhs.dim() != rhs.dim())
            throw std::invalid_argument("PauliStrings must have the same size");
        std::complex<double> phase = 1;
        std::vector<Pauli> prod(lhs.n_qubits());
        for (size_t q = 0; q < prod.size(); ++q)
        {
            auto [ph, p] = lhs.paulis[q] * rhs.paulis[q];
            phase *= ph;
            prod[q] = p;
        }
        return {phase, PauliString(std::move(prod))};
    }

    size_t n_qubits() const noexcept
    {
        return paulis.size();
    }
    size_t dim() const noexcept
    {
        return paulis.empty() ? 0 : size_t(1) << paulis.size();
    }
    std::string str() const
    {
        std::string s(paulis.size(), 'I');
        std::transform(paulis.begin(), paulis.end(), s.begin(), [](Pauli const &p) { return p.symbol(); });
        return s;
    }

    // ---- apply, 1-D: new_states[i] += c * m[i] * states[k[i]]                    (reference: PS:296-341)
    template <std::floating_point T>
    void apply(std::mdspan<std::complex<T>, std::dextents<size_t, 1>> new_states,
               std::mdspan<std::complex<T>, std::dextents<size_t, 1>> states, std::complex<T> const c = 1.0) const
    {
        apply(std::execution::seq, new_states, states, c);
    }
    template <std::floating_point T, execution_policy ExecutionPolicy>
    void apply(ExecutionPolicy &&, std::mdspan<std::complex<T>, std::dextents<size_t, 1>> new_states,
               std::mdspan<std::complex<T>, std::dextents<size_t, 1>> states, std::complex<T> const c = 1.0) const
    {
        if (states.size() != dim())
            throw std::invalid_argument("Input vector size must match the number of qubits");
        if (states.size() != new_states.size())
            throw std::invalid_argument("new_states must have the same dimensions as states");
        auto const codes = gpu::codes_of(paulis);
        gpu::check(fp_string_apply(gpu::context(), gpu::dtype_of<T>(), static_cast<int>(n_qubits()), codes.data(), &c,
                                   new_states.data_handle(), states.data_handle(), states.size(), 1, /*accumulate=*/1));
    }

    // ---- apply_batch: new_states_T(i,t) += c * m[i] * states_T(k[i], t), (dim, n_states) row-major (PS:377-436)
    template <std::floating_point T>
    void apply_batch(std::mdspan<std::complex<T>, std::dextents<size_t, 2>> new_states_T,
                     std::mdspan<std::complex<T>, std::dextents<size_t, 2>> const states_T,
                     std::complex<T> const c) const
    {
        apply_batch(std::execution::seq, new_states_T, states_T, c);
    }
    template <std::floating_point T, execution_policy ExecutionPolicy>
    void apply_batch(ExecutionPolicy &&, std::mdspan<std::complex<T>, std::dextents<size_t, 2>> new_states_T,
                     std::mdspan<std::complex<T>, std::dextents<size_t, 2>> const states_T,
                     std::complex<T> const c) const
    {
        if (states_T.extent(0) != dim())
            throw std::invalid_argument("[PauliString] states shape (" + std::to_string(states_T.extent(0)) +
                                        ") must match the dimension of the operators (" + std::to_string(dim()) + ")");
        if (states_T.extent(0) != new_states_T.extent(0) || states_T.extent(1) != new_states_T.extent(1))
            throw std::invalid_argument("[PauliString] new_states must have the same dimensions as states");
        auto const codes = gpu::codes_of(paulis);
        gpu::check(fp_string_apply(gpu::context(), gpu::dtype_of<T>(), static_cast<int>(n_qubits()), codes.data(), &c,
                                   new_states_T.data_handle(), states_T.data_handle(), states_T.extent(0),
                                   states_T.extent(1), /*accumulate=*/1));
    }

    // ---- expectation_value: out[t] += sum_i conj(states(i,t)) c m[i] states(k[i],t)               (PS:470-538)
    template <std::floating_point T>
    void expectation_value(std::mdspan<std::complex<T>, std::dextents<size_t, 1>> expectation_vals_out,
                           std::mdspan<std::complex<T>, std::dextents<size_t, 2>> states,
                           std::complex<T> const c = 1.0) const
    {
        expectation_value(std::execution::seq, expectation_vals_out, states, c);
    }
    template <std::floating_point T, execution_policy ExecutionPolicy>
    void expectation_value(ExecutionPolicy &&,
                           std::mdspan<std::complex<T>, std::dextents<size_t, 1>> expectation_vals_out,
                           std::mdspan<std::complex<T>, std::dextents<size_t, 2>> states,
                           std::complex<T> const c = 1.0) const
    {
        if (states.extent(0) != dim())
            throw std::invalid_argument("[PauliString] states shape (" + std::to_string(states.extent(0)) +
                                        ") must match the dimension of the operators (" + std::to_string(dim()) + ")");
        if (expectation_vals_out.extent(0) != states.extent(1))
            throw std::invalid_argument("[PauliString] expectation_vals_out shape must match the number of states");
        auto const codes = gpu::codes_of(paulis);
        gpu::check(fp_string_expval(gpu::context(), gpu::dtype_of<T>(), static_cast<int>(n_qubits()), codes.data(), &c,
                                    expectation_vals_out.data_handle(), states.data_handle(), states.extent(0),
                                    states.extent(1), /*accumulate=*/1));
    }

    // dense (dim x dim) matrix, host-side debug helper                                              (PS:549-560)
    template <std::floating_point T> void to_tensor(std::mdspan<std::complex<T>, std::dextents<size_t, 2>> output) const
    {
        if (output.extent(0) != dim() || output.extent(1) != dim())
            throw std::invalid_argument("Output tensor must have the same dimensions as the PauliString");
        auto [k, m] = get_sparse_repr<T>(paulis);
        for (size_t i = 0; i < k.size(); ++i)
            output(i, k[i]) = m[i];
    }

    friend std::ostream &operator<<(std::ostream &os, PauliString const &ps)
    {
        return os << ps.str();
    }

  private:
    void count_weight()
    {
        weight = static_cast<uint8_t>(std::count_if(paulis.begin(), paulis.end(), [](Pauli const &p) { return p.code > 0; }));
    }
};

} // namespace fast_pauli

template <> struct std::hash<fast_pauli::PauliString>
{
    std::size_t operator()(fast_pauli::PauliString const &key) const
    {
        // two bits per qubit folded into a 64-bit FNV hash: no string formatting on the hot path of dedupe maps
        uint64_t h = 1469598103934665603ull;
        for (auto const &p : key.paulis)
        {
            h ^= p.code;
            h *= 1099511628211ull;
        }
        return static_cast<std::size_t>(h);
    }
};
