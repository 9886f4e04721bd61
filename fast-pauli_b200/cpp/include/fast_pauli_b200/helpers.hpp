// String-set generators used by tests and benchmarks (reference API: __pauli_helpers.hpp:36-153).  The enumeration
// ORDER is part of the contract (the reference tests compare against literal lists, test_pauli_helpers.cpp:25-166):
// by letters first (X < Y < Z, left-most letter most significant), then by position combination in
// lexicographic order.
#pragma once
#include <string>
#include <vector>

#include "pauli_string.hpp"

namespace fast_pauli
{

// all 3^weight words over {X, Y, Z}, lexicographic
inline std::vector<std::string> get_nontrivial_paulis(size_t const weight)
{
    if (weight == 0)
        return {};
    size_t count = 1;
    for (size_t i = 0; i < weight; ++i)
        count *= 3;
    std::vector<std::string> words(count, std::string(weight, 'X'));
    for (size_t w = 0; w < count; ++w)
    {
        size_t v = w;
        for (size_t pos = weight; pos-- > 0; v /= 3)
            words[w][pos] = "XYZ"[v % 3];
    }
    return words;
}

// all k-subsets of {0..n-1}, lexicographic
inline std::vector<std::vector<size_t>> idx_combinations(size_t const n, size_t const k)
{
    std::vector<std::vector<size_t>> result;
    if (k > n)
        return result;
    std::vector<size_t> combo(k);
    for (size_t i = 0; i < k; ++i)
        combo[i] = i;
    while (true)
    {
        result.push_back(combo);
        // advance: right-most index that can still move
        size_t i = k;
        while (i > 0 && combo[i - 1] == n - k + (i - 1))
            --i;
        if (i == 0)
            break;
        ++combo[i - 1];
        for (size_t j = i; j < k; ++j)
            combo[j] = combo[j - 1] + 1;
    }
    return result;
}

inline std::vector<PauliString> calculate_pauli_strings(size_t const n_qubits, size_t const weight)
{
    if (weight == 0)
        return {PauliString(std::string(n_qubits, 'I'))};
    auto const words = get_nontrivial_paulis(weight);
    auto const combos = idx_combinations(n_qubits, weight);
    std::vector<PauliString> result;
    result.reserve(words.size() * combos.size());
    for (auto const &word : words)
        for (auto const &combo : combos)
        {
            std::string s(n_qubits, 'I');
            for (size_t k = 0; k < combo.size(); ++k)
                s[combo[k]] = word[k];
            result.emplace_back(s);
        }
    return result;
}

inline std::vector<PauliString> calculate_pauli_strings_max_weight(size_t n_qubits, size_t weight)
{
    std::vector<PauliString> result;
    for (size_t w = 0; w <= weight; ++w)
    {
        auto ps = calculate_pauli_strings(n_qubits, w);
        result.insert(result.end(), ps.begin(), ps.end());
    }
    return result;
}

} // namespace fast_pauli
