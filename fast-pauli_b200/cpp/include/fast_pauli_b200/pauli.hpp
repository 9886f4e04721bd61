// fast_pauli::Pauli -- one 2x2 Pauli matrix as a small value type (reference API: __pauli.hpp:39-212).
#pragma once
#include <array>
#include <chrono>
#include <compare>
#include <ostream>
#include <random>
#include <ranges>
#include <span>
#include <unordered_map>
#include <utility>

#include "detail.hpp"

// Source compatibility: the reference's headers make these names visible at global scope (__pauli.hpp:29,
// __pauli_op.hpp:26) and pull in <ranges>, <random>, <chrono>, <unordered_map>; code written against them (its own
// tests use `1i` literals and std::views without including anything else) must keep compiling.
using namespace std::literals;
using namespace std::experimental;

namespace fast_pauli
{

struct Pauli
{
    uint8_t code; // 0: I, 1: X, 2: Y, 3: Z

    constexpr Pauli() : code(0)
    {
    }
    template <class T>
        requires std::convertible_to<T, uint8_t> && (!std::same_as<T, char>)
    constexpr Pauli(T const c) : code(static_cast<uint8_t>(c))
    {
        if (c < 0 || c > 3)
            throw std::invalid_argument("Pauli code must be 0, 1, 2, or 3");
    }
    constexpr Pauli(char const symbol) : code(0)
    {
        constexpr char letters[4] = {'I', 'X', 'Y', 'Z'};
        for (uint8_t k = 0; k < 4; ++k)
            if (symbol == letters[k])
            {
                code = k;
                return;
            }
        throw std::invalid_argument("Invalid Pauli matrix symbol");
    }
    Pauli(Pauli const &) = default;
    Pauli &operator=(Pauli const &) noexcept = default;
    friend auto operator<=>(Pauli const &, Pauli const &) = default;

    constexpr char symbol() const
    {
        return "IXYZ"[code & 3];
    }

    // Product of two Pauli matrices as (phase, matrix).  With the (x, z) bit encoding I=(0,0) X=(1,0) Y=(1,1)
    // Z=(0,1) the result matrix is the bitwise XOR and the phase is i^k, k from the symplectic form -- no table.
    friend std::pair<std::complex<double>, Pauli> operator*(Pauli const &lhs, Pauli const &rhs)
    {
        if (lhs.code > 3 || rhs.code > 3)
            throw std::runtime_error("Unexpected Pauli code");
        auto xz = [](uint8_t c) { return std::pair<int, int>{c == 1 || c == 2, c == 2 || c == 3}; };
        auto [x1, z1] = xz(lhs.code);
        auto [x2, z2] = xz(rhs.code);
        int const x = x1 ^ x2, z = z1 ^ z2;
        uint8_t const out = static_cast<uint8_t>(x ? (z ? 2 : 1) : (z ? 3 : 0));
        // sigma(x,z) = i^{x z} X^x Z^z ; X^x1 Z^z1 X^x2 Z^z2 = (-1)^{z1 x2} X^{x1+x2} Z^{z1+z2}
        int k = x1 * z1 + x2 * z2 + 2 * (z1 * x2) - x * z;
        k = ((k % 4) + 4) % 4;
        constexpr std::complex<double> phases[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        return {phases[k], Pauli{static_cast<int>(out)}};
    }

    // dense 2x2 matrix (debug helper)
    template <std::floating_point T> void to_tensor(std::mdspan<std::complex<T>, std::dextents<size_t, 2>> output) const
    {
        if (output.extent(0) != 2 || output.extent(1) != 2)
            throw std::invalid_argument("Pauli::to_tensor needs a 2x2 output");
        using C = std::complex<T>;
        C const m[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, C(0, -1), C(0, 1), 0}, {1, 0, 0, -1}};
        for (size_t a = 0; a < 2; ++a)
            for (size_t b = 0; b < 2; ++b)
                output(a, b) = m[code & 3][2 * a + b];
    }

    friend std::ostream &operator<<(std::ostream &os, Pauli const &p)
    {
        return os << p.symbol();
    }
};

} // namespace fast_pauli
