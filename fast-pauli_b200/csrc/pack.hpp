// Host-side packing of Pauli strings into the (x, z, phase) form the kernels consume.
//
// Reference semantics being restated (fast_pauli/cpp/include/__pauli_string.hpp:49-118, get_sparse_repr):
//   k[i] = i ^ x,  m[i] = (-i)^nY (-1)^popcount(i & z),
//   bit (n-1-q) of x is set when codes[q] is X or Y, of z when it is Y or Z (the reference reverses
//   the string before building the table, PS:52-54).
// The reference rebuilds dim-sized (k, m) tables on every call (PS:323,408,499); here a string is three words
// and an operator is a list of them sorted by x-mask, so strings sharing an x-mask share one gather.
#pragma once
#include <algorithm>
#include <complex>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace fpk
{

struct StringMasks
{
    uint64_t x = 0, z = 0;
    uint32_t ny = 0; // number of Y (mod 4)
};

inline StringMasks make_masks(int n, uint8_t const *codes)
{
    StringMasks m;
    for (int q = 0; q < n; ++q)
    {
        uint8_t c = codes[q];
        if (c > 3)
            throw std::invalid_argument("Pauli code must be 0, 1, 2, or 3");
        uint64_t bit = 1ull << (n - 1 - q);
        if (c == 1 || c == 2)
            m.x |= bit;
        if (c == 2 || c == 3)
            m.z |= bit;
        if (c == 2)
            m.ny = (m.ny + 1) & 3u;
    }
    return m;
}

// c * (-i)^ny, exact (swap / negate only)
template <typename T> inline std::complex<T> times_phase(std::complex<T> c, uint32_t ny)
{
    switch (ny & 3u)
    {
    case 0:
        return c;
    case 1:
        return {c.imag(), -c.real()}; // * (-i)
    case 2:
        return {-c.real(), -c.imag()};
    default:
        return {-c.imag(), c.real()}; // * (+i)
    }
}

template <typename T> struct PackedOp
{
    int n_qubits = 0;
    size_t n_strings_in = 0;
    std::vector<uint64_t> gx;        // [G]
    std::vector<uint32_t> gstart;    // [G+1]
    std::vector<uint64_t> sz;        // [S]
    std::vector<std::complex<T>> sc; // [S]  h * (-i)^nY (duplicates merged when `merge`)
    std::vector<uint8_t> sodd;       // [S]  nY & 1
    std::vector<uint32_t> sny;       // [S]  nY & 3
    std::vector<uint32_t> perm;      // [S]  original index of each packed string (first one when merged)
};

// Sort by (x, z) (stable), optionally merge identical strings, build the group table.
template <typename T>
inline PackedOp<T> pack_op(int n, size_t S, uint8_t const *codes, std::complex<T> const *coeffs, bool merge)
{
    if (n < 0 || n > 62)
        throw std::invalid_argument("n_qubits must be in [0, 62]");
    PackedOp<T> p;
    p.n_qubits = n;
    p.n_strings_in = S;
    std::vector<StringMasks> mk(S);
    for (size_t s = 0; s < S; ++s)
        mk[s] = make_masks(n, codes + s * static_cast<size_t>(n));
    std::vector<uint32_t> order(S);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (mk[a].x != mk[b].x)
            return mk[a].x < mk[b].x;
        return mk[a].z < mk[b].z;
    });
    for (size_t r = 0; r < S; ++r)
    {
        uint32_t s = order[r];
        std::complex<T> c = coeffs ? times_phase(coeffs[s], mk[s].ny) : std::complex<T>(0);
        bool same_x = !p.gx.empty() && p.gx.back() == mk[s].x;
        if (merge && same_x && !p.sz.empty() && p.sz.back() == mk[s].z)
        {
            p.sc.back() += c; // identical string: same x, z (hence same nY parity... and same nY) -> sum coefficients
            continue;
        }
        if (!same_x)
        {
            p.gx.push_back(mk[s].x);
            p.gstart.push_back(static_cast<uint32_t>(p.sz.size()));
        }
        p.sz.push_back(mk[s].z);
        p.sc.push_back(c);
        p.sodd.push_back(static_cast<uint8_t>(mk[s].ny & 1u));
        p.sny.push_back(mk[s].ny);
        p.perm.push_back(s);
    }
    p.gstart.push_back(static_cast<uint32_t>(p.sz.size()));
    if (p.gx.empty())
        p.gstart.assign(1, 0u);
    return p;
}

} // namespace fpk
