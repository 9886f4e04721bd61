// Host-side planning of the coset-blocked ("state tile in shared memory") kernels.
//
// Idea.  For a set of x-masks spanning an r-dimensional GF(2) subspace V, the rows {i ^ v : v in V} form a coset
// that every gather i -> i ^ x_g (x_g in V) maps onto itself.  A CTA therefore loads the 2^r rows of ONE coset
// (times a narrow tile of the batch axis) into shared memory once, evaluates ALL groups of the pass out of shared
// memory and writes the 2^r output rows once: HBM traffic is the compulsory read + write no matter how many
// x-masks the pass contains.  Operators whose x-masks span more than r_max dimensions are split into several passes
// (each pass: one read of the batch + one read-modify-write of the output).
//
// With the basis b_0..b_{r-1} in reduced echelon form (pivot bit p_k set in b_k only):
//   row(c, l)   = deposit(c, non-pivot bits) ^ XOR_{k in l} b_k          (bijection (coset c, local l) <-> row)
//   row ^ x_g   = row(c, l ^ xl_g),   xl_g = bits p_k of x_g               (local gather)
//   par(row & z) = par(base & z) ^ par(l & zl),  zl_k = par(b_k & z)        (local sign)
// Reference semantics being accelerated: PauliOp::apply (PO:399-468), SummedPauliOp::apply_weighted (SPO:364-503).
#pragma once
#include <algorithm>
#include <complex>
#include <cstdint>
#include <vector>

#include "pack.hpp"

namespace fpk
{

constexpr int kCosetMaxRank = 12;
constexpr uint32_t kCosetChunkStrings = 128; // strings staged in shared memory at a time (small: more CTAs per SM)
constexpr uint32_t kCosetChunkGroups = 128;

struct CosetChunk
{
    uint32_t g_lo, g_hi; // groups of this chunk (indices into the pass' group arrays)
    uint32_t s_lo, s_hi; // strings of this chunk
};

struct Gf2Basis
{
    int r = 0;
    uint64_t b[kCosetMaxRank] = {};
    int pivot[kCosetMaxRank] = {};

    uint64_t reduce(uint64_t x) const
    {
        for (int k = 0; k < r; ++k)
            if ((x >> pivot[k]) & 1ull)
                x ^= b[k];
        return x;
    }
    bool contains(uint64_t x) const
    {
        return reduce(x) == 0;
    }
    // returns false (and leaves the basis untouched) when x is independent but the basis is full
    bool insert(uint64_t x, int max_rank)
    {
        x = reduce(x);
        if (x == 0)
            return true;
        if (r >= max_rank)
            return false;
        int p = 63 - __builtin_clzll(x);
        for (int k = 0; k < r; ++k)
            if ((b[k] >> p) & 1ull)
                b[k] ^= x;
        b[r] = x;
        pivot[r] = p;
        ++r;
        return true;
    }
    uint64_t pivot_mask() const
    {
        uint64_t m = 0;
        for (int k = 0; k < r; ++k)
            m |= 1ull << pivot[k];
        return m;
    }
    uint32_t coords(uint64_t x) const // x must be in the span
    {
        uint32_t c = 0;
        for (int k = 0; k < r; ++k)
            c |= static_cast<uint32_t>((x >> pivot[k]) & 1ull) << k;
        return c;
    }
    // order the basis by pivot position: local row index bit k <-> k-th lowest pivot, so consecutive local rows
    // (consecutive lanes of a warp) differ in the lowest address bits and neighbouring rows are neighbours in memory
    void sort_by_pivot()
    {
        for (int i = 1; i < r; ++i)
            for (int j = i; j > 0 && pivot[j - 1] > pivot[j]; --j)
            {
                std::swap(pivot[j - 1], pivot[j]);
                std::swap(b[j - 1], b[j]);
            }
    }
    uint32_t zlocal(uint64_t z) const
    {
        uint32_t c = 0;
        for (int k = 0; k < r; ++k)
            c |= static_cast<uint32_t>(__builtin_popcountll(b[k] & z) & 1) << k;
        return c;
    }
};

template <typename T> struct CosetPassHost
{
    Gf2Basis basis;
    uint64_t nonpivot_mask = 0;
    std::vector<uint32_t> gxl;       // [Gp]   local x of each (sub)group
    std::vector<uint32_t> gstart;    // [Gp+1]
    std::vector<uint32_t> szl;       // [Sp]   local z
    std::vector<uint64_t> sz;        // [Sp]   full z (coset-base parity)
    std::vector<std::complex<T>> sc; // [Sp]
    std::vector<uint32_t> sidx;      // [Sp]   index of the string in the packed operator (rows of W)
    std::vector<CosetChunk> chunks;
};

// Greedy pass construction.  Each pass first tries two candidate group sets and keeps the larger:
//  (a) bit-subset cover: grow a set T of <= rank bit positions, always adding the position that completes the most
//      still-uncovered masks (good for low-weight strings: every mask inside T lies in span{e_t});
//  (b) incremental rank: scan the groups in order and take every mask that keeps the GF(2) rank <= rank
//      (good for few, dense masks: any `rank` masks fit one pass).
// reserve_low_bits: number of lowest row-index bits forced into every pass' basis, so that a tile always contains
// runs of 2^reserve consecutive rows (narrow batches: a row is only 16-32 bytes and coalescing must come from the
// row index, SURVEY.md hard part 3).
template <typename T>
inline std::vector<CosetPassHost<T>> plan_coset(PackedOp<T> const &op, int n_qubits, int rank, int reserve_low_bits = 0)
{
    reserve_low_bits = std::max(0, std::min(reserve_low_bits, std::min(rank - 1, n_qubits)));
    int const free_rank = rank - reserve_low_bits;
    uint64_t const low_mask = (1ull << reserve_low_bits) - 1;
    std::vector<CosetPassHost<T>> passes;
    size_t const G = op.gx.size();
    std::vector<char> done(G, 0);
    size_t remaining = G;
    uint64_t const full = n_qubits >= 64 ? ~0ull : ((1ull << n_qubits) - 1);
    while (remaining)
    {
        // ---- candidate (a): bit-subset cover
        uint64_t Tmask = low_mask;
        if (G <= 4096) // the cover search is O(rank * n * G) per pass: only worth it for moderately sized operators
        {
            int used = reserve_low_bits;
            while (used < rank)
            {
                int best_bit = -1;
                long best_gain = -1, best_partial = -1;
                for (int bit = 0; bit < n_qubits; ++bit)
                {
                    if ((Tmask >> bit) & 1ull)
                        continue;
                    uint64_t T2 = Tmask | (1ull << bit);
                    long gain = 0, partial = 0;
                    for (size_t g = 0; g < G; ++g)
                    {
                        if (done[g] || !((op.gx[g] >> bit) & 1ull))
                            continue;
                        if ((op.gx[g] & ~T2) == 0)
                            ++gain;
                        else if (__builtin_popcountll(op.gx[g] & ~T2) == 1)
                            ++partial;
                    }
                    if (gain > best_gain || (gain == best_gain && partial > best_partial))
                    {
                        best_gain = gain;
                        best_partial = partial;
                        best_bit = bit;
                    }
                }
                if (best_bit < 0 || (best_gain <= 0 && best_partial <= 0))
                    break;
                Tmask |= 1ull << best_bit;
                ++used;
            }
        }
        size_t count_a = 0;
        for (size_t g = 0; g < G; ++g)
            if (!done[g] && (op.gx[g] & ~Tmask) == 0)
                ++count_a;
        // ---- candidate (b): incremental rank
        Gf2Basis bb;
        for (int bit = 0; bit < reserve_low_bits; ++bit)
            bb.insert(1ull << bit, rank);
        size_t count_b = 0;
        for (size_t g = 0; g < G; ++g)
            if (!done[g] && bb.insert(op.gx[g], rank))
                ++count_b;
        (void)free_rank;

        CosetPassHost<T> pass;
        if (count_a >= count_b)
        {
            for (int bit = 0; bit < n_qubits; ++bit)
                if ((Tmask >> bit) & 1ull)
                    pass.basis.insert(1ull << bit, rank);
        }
        else
        {
            for (int bit = 0; bit < reserve_low_bits; ++bit)
                pass.basis.insert(1ull << bit, rank);
            for (size_t g = 0; g < G; ++g)
                if (!done[g])
                    pass.basis.insert(op.gx[g], rank); // same scan as above: rebuilds bb
        }
        // pad to the full tile rank with free low bit positions (the tile then simply holds several cosets)
        for (int bit = 0; bit < n_qubits && pass.basis.r < rank; ++bit)
            pass.basis.insert(1ull << bit, rank);
        pass.basis.sort_by_pivot();
        pass.nonpivot_mask = full & ~pass.basis.pivot_mask();

        // ---- collect every not-yet-done group inside the span.  The diagonal group (x = 0) lies in every span: it goes
        // to the first pass that holds fewer than 8 other groups, or to the last pass, so that passes of exactly eight
        // independent x-masks (chain Hamiltonians: 8 bond masks + the ZZ terms) stay eligible for K3j
        size_t in_span = 0, others_left = 0;
        for (size_t g = 0; g < G; ++g)
            if (!done[g] && op.gx[g] != 0)
            {
                ++others_left;
                if (pass.basis.contains(op.gx[g]))
                    ++in_span;
            }
        bool const take_diagonal = in_span < 8 || in_span == others_left;
        pass.gstart.push_back(0);
        for (size_t g = 0; g < G; ++g)
        {
            if (done[g] || !pass.basis.contains(op.gx[g]) || (op.gx[g] == 0 && !take_diagonal))
                continue;
            done[g] = 1;
            --remaining;
            uint32_t xl = pass.basis.coords(op.gx[g]);
            uint32_t s0 = op.gstart[g], s1 = op.gstart[g + 1];
            for (uint32_t s = s0; s < s1; s += kCosetChunkStrings) // oversize groups become several sub-groups
            {
                uint32_t e = std::min<uint32_t>(s1, s + kCosetChunkStrings);
                for (uint32_t t = s; t < e; ++t)
                {
                    pass.szl.push_back(pass.basis.zlocal(op.sz[t]));
                    pass.sz.push_back(op.sz[t]);
                    pass.sc.push_back(op.sc[t]);
                    pass.sidx.push_back(t);
                }
                pass.gxl.push_back(xl);
                pass.gstart.push_back(static_cast<uint32_t>(pass.szl.size()));
            }
        }
        // ---- chunk the (sub)groups so that the metadata of one chunk fits the shared-memory staging area
        uint32_t g_lo = 0;
        uint32_t const Gp = static_cast<uint32_t>(pass.gxl.size());
        while (g_lo < Gp)
        {
            uint32_t g_hi = g_lo;
            while (g_hi < Gp && g_hi - g_lo < kCosetChunkGroups &&
                   pass.gstart[g_hi + 1] - pass.gstart[g_lo] <= kCosetChunkStrings)
                ++g_hi;
            pass.chunks.push_back(CosetChunk{g_lo, g_hi, pass.gstart[g_lo], pass.gstart[g_hi]});
            g_lo = g_hi;
        }
        passes.push_back(std::move(pass));
    }
    return passes;
}

// ---------------------------------------------------------------- K3j (coset4.cuh): paired-mask basis of a pass
// A pass of exactly eight INDEPENDENT x-masks m_0..m_7 (plan order): the masks themselves span the pass, so the local
// coordinates can be re-chosen such that masks (0,1), (2,3) [and (4,5)] differ in one of the row bits a lane owns:
//   nb2 (two pairs):   b_0..5 = m_0, m_2, m_4, m_5, m_6, m_7    b_6 = m_0^m_1   b_7 = m_2^m_3
//   nb3 (three pairs): b_0..4 = m_0, m_2, m_4, m_6, m_7         b_5 = m_0^m_1   b_6 = m_2^m_3   b_7 = m_4^m_5
// i.e. in the new coordinates  m_{2p} = e_p,  m_{2p+1} = e_p ^ e_{8-RB+p}  (p < RB),  m_{2RB+q} = e_{RB+q}.
// Returns false when the pass is not of that kind.
template <typename T> inline bool pair_basis(CosetPassHost<T> const &h, uint64_t (&nb3)[8], uint64_t (&nb2)[8])
{
    if (h.basis.r != 8 || h.gxl.size() != 8)
        return false;
    uint64_t m[8];
    for (int g = 0; g < 8; ++g)
    {
        m[g] = 0;
        for (int k = 0; k < 8; ++k)
            if ((h.gxl[g] >> k) & 1u)
                m[g] ^= h.basis.b[k];
    }
    uint32_t red[8], rk = 0; // GF(2) rank of the local coordinates
    for (int g = 0; g < 8; ++g)
    {
        uint32_t v = h.gxl[g];
        for (uint32_t j = 0; j < rk; ++j)
            v = std::min(v, v ^ red[j]);
        if (v)
            red[rk++] = v;
    }
    if (rk != 8)
        return false;
    uint64_t const a3[8] = {m[0], m[2], m[4], m[6], m[7], m[0] ^ m[1], m[2] ^ m[3], m[4] ^ m[5]};
    uint64_t const a2[8] = {m[0], m[2], m[4], m[5], m[6], m[7], m[0] ^ m[1], m[2] ^ m[3]};
    for (int k = 0; k < 8; ++k)
    {
        nb3[k] = a3[k];
        nb2[k] = a2[k];
    }
    return true;
}

// ---------------------------------------------------------------- single states (K3i, coset3.cuh)
// One state of 2^n amplitudes viewed as 2^(n-4) rows x 16 columns (the column is the 4 lowest index bits): the operator
// on the upper n-4 bits, (x >> 4, z >> 4), with every string's low nibbles kept beside it (column permutation j -> j ^ xlo,
// column sign (-1)^popc(j & zlo)).  Strings are sorted by (x >> 4, z >> 4) and grouped by x >> 4; no merging (two
// strings that differ only in their low nibbles stay two strings).
template <typename T> struct SingleStateOp
{
    PackedOp<T> r;            // gx, gstart, sz, sc on n - 4 qubits
    std::vector<uint8_t> xlo; // [S]
    std::vector<uint8_t> zlo; // [S]
};

template <typename T> inline SingleStateOp<T> single_state_reshape(PackedOp<T> const &op, int n_qubits)
{
    struct Rs
    {
        uint64_t x, z;
        uint8_t xlo, zlo;
        std::complex<T> c;
    };
    std::vector<Rs> rs;
    for (size_t g = 0; g + 1 < op.gstart.size(); ++g)
        for (uint32_t t = op.gstart[g]; t < op.gstart[g + 1]; ++t)
            rs.push_back(Rs{op.gx[g] >> 4, op.sz[t] >> 4, static_cast<uint8_t>(op.gx[g] & 15u),
                            static_cast<uint8_t>(op.sz[t] & 15u), op.sc[t]});
    std::stable_sort(rs.begin(), rs.end(), [](Rs const &a, Rs const &b) { return a.x != b.x ? a.x < b.x : a.z < b.z; });
    SingleStateOp<T> o;
    o.r.n_qubits = n_qubits - 4;
    o.r.n_strings_in = rs.size();
    for (size_t i = 0; i < rs.size(); ++i)
    {
        if (i == 0 || rs[i].x != rs[i - 1].x)
        {
            o.r.gx.push_back(rs[i].x);
            o.r.gstart.push_back(static_cast<uint32_t>(i));
        }
        o.r.sz.push_back(rs[i].z);
        o.r.sc.push_back(rs[i].c);
        o.xlo.push_back(rs[i].xlo);
        o.zlo.push_back(rs[i].zlo);
    }
    o.r.gstart.push_back(static_cast<uint32_t>(rs.size()));
    return o;
}

} // namespace fpk
