// CPU test of the host-side planners (fast-pauli_b200/csrc/pack.hpp, coset_plan.hpp): no GPU, no CUDA headers.
//
// For random operators of every family the kernels care about, the packed / planned form is EVALUATED ON THE HOST
// exactly the way the coset kernels index it --
//     row(c, l)  = deposit(c, non-pivot bits) ^ XOR_{k in l} b_k
//     out[row(c, l)] += sum_groups sum_strings sc * (-1)^{par(base & sz) ^ par(l & szl)} * psi[row(c, l ^ gxl)]
// -- and compared with the definition  out[i] = sum_s h_s (-i)^{nY_s} (-1)^{popc(i & z_s)} psi[i ^ x_s]
// (reference: get_sparse_repr, __pauli_string.hpp:49-118; PauliOp::apply, __pauli_op.hpp:399-468).  That pins the
// GF(2) basis construction, the local x / z coordinates, the pass partition (every group in exactly one pass), the
// chunking limits and the duplicate merge of the packer for tile ranks 1..12 and reserved low bits 0..2.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "coset_plan.hpp"

using cd = std::complex<double>;
static int failures = 0, checks = 0;
#define EXPECT(cond, ...)                                                                                              \
    do                                                                                                                 \
    {                                                                                                                  \
        ++checks;                                                                                                      \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            ++failures;                                                                                                \
            std::printf("FAIL %s:%d: %s | ", __FILE__, __LINE__, #cond);                                               \
            std::printf(__VA_ARGS__);                                                                                  \
            std::printf("\n");                                                                                         \
        }                                                                                                              \
    } while (0)

static uint64_t deposit(uint64_t src, uint64_t mask)
{
    uint64_t res = 0;
    for (uint64_t bb = 1; mask; bb <<= 1)
    {
        uint64_t low = mask & (~mask + 1);
        if (src & bb)
            res |= low;
        mask &= mask - 1;
    }
    return res;
}
static int par(uint64_t v)
{
    return __builtin_popcountll(v) & 1;
}

static std::vector<uint8_t> make_codes(std::mt19937_64 &rng, int n, int family, size_t &S)
{
    std::vector<uint8_t> codes;
    auto push = [&](std::vector<uint8_t> const &s) { codes.insert(codes.end(), s.begin(), s.end()); };
    std::uniform_int_distribution<int> letter(0, 3), xyz(1, 3), bit(0, 1);
    switch (family)
    {
    case 0: // i.i.d.
        S = 1 + rng() % 60;
        for (size_t s = 0; s < S; ++s)
        {
            std::vector<uint8_t> st(n);
            for (auto &c : st)
                c = static_cast<uint8_t>(letter(rng));
            push(st);
        }
        break;
    case 1: // weight <= 4
        S = 1 + rng() % 150;
        for (size_t s = 0; s < S; ++s)
        {
            std::vector<uint8_t> st(n, 0);
            int w = 1 + static_cast<int>(rng() % 4);
            for (int k = 0; k < w; ++k)
                st[rng() % n] = static_cast<uint8_t>(xyz(rng));
            push(st);
        }
        break;
    case 2: // few x-masks, many z variants
    {
        int G = 1 + static_cast<int>(rng() % 9);
        S = 0;
        for (int g = 0; g < G; ++g)
        {
            std::vector<int> x(n);
            for (auto &v : x)
                v = bit(rng);
            int zv = 1 + static_cast<int>(rng() % 8);
            for (int k = 0; k < zv; ++k, ++S)
            {
                std::vector<uint8_t> st(n);
                for (int q = 0; q < n; ++q)
                {
                    int z = bit(rng);
                    st[q] = static_cast<uint8_t>(x[q] ? (z ? 2 : 1) : (z ? 3 : 0));
                }
                push(st);
            }
        }
        break;
    }
    case 3: // nearest-neighbour chain
        S = 0;
        for (int q = 0; q + 1 < n; ++q)
            for (uint8_t c = 1; c <= 3; ++c, ++S)
            {
                std::vector<uint8_t> st(n, 0);
                st[q] = st[q + 1] = c;
                push(st);
            }
        if (S == 0)
        {
            push(std::vector<uint8_t>(n, 3));
            S = 1;
        }
        break;
    default: // duplicates + diagonal
        S = 2 + rng() % 20;
        {
            std::vector<uint8_t> base(n);
            for (auto &c : base)
                c = static_cast<uint8_t>(letter(rng));
            for (size_t s = 0; s + 1 < S; ++s)
                push(base);
            push(std::vector<uint8_t>(n, 3));
        }
    }
    return codes;
}

int main()
{
    std::mt19937_64 rng(20241017);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    int cases = 0;
    for (int n = 1; n <= 11; ++n)
        for (int family = 0; family < 5; ++family)
            for (int rep = 0; rep < 3; ++rep)
            {
                size_t S = 0;
                std::vector<uint8_t> codes = make_codes(rng, n, family, S);
                std::vector<cd> h(S);
                for (auto &c : h)
                    c = cd(u(rng), u(rng));
                uint64_t const dim = 1ull << n;
                std::vector<cd> psi(dim), want(dim, cd(0));
                for (auto &a : psi)
                    a = cd(u(rng), u(rng));
                // the definition
                for (size_t s = 0; s < S; ++s)
                {
                    fpk::StringMasks m = fpk::make_masks(n, codes.data() + s * n);
                    cd const c = fpk::times_phase(h[s], m.ny);
                    for (uint64_t i = 0; i < dim; ++i)
                        want[i] += (par(i & m.z) ? -c : c) * psi[i ^ m.x];
                }
                fpk::PackedOp<double> op = fpk::pack_op<double>(n, S, codes.data(), h.data(), /*merge=*/true);
                // packer invariants
                EXPECT(op.gstart.size() == op.gx.size() + 1 && op.gstart.back() == op.sz.size(), "group table");
                for (size_t g = 1; g < op.gx.size(); ++g)
                    EXPECT(op.gx[g - 1] < op.gx[g], "x-masks strictly increasing");
                for (size_t g = 0; g < op.gx.size(); ++g)
                    for (uint32_t s = op.gstart[g] + 1; s < op.gstart[g + 1]; ++s)
                        EXPECT(op.sz[s - 1] < op.sz[s], "z-masks strictly increasing inside a group (merged)");
                for (int rank = 1; rank <= std::min(n, fpk::kCosetMaxRank); ++rank)
                    for (int reserve = 0; reserve <= std::min(2, rank - 1); ++reserve)
                    {
                        auto passes = fpk::plan_coset<double>(op, n, rank, reserve);
                        std::vector<cd> got(dim, cd(0));
                        size_t groups_seen = 0;
                        for (auto const &p : passes)
                        {
                            EXPECT(p.basis.r == rank, "tile rank %d != %d", p.basis.r, rank);
                            for (int b = 0; b < reserve; ++b)
                                EXPECT(p.basis.contains(1ull << b), "reserved low bit %d missing", b);
                            for (int k = 0; k < p.basis.r; ++k)
                                for (int j = 0; j < p.basis.r; ++j)
                                    EXPECT(((p.basis.b[j] >> p.basis.pivot[k]) & 1ull) == (j == k ? 1u : 0u),
                                           "basis not in reduced echelon form");
                            for (int k = 1; k < p.basis.r; ++k)
                                EXPECT(p.basis.pivot[k - 1] < p.basis.pivot[k], "pivots not sorted");
                            for (auto const &ch : p.chunks)
                                EXPECT(ch.g_hi > ch.g_lo && ch.g_hi - ch.g_lo <= fpk::kCosetChunkGroups &&
                                           ch.s_hi - ch.s_lo <= fpk::kCosetChunkStrings,
                                       "chunk limits");
                            EXPECT(!p.chunks.empty() && p.chunks.front().g_lo == 0 &&
                                       p.chunks.back().g_hi == p.gxl.size(),
                                   "chunks cover the pass");
                            groups_seen += p.gxl.size();
                            uint64_t const n_cosets = dim >> rank;
                            for (uint64_t c = 0; c < n_cosets; ++c)
                            {
                                uint64_t const base = deposit(c, p.nonpivot_mask);
                                for (uint32_t l = 0; l < (1u << rank); ++l)
                                {
                                    uint64_t row = base;
                                    for (int k = 0; k < rank; ++k)
                                        if ((l >> k) & 1u)
                                            row ^= p.basis.b[k];
                                    cd acc(0);
                                    for (size_t g = 0; g < p.gxl.size(); ++g)
                                    {
                                        uint32_t const lsrc = l ^ p.gxl[g];
                                        uint64_t src = base;
                                        for (int k = 0; k < rank; ++k)
                                            if ((lsrc >> k) & 1u)
                                                src ^= p.basis.b[k];
                                        cd d(0);
                                        for (uint32_t s = p.gstart[g]; s < p.gstart[g + 1]; ++s)
                                            d += (par(base & p.sz[s]) ^ par(l & p.szl[s])) ? -p.sc[s] : p.sc[s];
                                        acc += d * psi[src];
                                    }
                                    got[row] += acc;
                                }
                            }
                        }
                        EXPECT(groups_seen >= op.gx.size(), "every x-group lands in a pass");
                        double err = 0, scale = 1e-300;
                        for (uint64_t i = 0; i < dim; ++i)
                        {
                            err = std::max(err, std::abs(got[i] - want[i]));
                            scale = std::max(scale, std::abs(want[i]));
                        }
                        EXPECT(err / scale < 1e-12, "n=%d family=%d rank=%d reserve=%d passes=%zu rel err %.3e", n, family,
                               rank, reserve, passes.size(), err / scale);
                    }
                ++cases;
            }
    std::printf("[host-plan] operators: %d | assertions: %d | failed: %d\n", cases, checks, failures);
    return failures ? 1 : 0;
}
