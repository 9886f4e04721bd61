// CPU test of the host-side planners (fast-pauli_b200/csrc/pack.hpp, coset_plan.hpp): no GPU, no CUDA headers.
//
// For random operators of every family the kernels care about, the packed / planned form is EVALUATED ON THE HOST
// exactly the way the coset kernels index it --
//     row(c, l)  = deposit(c, non-pivot bits) ^ XOR_{k in l} b_k
//     out[row(c, l)] += sum_groups sum_strings sc * (-1)^{par(base & sz) ^ par(l & szl)} * psi[row(c, l ^ gxl)]
// -- and compared with the definition  out[i] = sum_s h_s (-i)^{nY_s} (-1)^{popc(i & z_s)} psi[i ^ x_s]
// (reference: get_sparse_repr, __pauli_string.hpp:49-118; PauliOp::apply, __pauli_op.hpp:399-468).  That pins the
// GF(2) basis construction, the local x / z coordinates, the pass partition (every group in exactly one pass), the
// chunking limits and the duplicate merge of the packer for tile ranks 1..12 and reserved low bits 0..2.  Further down:
// the paired-mask basis of K3j and the single-state view of K3i, evaluated with the kernels' indexing.
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
    // ---------------------------------------------------------------- K3j: paired-mask basis (coset_plan.hpp: pair_basis)
    // The kernel hard-codes the local x of the eight masks in the re-chosen basis (m_{2p} = e_p, m_{2p+1} = e_p ^
    // e_{8-RB+p}, singles e_{RB+q}) and forms the row factors from the full row index; evaluated on the host that way,
    // for both lane mappings, the pass must reproduce the definition.
    int pair_cases = 0;
    for (int n = 9; n <= 12; ++n)
        for (int per = 1; per <= 3; ++per)
            for (int rep = 0; rep < 6; ++rep)
            {
                std::uniform_int_distribution<int> bit(0, 1);
                std::vector<uint8_t> codes;
                size_t S = 0;
                for (int g = 0; g < 8; ++g)
                {
                    std::vector<int> x(n);
                    for (auto &v : x)
                        v = bit(rng);
                    for (int k = 0; k < per; ++k, ++S)
                        for (int q = 0; q < n; ++q)
                        {
                            int z = bit(rng);
                            codes.push_back(static_cast<uint8_t>(x[q] ? (z ? 2 : 1) : (z ? 3 : 0)));
                        }
                }
                std::vector<cd> h(S);
                for (auto &c : h)
                    c = cd(u(rng), u(rng));
                uint64_t const dim = 1ull << n;
                std::vector<cd> psi(dim), want(dim, cd(0));
                for (auto &a : psi)
                    a = cd(u(rng), u(rng));
                for (size_t s = 0; s < S; ++s)
                {
                    fpk::StringMasks m = fpk::make_masks(n, codes.data() + s * n);
                    cd const c = fpk::times_phase(h[s], m.ny);
                    for (uint64_t i = 0; i < dim; ++i)
                        want[i] += (par(i & m.z) ? -c : c) * psi[i ^ m.x];
                }
                fpk::PackedOp<double> op = fpk::pack_op<double>(n, S, codes.data(), h.data(), true);
                auto passes = fpk::plan_coset<double>(op, n, 8, 0);
                uint64_t nb3[8], nb2[8];
                if (passes.size() != 1 || !fpk::pair_basis<double>(passes[0], nb3, nb2))
                    continue; // dependent or repeated masks: not a K3j pass
                auto const &p = passes[0];
                for (int RB = 2; RB <= 3; ++RB)
                {
                    uint64_t const *nb = RB == 3 ? nb3 : nb2;
                    uint32_t xl[8];
                    for (int q = 0; q < RB; ++q)
                    {
                        xl[2 * q] = 1u << q;
                        xl[2 * q + 1] = (1u << q) | (1u << (8 - RB + q));
                    }
                    for (int q = 0; q < 8 - 2 * RB; ++q)
                        xl[2 * RB + q] = 1u << (RB + q);
                    auto comb = [&](uint32_t l) {
                        uint64_t r = 0;
                        for (int k = 0; k < 8; ++k)
                            if ((l >> k) & 1u)
                                r ^= nb[k];
                        return r;
                    };
                    std::vector<cd> got(dim, cd(0));
                    std::vector<char> hit(dim, 0);
                    for (uint64_t c = 0; c < (dim >> 8); ++c)
                    {
                        uint64_t const base = deposit(c, p.nonpivot_mask);
                        for (uint32_t l = 0; l < 256; ++l)
                        {
                            uint64_t const row = base ^ comb(l);
                            hit[row]++;
                            for (int g = 0; g < 8; ++g)
                            {
                                cd d(0);
                                for (uint32_t s = p.gstart[g]; s < p.gstart[g + 1]; ++s)
                                    d += par(row & p.sz[s]) ? -p.sc[s] : p.sc[s];
                                got[row] += d * psi[base ^ comb(l ^ xl[g])];
                            }
                        }
                    }
                    double err = 0, scale = 1e-300;
                    bool once = true;
                    for (uint64_t i = 0; i < dim; ++i)
                    {
                        once = once && hit[i] == 1;
                        err = std::max(err, std::abs(got[i] - want[i]));
                        scale = std::max(scale, std::abs(want[i]));
                    }
                    EXPECT(once, "K3j basis RB=%d: the tiles do not cover every row exactly once", RB);
                    EXPECT(err / scale < 1e-12, "K3j basis RB=%d n=%d per=%d rel err %.3e", RB, n, per, err / scale);
                }
                ++pair_cases;
            }
    EXPECT(pair_cases >= 20, "only %d paired-mask passes were checked", pair_cases);

    // ---------------------------------------------------------------- single states (coset_plan.hpp: single_state_reshape)
    // One state viewed as 2^(n-4) rows x 16 columns; every pass evaluated the way K3i indexes it: gather row
    // base ^ comb(l ^ xl), column j ^ xlo, sign par(base & z') ^ par(l & zl) ^ par(j & zlo).
    int single_cases = 0;
    for (int n = 6; n <= 13; ++n)
        for (int rep = 0; rep < 4; ++rep)
        {
            size_t S = 0;
            std::vector<uint8_t> codes = make_codes(rng, n, rep == 3 ? 3 : 0, S);
            std::vector<cd> h(S);
            for (auto &c : h)
                c = cd(u(rng), u(rng));
            uint64_t const dim = 1ull << n;
            std::vector<cd> psi(dim), want(dim, cd(0));
            for (auto &a : psi)
                a = cd(u(rng), u(rng));
            for (size_t s = 0; s < S; ++s)
            {
                fpk::StringMasks m = fpk::make_masks(n, codes.data() + s * n);
                cd const c = fpk::times_phase(h[s], m.ny);
                for (uint64_t i = 0; i < dim; ++i)
                    want[i] += (par(i & m.z) ? -c : c) * psi[i ^ m.x];
            }
            fpk::PackedOp<double> op = fpk::pack_op<double>(n, S, codes.data(), h.data(), true);
            fpk::SingleStateOp<double> ss = fpk::single_state_reshape<double>(op, n);
            EXPECT(ss.r.sz.size() == op.sz.size() && ss.xlo.size() == op.sz.size(), "single state: string count");
            int const nr = n - 4, rank = std::min(8, nr);
            auto passes = fpk::plan_coset<double>(ss.r, nr, rank, 0);
            std::vector<cd> got(dim, cd(0));
            for (auto const &p : passes)
                for (uint64_t c = 0; c < ((1ull << nr) >> rank); ++c)
                {
                    uint64_t const base = deposit(c, p.nonpivot_mask);
                    for (uint32_t l = 0; l < (1u << rank); ++l)
                    {
                        auto comb = [&](uint32_t ll) {
                            uint64_t r = base;
                            for (int k = 0; k < rank; ++k)
                                if ((ll >> k) & 1u)
                                    r ^= p.basis.b[k];
                            return r;
                        };
                        uint64_t const row = comb(l);
                        for (size_t g = 0; g < p.gxl.size(); ++g)
                        {
                            uint64_t const src = comb(l ^ p.gxl[g]);
                            for (uint32_t t = p.gstart[g]; t < p.gstart[g + 1]; ++t)
                            {
                                uint8_t const xlo = ss.xlo[p.sidx[t]], zlo = ss.zlo[p.sidx[t]];
                                for (uint32_t j = 0; j < 16; ++j)
                                {
                                    int const odd = par(base & p.sz[t]) ^ par(l & p.szl[t]) ^ par(j & zlo);
                                    got[row * 16 + j] += (odd ? -p.sc[t] : p.sc[t]) * psi[src * 16 + (j ^ xlo)];
                                }
                            }
                        }
                    }
                }
            double err = 0, scale = 1e-300;
            for (uint64_t i = 0; i < dim; ++i)
            {
                err = std::max(err, std::abs(got[i] - want[i]));
                scale = std::max(scale, std::abs(want[i]));
            }
            EXPECT(err / scale < 1e-12, "single state n=%d rep=%d passes=%zu rel err %.3e", n, rep, passes.size(), err / scale);
            ++single_cases;
        }
    std::printf("[host-plan] paired-mask passes: %d | single-state operators: %d\n", pair_cases, single_cases);
    std::printf("[host-plan] operators: %d | assertions: %d | failed: %d\n", cases, checks, failures);
    return failures ? 1 : 0;
}
