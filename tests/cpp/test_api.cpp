// C++ API tests of the B200 build, written the way the reference's doctest files are (dense-matrix oracles built
// in-test by explicit Kronecker products; every hot-path method exercised with both execution policies), but
// against the GPU-backed classes.  Reference counterparts: fast_pauli/cpp/tests/test_pauli_string.cpp,
// test_pauli_op.cpp, test_summed_pauli_op.cpp, test_pauli.cpp, test_pauli_helpers.cpp, test_factory.cpp.
//
//   test_api --host-only   : everything that needs no GPU (value types, algebra, generators, exceptions)
//   test_api               : the above + all nine hot-path methods on the GPU for complex128 and complex64
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "fast_pauli.hpp"

namespace fp = fast_pauli;
using cd = std::complex<double>;

static int g_failures = 0, g_checks = 0;
#define CHECK(cond)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        ++g_checks;                                                                                                    \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            ++g_failures;                                                                                              \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);                                                \
        }                                                                                                              \
    } while (0)
#define CHECK_THROWS(expr)                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        ++g_checks;                                                                                                    \
        bool threw = false;                                                                                            \
        try                                                                                                            \
        {                                                                                                              \
            expr;                                                                                                      \
        }                                                                                                              \
        catch (std::invalid_argument const &)                                                                          \
        {                                                                                                              \
            threw = true;                                                                                              \
        }                                                                                                              \
        if (!threw)                                                                                                    \
        {                                                                                                              \
            ++g_failures;                                                                                              \
            std::printf("FAIL %s:%d  expected std::invalid_argument: %s\n", __FILE__, __LINE__, #expr);                \
        }                                                                                                              \
    } while (0)

// ---- dense oracle: explicit Kronecker product of 2x2 matrices, left-most factor most significant
template <class T> std::vector<std::complex<T>> dense_of(std::string const &s)
{
    using C = std::complex<T>;
    size_t const n = s.size(), dim = size_t(1) << n;
    std::vector<C> m(dim * dim, C(0));
    auto el = [](char p, int a, int b) -> C {
        switch (p)
        {
        case 'I':
            return a == b ? C(1) : C(0);
        case 'X':
            return a != b ? C(1) : C(0);
        case 'Y':
            return a == b ? C(0) : (a == 0 ? C(0, -1) : C(0, 1));
        default:
            return a == b ? (a == 0 ? C(1) : C(-1)) : C(0);
        }
    };
    for (size_t i = 0; i < dim; ++i)
        for (size_t j = 0; j < dim; ++j)
        {
            C v(1);
            for (size_t q = 0; q < n; ++q)
            {
                int a = (i >> (n - 1 - q)) & 1, b = (j >> (n - 1 - q)) & 1;
                v *= el(s[q], a, b);
            }
            m[i * dim + j] = v;
        }
    return m;
}

template <class T> double tol()
{
    return std::is_same_v<T, double> ? 1e-12 : 2e-5;
}

template <class T> double max_abs(std::vector<std::complex<T>> const &v)
{
    double m = 0;
    for (auto const &x : v)
        m = std::max<double>(m, std::abs(x));
    return m;
}

// ------------------------------------------------------------------------------------------------ host-only
static void test_pauli_value_type()
{
    // Cayley table against dense 2x2 products (reference: test_pauli.cpp:60-101)
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
        {
            auto [phase, p] = fp::Pauli(a) * fp::Pauli(b);
            auto A = dense_of<double>(std::string(1, "IXYZ"[a])), B = dense_of<double>(std::string(1, "IXYZ"[b]));
            auto P = dense_of<double>(std::string(1, "IXYZ"[p.code]));
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                {
                    cd prod = A[i * 2] * B[j] + A[i * 2 + 1] * B[2 + j];
                    CHECK(std::abs(prod - phase * P[i * 2 + j]) < 1e-15);
                }
        }
    CHECK(fp::Pauli('Y').code == 2);
    CHECK_THROWS(fp::Pauli(4));
    CHECK_THROWS(fp::Pauli('Q'));
    std::vector<cd> blob(4);
    std::mdspan<cd, std::dextents<size_t, 2>> m(blob.data(), 2, 2);
    fp::Pauli('Y').to_tensor(m);
    CHECK(m(0, 1) == cd(0, -1) && m(1, 0) == cd(0, 1));
}

static void test_string_host()
{
    fp::PauliString ps("IXYZ");
    CHECK(ps.n_qubits() == 4 && ps.dim() == 16 && ps.weight == 3);
    CHECK(fp::PauliString().dim() == 0);
    CHECK(fp::PauliString("IIII").weight == 0);
    CHECK_THROWS(fp::PauliString("IXQZ"));
    CHECK(ps.str() == "IXYZ");
    // get_sparse_repr against the dense matrix for the reference's test strings (test_pauli_string.cpp:255)
    for (std::string s : {"IXYZ", "YYIX", "XXYIYZ", "IZIXYYZ", "X", "Y", "Z", "I"})
    {
        auto [k, mvals] = fp::get_sparse_repr<double>(fp::PauliString(s).paulis);
        auto D = dense_of<double>(s);
        size_t dim = size_t(1) << s.size();
        for (size_t i = 0; i < dim; ++i)
            for (size_t j = 0; j < dim; ++j)
                CHECK(D[i * dim + j] == (j == k[i] ? mvals[i] : cd(0)));
    }
    // string product vs dense product
    auto [phase, prod] = fp::PauliString("XYZ") * fp::PauliString("ZZX");
    auto A = dense_of<double>("XYZ"), B = dense_of<double>("ZZX"), P = dense_of<double>(prod.str());
    for (size_t i = 0; i < 8; ++i)
        for (size_t j = 0; j < 8; ++j)
        {
            cd acc = 0;
            for (size_t l = 0; l < 8; ++l)
                acc += A[i * 8 + l] * B[l * 8 + j];
            CHECK(std::abs(acc - phase * P[i * 8 + j]) < 1e-15);
        }
    CHECK_THROWS(fp::PauliString("XY") * fp::PauliString("XYZ"));
    CHECK(std::hash<fp::PauliString>()(fp::PauliString("XYZ")) == std::hash<fp::PauliString>()(fp::PauliString("XYZ")));
}

static void test_helpers_and_factory()
{
    // literal enumeration order (reference: test_pauli_helpers.cpp:25-166)
    auto w1 = fp::get_nontrivial_paulis(1);
    CHECK((w1 == std::vector<std::string>{"X", "Y", "Z"}));
    auto w2 = fp::get_nontrivial_paulis(2);
    CHECK((w2 == std::vector<std::string>{"XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"}));
    CHECK(fp::get_nontrivial_paulis(0).empty());
    auto c42 = fp::idx_combinations(4, 2);
    CHECK((c42 == std::vector<std::vector<size_t>>{{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}}));
    auto s21 = fp::calculate_pauli_strings(2, 1);
    std::vector<std::string> got;
    for (auto const &p : s21)
        got.push_back(p.str());
    CHECK((got == std::vector<std::string>{"XI", "IX", "YI", "IY", "ZI", "IZ"}));
    auto all = fp::calculate_pauli_strings_max_weight(3, 3);
    CHECK(all.size() == 64);
    CHECK(all.front().str() == "III");
    CHECK(fp::calculate_pauli_strings_max_weight(4, 2).size() == 1 + 12 + 54);
    std::vector<cd> blob;
    auto r = fp::rand<cd, 2>(blob, {3, 5});
    CHECK(r.extent(0) == 3 && r.extent(1) == 5 && blob.size() == 15);
    for (auto const &v : blob)
        CHECK(v.real() >= 0 && v.real() < 1 && v.imag() >= 0 && v.imag() < 1);
    std::vector<cd> blob2;
    fp::rand<cd, 2>(blob2, {3, 5});
    CHECK(blob == blob2); // deterministic: seed 18 (__factory.hpp:95)
    auto z = fp::zeros<cd, 1>(blob2, {7});
    CHECK(z.extent(0) == 7 && blob2[3] == cd(0));
}

static void test_op_host()
{
    CHECK_THROWS((fp::PauliOp<double>({cd(1)}, {fp::PauliString("XX"), fp::PauliString("YY")})));
    CHECK_THROWS((fp::PauliOp<double>({cd(1), cd(1)}, {fp::PauliString("XX"), fp::PauliString("YYY")})));
    fp::PauliOp<double> a({cd(1, 1), cd(0.5)}, {fp::PauliString("XY"), fp::PauliString("ZI")});
    fp::PauliOp<double> b({cd(2), cd(0, -1)}, {fp::PauliString("YY"), fp::PauliString("XY")});
    CHECK(a.dim() == 4 && a.n_qubits() == 2 && a.n_pauli_strings() == 2);
    auto ab = a * b;
    // dense check of op * op (with dedupe) through to_tensor
    auto dense = [](fp::PauliOp<double> const &op) {
        std::vector<cd> blob(op.dim() * op.dim());
        std::mdspan<cd, std::dextents<size_t, 2>> m(blob.data(), op.dim(), op.dim());
        op.to_tensor(m);
        return blob;
    };
    auto A = dense(a), B = dense(b), AB = dense(ab);
    for (size_t i = 0; i < 4; ++i)
        for (size_t j = 0; j < 4; ++j)
        {
            cd acc = 0;
            for (size_t l = 0; l < 4; ++l)
                acc += A[i * 4 + l] * B[l * 4 + j];
            CHECK(std::abs(acc - AB[i * 4 + j]) < 1e-14);
        }
    a.extend(fp::PauliString("XY"), cd(1), true);
    CHECK(a.n_pauli_strings() == 2 && a.coeffs[0] == cd(2, 1));
    a.extend(fp::PauliString("XX"), cd(1), true);
    CHECK(a.n_pauli_strings() == 3);
    a.scale(cd(2));
    CHECK(a.coeffs[0] == cd(4, 2));
    CHECK((-a).coeffs[0] == cd(-4, -2));
    CHECK_THROWS(a.extend(fp::PauliString("XXX"), cd(1)));
    // SummedPauliOp ctor checks (SPO:47-65)
    std::vector<cd> cblob(6, cd(1));
    std::mdspan<cd, std::dextents<size_t, 2>> c32(cblob.data(), 3, 2);
    CHECK_THROWS((fp::SummedPauliOp<double>(std::vector<std::string>{"XX", "YY"}, c32)));
    std::mdspan<cd, std::dextents<size_t, 2>> c23(cblob.data(), 2, 3);
    fp::SummedPauliOp<double> sop(std::vector<std::string>{"XX", "YY"}, c23);
    CHECK(sop.n_operators() == 3 && sop.n_pauli_strings() == 2 && sop.dim() == 4);
    fp::SummedPauliOp<double> copy(sop); // copies own their coefficient storage
    CHECK(copy.coeffs.data_handle() == copy.coeffs_raw.data() && copy.coeffs.data_handle() != sop.coeffs.data_handle());
    CHECK(sop.split().size() == 3);
}

// ------------------------------------------------------------------------------------------------ GPU
template <class T> void test_string_gpu()
{
    using C = std::complex<T>;
    for (std::string s : {"IXYZ", "YYIX", "XXYIYZ", "IZIXYYZ", "IZIXYYZIXYZ", "IIII", "Z"})
    {
        fp::PauliString ps(s);
        size_t const dim = ps.dim(), B = 10;
        auto D = dense_of<T>(s.size() <= 7 ? s : std::string("I"));
        std::vector<C> sb, ob(dim * B), eb(B), v_in(dim), v_out(dim);
        auto states = fp::rand<C, 2>(sb, {dim, B});
        for (auto &x : sb)
            x += C(1, 1); // amplitudes in [1,2)^2 like test_pauli_string.cpp:249-278
        C const c(0.5, -2.0);
        std::mdspan<C, std::dextents<size_t, 2>> out(ob.data(), dim, B);
        ps.apply_batch(std::execution::par, out, states, c);
        std::vector<C> ob2(dim * B);
        std::mdspan<C, std::dextents<size_t, 2>> out2(ob2.data(), dim, B);
        ps.apply_batch(out2, states, c); // seq overload
        CHECK(ob == ob2);
        std::mdspan<C, std::dextents<size_t, 1>> ev(eb.data(), B);
        ps.expectation_value(std::execution::par, ev, states, c);
        for (size_t i = 0; i < dim; ++i)
            v_in[i] = states(i, 0);
        std::mdspan<C, std::dextents<size_t, 1>> vin(v_in.data(), dim), vout(v_out.data(), dim);
        ps.apply(std::execution::par, vout, vin, c);
        if (s.size() <= 7)
        {
            double const scale = 2.9 * std::abs(c);
            for (size_t t = 0; t < B; ++t)
            {
                C e = 0;
                for (size_t i = 0; i < dim; ++i)
                {
                    C acc = 0;
                    for (size_t j = 0; j < dim; ++j)
                        acc += D[i * dim + j] * states(j, t);
                    acc *= c;
                    CHECK(std::abs(out(i, t) - acc) < tol<T>() * scale);
                    if (t == 0)
                        CHECK(std::abs(vout(i) - acc) < tol<T>() * scale);
                    e += std::conj(states(i, t)) * acc;
                }
                CHECK(std::abs(ev(t) - e) < tol<T>() * scale * 8 * dim);
            }
        }
        // accumulate semantics: a second call adds on top (PS:419,432)
        ps.apply_batch(std::execution::par, out, states, c);
        for (size_t i = 0; i < dim * B; i += 7)
            CHECK(std::abs(ob[i] - T(2) * ob2[i]) < tol<T>() * 8);
        // shape errors (PS:271-283, 343-359, 438-450)
        std::vector<C> wrong(2 * dim * B);
        std::mdspan<C, std::dextents<size_t, 2>> wst(wrong.data(), 2 * dim, B);
        CHECK_THROWS(ps.apply_batch(std::execution::par, out, wst, c));
        CHECK_THROWS(ps.expectation_value(std::execution::par, ev, wst, c));
        std::mdspan<C, std::dextents<size_t, 1>> wev(eb.data(), B - 1);
        CHECK_THROWS(ps.expectation_value(std::execution::par, wev, states, c));
    }
    // "IXI" on e6 + e7 (test_pauli_string.cpp:231-247)
    std::vector<C> st(8, C(0)), nw(8, C(0));
    st[6] = st[7] = 1;
    fp::PauliString("IXI").apply(std::mdspan<C, std::dextents<size_t, 1>>(nw.data(), 8),
                                 std::mdspan<C, std::dextents<size_t, 1>>(st.data(), 8));
    CHECK(nw[4] == C(1) && nw[5] == C(1) && nw[6] == C(0) && nw[7] == C(0));
}

template <class T> void check_op(fp::PauliOp<T> const &op, size_t B)
{
    using C = std::complex<T>;
    size_t const dim = op.dim();
    std::vector<C> dense_blob(dim * dim), sb, ob(dim * B), eb(B);
    std::mdspan<C, std::dextents<size_t, 2>> dense(dense_blob.data(), dim, dim);
    op.to_tensor(dense);
    auto states = fp::rand<C, 2>(sb, {dim, B});
    std::mdspan<C, std::dextents<size_t, 2>> out(ob.data(), dim, B);
    op.apply(std::execution::par, out, states);
    std::mdspan<C, std::dextents<size_t, 1>> ev(eb.data(), B);
    op.expectation_value(std::execution::par, ev, states);
    double scale = 0;
    for (auto const &c : op.coeffs)
        scale += std::abs(c);
    scale = std::max(1.0, scale) * 2;
    for (size_t t = 0; t < B; ++t)
    {
        C e = 0;
        for (size_t i = 0; i < dim; ++i)
        {
            C acc = 0;
            for (size_t j = 0; j < dim; ++j)
                acc += dense(i, j) * states(j, t);
            CHECK(std::abs(out(i, t) - acc) < tol<T>() * scale);
            e += std::conj(states(i, t)) * acc;
        }
        CHECK(std::abs(ev(t) - e) < tol<T>() * scale * dim);
    }
}

template <class T> void test_op_gpu()
{
    using C = std::complex<T>;
    // 1/2/10 states x 1/2/6 strings (test_pauli_op.cpp:183-226)
    check_op<T>(fp::PauliOp<T>({C(1)}, {fp::PauliString("IXYZ")}), 1);
    check_op<T>(fp::PauliOp<T>({C(0.5, 1), C(-2, 0.25)}, {fp::PauliString("IXYZ"), fp::PauliString("YYIX")}), 2);
    check_op<T>(fp::PauliOp<T>({C(1), C(0, 1), C(-1), C(0.3), C(2, 2), C(0, -0.5)},
                               {fp::PauliString("XXYIYZ"), fp::PauliString("ZZZIII"), fp::PauliString("IIIIII"),
                                fp::PauliString("XXYIYZ"), fp::PauliString("YIXZIZ"), fp::PauliString("ZIZIZI")}),
                10);
    // all weight <= 2 strings on 6 qubits with random coefficients (test_pauli_op.cpp:269-327)
    {
        auto strings = fp::calculate_pauli_strings_max_weight(6, 2);
        std::vector<C> cb;
        fp::rand<C, 1>(cb, {strings.size()});
        check_op<T>(fp::PauliOp<T>(cb, strings), 10);
    }
    // 16 identical IIII strings with coefficient 1/16 act as the identity (test_pauli_op.cpp:228-267)
    {
        fp::PauliOp<T> op(std::vector<C>(16, C(1.0 / 16)), std::vector<fp::PauliString>(16, fp::PauliString("IIII")));
        std::vector<C> sb, ob(16 * 3);
        auto states = fp::rand<C, 2>(sb, {16, 3});
        std::mdspan<C, std::dextents<size_t, 2>> out(ob.data(), 16, 3);
        op.apply(out, states);
        for (size_t i = 0; i < 48; ++i)
            CHECK(std::abs(ob[i] - sb[i]) < tol<T>());
    }
    // 1-D apply equals PauliString::apply (test_pauli_op.cpp:159-162) and the plan cache follows member edits
    {
        std::vector<C> st(16), o1(16), o2(16);
        for (size_t i = 0; i < 16; ++i)
            st[i] = C(0.25 * i, 0.5 * (i % 5));
        fp::PauliOp<T> op({C(1)}, {fp::PauliString("IXYZ")});
        std::mdspan<C, std::dextents<size_t, 1>> s1(st.data(), 16), a(o1.data(), 16), b(o2.data(), 16);
        op.apply(std::execution::par, a, s1);
        fp::PauliString("IXYZ").apply(std::execution::par, b, s1);
        CHECK(o1 == o2);
        op.coeffs[0] = C(3); // public member edit must invalidate the cached plan
        std::fill(o1.begin(), o1.end(), C(0));
        op.apply(a, s1);
        for (size_t i = 0; i < 16; ++i)
            CHECK(std::abs(o1[i] - T(3) * o2[i]) < tol<T>() * 8);
    }
    // error paths (test_pauli_op.cpp:101-117)
    {
        fp::PauliOp<T> op({C(1), C(1)}, {fp::PauliString("XYZ"), fp::PauliString("III")});
        std::vector<C> a(4 * 2), b(4 * 2), e(2);
        std::mdspan<C, std::dextents<size_t, 2>> ma(a.data(), 4, 2), mb(b.data(), 4, 2);
        CHECK_THROWS(op.apply(std::execution::par, ma, mb));
        CHECK_THROWS(op.expectation_value(std::execution::par, std::mdspan<C, std::dextents<size_t, 1>>(e.data(), 2), mb));
    }
}

template <class T> void test_summed_gpu()
{
    using C = std::complex<T>;
    // all weight <= 2 strings on 6 qubits x 100 operators x 10 states; apply vs the sum of per-operator
    // PauliOp::apply, apply_weighted vs the manual triple loop, expectation_value vs per-string expectation
    // values contracted by hand (test_summed_pauli_op.cpp:30-409)
    size_t const n = 6, K = 20, B = 10;
    auto strings = fp::calculate_pauli_strings_max_weight(n, 2);
    size_t const S = strings.size(), dim = size_t(1) << n;
    std::vector<C> cb, sb;
    auto coeffs = fp::rand<C, 2>(cb, {S, K});
    auto states = fp::rand<C, 2>(sb, {dim, B});
    fp::SummedPauliOp<T> sop(strings, coeffs);
    std::vector<T> db(K * B);
    for (size_t i = 0; i < db.size(); ++i)
        db[i] = T(0.1) + T(i % 7) / 7;
    std::mdspan<T, std::dextents<size_t, 2>> data(db.data(), K, B);

    std::vector<C> ob(dim * B), wb(dim * B), eb(K * B), ref_apply(dim * B), ref_w(dim * B), ref_e(K * B);
    std::mdspan<C, std::dextents<size_t, 2>> out(ob.data(), dim, B), wout(wb.data(), dim, B), ev(eb.data(), K, B);
    sop.apply(std::execution::par, out, states);
    sop.apply_weighted(std::execution::par, wout, states, data);
    sop.expectation_value(std::execution::par, ev, states);

    auto ops = sop.split();
    for (size_t k = 0; k < K; ++k)
    {
        std::vector<C> tmp(dim * B), etmp(B);
        std::mdspan<C, std::dextents<size_t, 2>> t(tmp.data(), dim, B);
        ops[k].apply(std::execution::par, t, states);
        std::mdspan<C, std::dextents<size_t, 1>> e(etmp.data(), B);
        ops[k].expectation_value(std::execution::par, e, states);
        for (size_t i = 0; i < dim; ++i)
            for (size_t b = 0; b < B; ++b)
            {
                ref_apply[i * B + b] += tmp[i * B + b];
                ref_w[i * B + b] += tmp[i * B + b] * data(k, b);
            }
        for (size_t b = 0; b < B; ++b)
            ref_e[k * B + b] = etmp[b];
    }
    double const sa = max_abs(ref_apply), sw = max_abs(ref_w), se = max_abs(ref_e);
    for (size_t i = 0; i < dim * B; ++i)
    {
        CHECK(std::abs(ob[i] - ref_apply[i]) < tol<T>() * sa);
        CHECK(std::abs(wb[i] - ref_w[i]) < tol<T>() * sw);
    }
    for (size_t i = 0; i < K * B; ++i)
        CHECK(std::abs(eb[i] - ref_e[i]) < tol<T>() * se);
    // shape errors (SPO:384-399, 539-558; test_summed_pauli_op.cpp:203-215)
    std::vector<T> bad(3 * B);
    CHECK_THROWS(sop.apply_weighted(std::execution::par, wout, states, std::mdspan<T, std::dextents<size_t, 2>>(bad.data(), 3, B)));
    std::vector<C> bad_e((K + 1) * B);
    CHECK_THROWS(sop.expectation_value(std::execution::par, std::mdspan<C, std::dextents<size_t, 2>>(bad_e.data(), K + 1, B), states));
    // square() against the dense definition on a small case (test_summed_pauli_op.cpp:411-470)
    {
        auto s2 = fp::calculate_pauli_strings_max_weight(3, 1);
        std::vector<C> c2;
        auto cc = fp::rand<C, 2>(c2, {s2.size(), 3});
        fp::SummedPauliOp<T> small(s2, cc);
        auto sq = small.square();
        size_t const d = 8;
        std::vector<C> A(3 * d * d), A2(3 * d * d);
        small.to_tensor(std::mdspan<C, std::dextents<size_t, 3>>(A.data(), 3, d, d));
        sq.to_tensor(std::mdspan<C, std::dextents<size_t, 3>>(A2.data(), 3, d, d));
        for (size_t k = 0; k < 3; ++k)
            for (size_t i = 0; i < d; ++i)
                for (size_t j = 0; j < d; ++j)
                {
                    C acc = 0;
                    for (size_t l = 0; l < d; ++l)
                        acc += A[(k * d + i) * d + l] * A[(k * d + l) * d + j];
                    CHECK(std::abs(acc - A2[(k * d + i) * d + j]) < (std::is_same_v<T, double> ? 1e-12 : 1e-4));
                }
    }
}

int main(int argc, char **argv)
{
    bool host_only = argc > 1 && std::strcmp(argv[1], "--host-only") == 0;
    test_pauli_value_type();
    test_string_host();
    test_helpers_and_factory();
    test_op_host();
    if (!host_only)
    {
        int n = 0;
        if (fp_device_count(&n) != FP_OK || n == 0)
        {
            std::printf("no CUDA device: the GPU tests cannot run (there is no CPU fallback)\n");
            return 2;
        }
        test_string_gpu<double>();
        test_string_gpu<float>();
        test_op_gpu<double>();
        test_op_gpu<float>();
        test_summed_gpu<double>();
        test_summed_gpu<float>();
    }
    std::printf("%s: %d checks, %d failures%s\n", g_failures ? "FAILED" : "OK", g_checks, g_failures,
                host_only ? " (host-only)" : "");
    return g_failures ? 1 : 0;
}
