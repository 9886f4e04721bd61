// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// extern "C" wrapper around the UNMODIFIED fast-pauli reference headers, which
// are compiled where they lie (-I/root/reference/fast_pauli/cpp/include, see
// oracle/Makefile).  No reference source is copied into this repository.  The
// resulting oracle/_ref/libfastpauli_ref.so is the "real reference" leg of the
// oracle: it validates the C restatement in oracle/pauli_oracle.c, generates /
// checks golden vectors, and is the CPU baseline bench.py times
// (cpu_baseline.kind == "reference").  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Every entry point maps 1:1 onto one reference method (file:line cited on
// each) and takes the same raw arguments as the matching orc_* function in
// oracle/pauli_oracle.c and fp_* one-shot function in include/fastpauli_b200.h:
//   codes : n_strings x n_qubits uint8 (0:I 1:X 2:Y 3:Z), codes[s*n + 0] is the
//           LEFT-most character of the string (most significant qubit, PS:52-54)
//   arrays: row-major (dim, n_states), interleaved (re, im)
//   par   : 0 -> std::execution::seq overload, 1 -> std::execution::par overload
// Return 0 on success, 1 on std::invalid_argument (message via ref_last_error),
// 2 on any other exception.
#include <complex>
#include <cstdint>
#include <cstring>
#include <exception>
#include <execution>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <vector>

#include "fast_pauli.hpp"

#include <omp.h>

namespace
{
thread_local std::string g_err;

template <class T> using C = std::complex<T>;
template <class T> using M1 = std::mdspan<C<T>, std::dextents<size_t, 1>>;
template <class T> using M2 = std::mdspan<C<T>, std::dextents<size_t, 2>>;

fast_pauli::PauliString make_string(int n, uint8_t const *codes)
{
    std::vector<fast_pauli::Pauli> p;
    p.reserve(n);
    for (int q = 0; q < n; ++q)
        p.emplace_back(static_cast<int>(codes[q]));
    return fast_pauli::PauliString(std::move(p));
}

std::vector<fast_pauli::PauliString> make_strings(int n, size_t S, uint8_t const *codes)
{
    std::vector<fast_pauli::PauliString> v;
    v.reserve(S);
    for (size_t s = 0; s < S; ++s)
        v.push_back(make_string(n, codes + s * n));
    return v;
}

template <class F> int guarded(F &&f)
{
    try
    {
        f();
        return 0;
    }
    catch (std::invalid_argument const &e)
    {
        g_err = e.what();
        return 1;
    }
    catch (std::exception const &e)
    {
        g_err = e.what();
        return 2;
    }
}

// PauliString::apply, 1-D (PS:296-341)
template <class T>
int string_apply1d(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, int par)
{
    return guarded([&] {
        auto ps = make_string(n, codes);
        M1<T> o(reinterpret_cast<C<T> *>(out), dim);
        M1<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim);
        C<T> cc(c[0], c[1]);
        if (par)
            ps.apply(std::execution::par, o, i, cc);
        else
            ps.apply(std::execution::seq, o, i, cc);
    });
}

// PauliString::apply_batch (PS:377-436)
template <class T>
int string_apply(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, size_t B, int par)
{
    return guarded([&] {
        auto ps = make_string(n, codes);
        M2<T> o(reinterpret_cast<C<T> *>(out), dim, B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        C<T> cc(c[0], c[1]);
        if (par)
            ps.apply_batch(std::execution::par, o, i, cc);
        else
            ps.apply_batch(std::execution::seq, o, i, cc);
    });
}

// PauliString::expectation_value (PS:470-538)
template <class T>
int string_expval(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, size_t B, int par)
{
    return guarded([&] {
        auto ps = make_string(n, codes);
        M1<T> o(reinterpret_cast<C<T> *>(out), B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        C<T> cc(c[0], c[1]);
        if (par)
            ps.expectation_value(std::execution::par, o, i, cc);
        else
            ps.expectation_value(std::execution::seq, o, i, cc);
    });
}

template <class T> fast_pauli::PauliOp<T> make_op(int n, size_t S, uint8_t const *codes, T const *coeffs)
{
    std::vector<C<T>> h(S);
    for (size_t s = 0; s < S; ++s)
        h[s] = C<T>(coeffs[2 * s], coeffs[2 * s + 1]);
    return fast_pauli::PauliOp<T>(std::move(h), make_strings(n, S, codes));
}

// PauliOp::apply, 1-D (PO:362-383)
template <class T>
int op_apply1d(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim, int par)
{
    return guarded([&] {
        auto op = make_op<T>(n, S, codes, coeffs);
        M1<T> o(reinterpret_cast<C<T> *>(out), dim);
        M1<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim);
        if (par)
            op.apply(std::execution::par, o, i);
        else
            op.apply(std::execution::seq, o, i);
    });
}

// PauliOp::apply, 2-D (PO:399-468)
template <class T>
int op_apply(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim, size_t B,
             int par)
{
    return guarded([&] {
        auto op = make_op<T>(n, S, codes, coeffs);
        M2<T> o(reinterpret_cast<C<T> *>(out), dim, B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        if (par)
            op.apply(std::execution::par, o, i);
        else
            op.apply(std::execution::seq, o, i);
    });
}

// PauliOp::expectation_value (PO:482-549)
template <class T>
int op_expval(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim, size_t B,
              int par)
{
    return guarded([&] {
        auto op = make_op<T>(n, S, codes, coeffs);
        M1<T> o(reinterpret_cast<C<T> *>(out), B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        if (par)
            op.expectation_value(std::execution::par, o, i);
        else
            op.expectation_value(std::execution::seq, o, i);
    });
}

template <class T>
fast_pauli::SummedPauliOp<T> make_sop(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs)
{
    std::vector<C<T>> h(S * K);
    for (size_t s = 0; s < S * K; ++s)
        h[s] = C<T>(coeffs[2 * s], coeffs[2 * s + 1]);
    // (strings, flat coeffs) ctor, SPO:83-92; coeffs are (n_strings, n_operators) row-major
    return fast_pauli::SummedPauliOp<T>(make_strings(n, S, codes), h);
}

// SummedPauliOp::apply (SPO:277-349)
template <class T>
int sop_apply(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out, T const *in, size_t dim,
              size_t B, int par)
{
    return guarded([&] {
        auto op = make_sop<T>(n, S, codes, K, coeffs);
        M2<T> o(reinterpret_cast<C<T> *>(out), dim, B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        if (par)
            op.apply(std::execution::par, o, i);
        else
            op.apply(std::execution::seq, o, i);
    });
}

// SummedPauliOp::apply_weighted (SPO:364-503); data is (n_operators, n_states) real
template <class T, class D>
int sop_apply_weighted(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out, T const *in,
                       D const *data, size_t dim, size_t B, int par)
{
    if constexpr (!std::is_same_v<T, D>)
    {
        // The reference template does not instantiate for data_dtype != T:
        // `coeffs(j, k) * data(k, t)` (SPO:429,484) is std::complex<T> * D, for
        // which <complex> has no operator.  Report "not available in reference".
        g_err = "reference apply_weighted does not compile for data_dtype != T (SPO:484)";
        return 3;
    }
    else
    {
        return guarded([&] {
            auto op = make_sop<T>(n, S, codes, K, coeffs);
            M2<T> o(reinterpret_cast<C<T> *>(out), dim, B);
            M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
            std::mdspan<D, std::dextents<size_t, 2>> d(const_cast<D *>(data), K, B);
            if (par)
                op.apply_weighted(std::execution::par, o, i, d);
            else
                op.apply_weighted(std::execution::seq, o, i, d);
        });
    }
}

// SummedPauliOp::expectation_value (SPO:520-614); out is (n_operators, n_states)
template <class T>
int sop_expval(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out, T const *in, size_t dim,
               size_t B, int par)
{
    return guarded([&] {
        auto op = make_sop<T>(n, S, codes, K, coeffs);
        M2<T> o(reinterpret_cast<C<T> *>(out), K, B);
        M2<T> i(reinterpret_cast<C<T> *>(const_cast<T *>(in)), dim, B);
        if (par)
            op.expectation_value(std::execution::par, o, i);
        else
            op.expectation_value(std::execution::seq, o, i);
    });
}
// SummedPauliOp::square() of the reference (SPO:197-268): output strings as codes + coefficients (n_sq, K) row-major
template <typename T>
int sop_square(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, size_t cap, size_t *n_sq,
               uint8_t *sq_codes, T *coeffs_sq)
{
    return guarded([&] {
        auto op = make_sop<T>(n, S, codes, K, coeffs);
        auto sq = op.square();
        *n_sq = sq.n_pauli_strings();
        if (sq.n_pauli_strings() > cap)
            throw std::invalid_argument("square: output capacity too small");
        for (size_t c = 0; c < sq.n_pauli_strings(); ++c)
        {
            for (int q = 0; q < n; ++q)
                sq_codes[c * static_cast<size_t>(n) + q] = sq.pauli_strings[c].paulis[q].code;
            for (size_t k = 0; k < K; ++k)
            {
                coeffs_sq[2 * (c * K + k)] = sq.coeffs(c, k).real();
                coeffs_sq[2 * (c * K + k) + 1] = sq.coeffs(c, k).imag();
            }
        }
    });
}
} // namespace

extern "C"
{
    char const *ref_last_error(void)
    {
        return g_err.c_str();
    }
    int ref_max_threads(void)
    {
        return omp_get_max_threads();
    }
    void ref_set_threads(int n)
    {
        omp_set_num_threads(n);
    }

#define FP_REF_STAMP(SFX, T)                                                                                           \
    int ref_string_apply1d_##SFX(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, int par)    \
    {                                                                                                                  \
        return string_apply1d<T>(n, codes, c, out, in, dim, par);                                                      \
    }                                                                                                                  \
    int ref_string_apply_##SFX(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, size_t B,     \
                               int par)                                                                                \
    {                                                                                                                  \
        return string_apply<T>(n, codes, c, out, in, dim, B, par);                                                     \
    }                                                                                                                  \
    int ref_string_expval_##SFX(int n, uint8_t const *codes, T const *c, T *out, T const *in, size_t dim, size_t B,    \
                                int par)                                                                               \
    {                                                                                                                  \
        return string_expval<T>(n, codes, c, out, in, dim, B, par);                                                    \
    }                                                                                                                  \
    int ref_op_apply1d_##SFX(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim,  \
                             int par)                                                                                  \
    {                                                                                                                  \
        return op_apply1d<T>(n, S, codes, coeffs, out, in, dim, par);                                                  \
    }                                                                                                                  \
    int ref_op_apply_##SFX(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim,    \
                           size_t B, int par)                                                                          \
    {                                                                                                                  \
        return op_apply<T>(n, S, codes, coeffs, out, in, dim, B, par);                                                 \
    }                                                                                                                  \
    int ref_op_expval_##SFX(int n, size_t S, uint8_t const *codes, T const *coeffs, T *out, T const *in, size_t dim,   \
                            size_t B, int par)                                                                         \
    {                                                                                                                  \
        return op_expval<T>(n, S, codes, coeffs, out, in, dim, B, par);                                                \
    }                                                                                                                  \
    int ref_sop_apply_##SFX(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out, T const *in,     \
                            size_t dim, size_t B, int par)                                                             \
    {                                                                                                                  \
        return sop_apply<T>(n, S, codes, K, coeffs, out, in, dim, B, par);                                             \
    }                                                                                                                  \
    int ref_sop_apply_weighted_##SFX(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out,         \
                                     T const *in, void const *data, int data_is_f64, size_t dim, size_t B, int par)    \
    {                                                                                                                  \
        if (data_is_f64)                                                                                               \
            return sop_apply_weighted<T, double>(n, S, codes, K, coeffs, out, in, static_cast<double const *>(data),   \
                                                 dim, B, par);                                                         \
        return sop_apply_weighted<T, float>(n, S, codes, K, coeffs, out, in, static_cast<float const *>(data), dim, B, \
                                            par);                                                                      \
    }                                                                                                                  \
    int ref_sop_expval_##SFX(int n, size_t S, uint8_t const *codes, size_t K, T const *coeffs, T *out, T const *in,    \
                             size_t dim, size_t B, int par)                                                            \
    {                                                                                                                  \
        return sop_expval<T>(n, S, codes, K, coeffs, out, in, dim, B, par);                                            \
    }

    FP_REF_STAMP(c128, double)
    FP_REF_STAMP(c64, float)

    int ref_sop_square_c128(int n, size_t S, uint8_t const *codes, size_t K, double const *coeffs, size_t cap,
                            size_t *n_sq, uint8_t *sq_codes, double *coeffs_sq)
    {
        return sop_square<double>(n, S, codes, K, coeffs, cap, n_sq, sq_codes, coeffs_sq);
    }
    int ref_sop_square_c64(int n, size_t S, uint8_t const *codes, size_t K, float const *coeffs, size_t cap,
                           size_t *n_sq, uint8_t *sq_codes, float *coeffs_sq)
    {
        return sop_square<float>(n, S, codes, K, coeffs, cap, n_sq, sq_codes, coeffs_sq);
    }
}
