// Optional {fmt} integration.  The reference's headers pull in <fmt/format.h> and teach fmt to print Pauli,
// PauliString and std::complex<double> (__pauli.hpp:216-262, __pauli_string.hpp:566-581); user code written against
// them (the reference's examples 03 and 05, for instance) relies on both.  This build does not depend on fmt: the
// header is included by fast_pauli.hpp only when <fmt/format.h> is on the include path.
#pragma once
#include <fmt/format.h>
#include <fmt/ranges.h>

#include <complex>
#include <string>

#include "pauli.hpp"
#include "pauli_string.hpp"

template <> struct fmt::formatter<fast_pauli::Pauli> : fmt::formatter<char>
{
    auto format(fast_pauli::Pauli const &p, fmt::format_context &ctx) const
    {
        return fmt::formatter<char>::format(p.symbol(), ctx);
    }
};

template <> struct fmt::formatter<fast_pauli::PauliString> : fmt::formatter<std::string>
{
    auto format(fast_pauli::PauliString const &ps, fmt::format_context &ctx) const
    {
        return fmt::formatter<std::string>::format(ps.str(), ctx);
    }
};

// "(re, imi)" like the reference; skipped when <fmt/std.h> (which has its own complex formatter) was included first
#ifndef FMT_STD_H_
template <> struct fmt::formatter<std::complex<double>>
{
    constexpr auto parse(fmt::format_parse_context &ctx)
    {
        return ctx.begin();
    }
    auto format(std::complex<double> const &v, fmt::format_context &ctx) const
    {
        return fmt::format_to(ctx.out(), "({}, {}i)", v.real(), v.imag());
    }
};
#endif
