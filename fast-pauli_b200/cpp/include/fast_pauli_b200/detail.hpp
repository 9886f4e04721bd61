// Shared plumbing of the C++ API layer: C ABI, error translation, execution-policy concept, device context.
#pragma once
#include <complex>
#include <concepts>
#include <cstdint>
#include <execution>
#include <experimental/mdspan>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "fastpauli_b200.h"

namespace fast_pauli
{

// Same concept / trait names as the reference's __type_traits.hpp:26-53.  Both policies route to the GPU: the tag
// only selected the OpenMP code path in the reference, and there is no CPU path here.
template <typename T> struct is_complex : std::false_type
{
};
template <std::floating_point T> struct is_complex<std::complex<T>> : std::true_type
{
};
template <typename T>
concept execution_policy = std::is_execution_policy_v<std::remove_cvref_t<T>>;
template <typename T>
inline constexpr bool is_parallel_execution_policy_v =
    std::is_same_v<std::execution::parallel_policy, std::remove_cvref_t<T>>;

namespace gpu
{
template <std::floating_point T> constexpr int dtype_of()
{
    static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "float or double only");
    return std::is_same_v<T, double> ? FP_C128 : FP_C64;
}

// FP_INVALID_ARGUMENT is raised exactly where the reference throws std::invalid_argument; everything else is a
// runtime failure (CUDA error, no device, out of memory) that the reference could not have.
inline void check(int rc)
{
    if (rc == FP_OK)
        return;
    std::string msg = fp_last_error();
    if (rc == FP_INVALID_ARGUMENT)
        throw std::invalid_argument(msg);
    throw std::runtime_error("fastpauli_b200: " + msg);
}

inline fp_ctx *context()
{
    fp_ctx *ctx = nullptr;
    check(fp_default_ctx(&ctx));
    return ctx;
}

inline uint64_t fnv1a(void const *data, size_t bytes, uint64_t h = 1469598103934665603ull)
{
    auto const *p = static_cast<unsigned char const *>(data);
    for (size_t i = 0; i < bytes; ++i)
    {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}
} // namespace gpu
} // namespace fast_pauli
