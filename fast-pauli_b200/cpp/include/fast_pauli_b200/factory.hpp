// torch-like tensor factories over a caller-owned std::vector blob (reference API: __factory.hpp:40-128).
// rand() keeps the reference's input distribution -- mt19937(seed = 18), U[0,1) for real and imaginary parts --
// because it defines the reference tests' data.
#pragma once
#include <array>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

#include "detail.hpp"

namespace fast_pauli
{
namespace factory_detail
{
template <size_t n_dim> size_t volume(std::array<size_t, n_dim> const &extents)
{
    return std::accumulate(extents.begin(), extents.end(), size_t(1), std::multiplies<>());
}
// mdspan over `data` with the given extents (index pack instead of passing std::array: works with both the C++23
// <mdspan> and CCCL's cuda::std::mdspan, whose extents constructor wants its own array type)
template <typename T, size_t n_dim, size_t... I>
auto view_impl(T *data, std::array<size_t, n_dim> const &e, std::index_sequence<I...>)
{
    return std::mdspan<T, std::dextents<size_t, n_dim>>(data, e[I]...);
}
template <typename T, size_t n_dim> auto view(T *data, std::array<size_t, n_dim> const &e)
{
    return view_impl<T, n_dim>(data, e, std::make_index_sequence<n_dim>{});
}
} // namespace factory_detail

template <typename T, size_t n_dim>
    requires is_complex<T>::value || std::floating_point<T>
auto empty(std::vector<T> &blob, std::array<size_t, n_dim> extents)
{
    blob.resize(factory_detail::volume(extents));
    return factory_detail::view<T, n_dim>(blob.data(), extents);
}

template <typename T, typename... Is>
    requires(is_complex<T>::value || std::floating_point<T>) && (std::integral<Is> && ...)
auto empty(std::vector<T> &blob, Is... dims)
{
    return empty<T, sizeof...(Is)>(blob, std::array<size_t, sizeof...(Is)>{static_cast<size_t>(dims)...});
}

template <typename T, size_t n_dim>
    requires is_complex<T>::value || std::floating_point<T>
auto zeros(std::vector<T> &blob, std::array<size_t, n_dim> extents)
{
    blob.assign(factory_detail::volume(extents), T(0));
    return factory_detail::view<T, n_dim>(blob.data(), extents);
}

template <typename T, size_t n_dim>
    requires is_complex<T>::value || std::floating_point<T>
auto rand(std::vector<T> &blob, std::array<size_t, n_dim> extents, size_t seed = 18)
{
    blob.assign(factory_detail::volume(extents), T(0));
    std::mt19937 gen(seed);
    if constexpr (is_complex<T>::value)
    {
        std::uniform_real_distribution<typename T::value_type> dis(0, 1.0);
        for (auto &v : blob)
        {
            auto re = dis(gen);
            auto im = dis(gen);
            v = T{re, im};
        }
    }
    else
    {
        std::uniform_real_distribution<T> dis(0, 1.0);
        for (auto &v : blob)
            v = dis(gen);
    }
    return factory_detail::view<T, n_dim>(blob.data(), extents);
}

} // namespace fast_pauli
