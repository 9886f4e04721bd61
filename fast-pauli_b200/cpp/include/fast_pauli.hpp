// fast_pauli.hpp -- umbrella header of the B200-native build of the fast-pauli C++ API.
//
// Same include name, namespace and class names as the reference (fast_pauli/cpp/include/fast_pauli.hpp:18-23):
//     #include "fast_pauli.hpp"
//     fast_pauli::PauliOp<double> op(coeffs, strings);  op.apply(std::execution::par, new_states, states);
// The nine hot-path methods run on the GPU through libfastpauli_b200.so (include/fastpauli_b200.h); everything
// else (operator algebra, dense debug tensors, generators) is small host code kept so user code compiles.
// Build: g++ -std=c++20 -I<repo>/fast-pauli_b200/cpp/include -I<repo>/include -I$CUDA_HOME/include app.cpp
//        -L<repo>/fast-pauli_b200/lib -lfastpauli_b200
#pragma once
#include "fast_pauli_b200/detail.hpp"
#include "fast_pauli_b200/factory.hpp"
#include "fast_pauli_b200/helpers.hpp"
#include "fast_pauli_b200/pauli.hpp"
#include "fast_pauli_b200/pauli_op.hpp"
#include "fast_pauli_b200/pauli_string.hpp"
#include "fast_pauli_b200/sharded.hpp"
#include "fast_pauli_b200/summed_pauli_op.hpp"
#if __has_include(<fmt/format.h>)
#include "fast_pauli_b200/fmt_support.hpp" // fmt::formatter<Pauli / PauliString / std::complex<double>>, as the reference
#endif
