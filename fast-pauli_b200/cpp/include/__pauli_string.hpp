// Forwarding header: user code written against the reference includes "__pauli_string.hpp" by name
// (e.g. its examples/01, 03, 05); here the declarations live in fast_pauli_b200/pauli_string.hpp.
#pragma once
#include "fast_pauli_b200/pauli_string.hpp"
