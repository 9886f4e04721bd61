"""``helpers`` submodule of the reference Python package (fast_pauli/cpp/src/fast_pauli.cpp:44-104): string-set
generators and the sparse representation.  The enumeration ORDER is part of the contract (the reference tests compare
against literal lists): letters first (X < Y < Z, left-most letter most significant), then position combinations in
lexicographic order."""
from __future__ import annotations

import itertools as it

import numpy as np


def get_nontrivial_paulis(weight: int) -> list[str]:
    """All 3^weight words over X, Y, Z (empty list for weight 0)."""
    if weight <= 0:
        return []
    return ["".join(p) for p in it.product("XYZ", repeat=weight)]


def calculate_pauli_strings(n_qubits: int, weight: int):
    """All Pauli strings on ``n_qubits`` qubits with exactly ``weight`` non-identity letters."""
    from . import PauliString

    if weight == 0:
        return [PauliString("I" * n_qubits)]
    out = []
    combos = list(it.combinations(range(n_qubits), weight))
    for word in get_nontrivial_paulis(weight):
        for combo in combos:
            s = ["I"] * n_qubits
            for pos, ch in zip(combo, word):
                s[pos] = ch
            out.append(PauliString("".join(s)))
    return out


def calculate_pauli_strings_max_weight(n_qubits: int, weight: int):
    """All Pauli strings of weight <= ``weight``, by weight then in calculate_pauli_strings order."""
    out = []
    for w in range(weight + 1):
        out.extend(calculate_pauli_strings(n_qubits, w))
    return out


def pauli_string_sparse_repr(paulis) -> tuple[np.ndarray, np.ndarray]:
    """(k, m) with P[i, k[i]] = m[i] (reference: get_sparse_repr, __pauli_string.hpp:49-118) from the closed form
    k[i] = i ^ x, m[i] = (-i)^nY (-1)^popcount(i & z)."""
    string = "".join(str(p) for p in paulis) if not isinstance(paulis, str) else paulis
    n = len(string)
    dim = 1 << n if n else 0
    x = z = 0
    for q, ch in enumerate(string):
        bit = 1 << (n - 1 - q)
        x |= bit if ch in "XY" else 0
        z |= bit if ch in "YZ" else 0
    i = np.arange(dim, dtype=np.int64)
    par = np.zeros(dim, dtype=np.int64)
    zz = i & z
    while np.any(zz):
        par ^= zz & 1
        zz >>= 1
    m = np.array([1, -1j, -1, 1j])[string.count("Y") & 3] * (1 - 2 * par)
    return (i ^ x).astype(np.uint64), m.astype(np.complex128)
