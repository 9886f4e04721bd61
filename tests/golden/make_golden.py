"""Generate tests/golden/*.npz from the REFERENCE's own Python implementation.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

What is imported from the reference (never copied into this repo):
  * fast_pauli.pypauli.PauliString / PauliOp       (numpy implementation, pypauli/pauli_string.py:75-112,
                                                    pypauli/pauli_op.py:96-135)
  * fast_pauli.pypauli.helpers.naive_pauli_converter / naive_pauli_operator
                                                   (dense np.kron oracle, pypauli/helpers.py:40-85)
The package's __init__ needs the compiled nanobind module and qiskit, so the
sub-package is imported through a stub parent.

Inputs follow the reference's own test fixtures: states/coefficients from
default_rng(321) uniform [0,1) real and imaginary parts (tests/conftest.py:86-90), string sets from
tests/conftest.py:43-67 and fast_pauli/cpp/tests/test_pauli_string.cpp:255.
SummedPauliOp expectations are formed exactly as the reference's pytest files form their
"trusted" values (tests/fast_pauli/test_summed_pauli_op.py:83-87,128-132,160-177).
"""
from __future__ import annotations

import itertools as it
import os
import sys
import types

import numpy as np

REF = "/root/reference"
pkg = types.ModuleType("fast_pauli")
pkg.__path__ = [os.path.join(REF, "fast_pauli")]
sys.modules["fast_pauli"] = pkg
import fast_pauli.pypauli as pp  # noqa: E402
from fast_pauli.pypauli.helpers import naive_pauli_converter, naive_pauli_operator  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(321)


def rand_c(*shape):
    return rng.random(shape) + 1j * rng.random(shape)


def sample_pauli_strings() -> list[str]:
    strings = it.chain(
        ["I", "X", "Y", "Z"],
        it.product("IXYZ", repeat=2),
        it.product("IXYZ", repeat=3),
        ["XYZXYZ", "ZZZIII", "XYIZXYZ", "XXIYYIZZ", "ZIXIZYXX"],
        ["IXYZ", "YYIX", "XXYIYZ", "IZIXYYZ", "IZIXYYZIXYZ"],  # cpp/tests/test_pauli_string.cpp:255
    )
    return list(map("".join, strings))


def first_strings(size: int, limit: int) -> list[str]:
    out = []
    for s in it.product("IXYZ", repeat=size):
        if len(out) >= limit:
            break
        out.append("".join(s))
    return out


def weight_le2_strings(n: int) -> list[str]:
    """All strings of weight <= 2 in the order of calculate_pauli_strings_max_weight
    (__pauli_helpers.hpp:99-153): by weight, then letters-major, then position combination."""
    res = ["I" * n]
    for w in (1, 2):
        if w > n:
            break
        letters = ["".join(p) for p in it.product("XYZ", repeat=w)]
        combos = list(it.combinations(range(n), w))
        for let in letters:
            for combo in combos:
                s = ["I"] * n
                for pos, ch in zip(combo, let):
                    s[pos] = ch
                res.append("".join(s))
    return res


def gen_pauli_string() -> None:
    rec: dict[str, np.ndarray] = {}
    names = []
    for idx, s in enumerate(sample_pauli_strings()):
        n = len(s)
        ps = pp.PauliString(s)
        B = 3
        psi = rand_c(2**n, B)
        coeff = complex(rand_c(1)[0])
        cols, vals = pp.pauli_string.compose_sparse_pauli(s)
        dense = naive_pauli_converter(s)
        rec[f"{idx}_states"] = psi
        rec[f"{idx}_coeff"] = np.array(coeff)
        rec[f"{idx}_k"] = cols.astype(np.int64)
        rec[f"{idx}_m"] = vals
        rec[f"{idx}_apply2d"] = ps.apply(psi, coeff)
        rec[f"{idx}_apply1d"] = ps.apply(psi[:, 0].copy())
        rec[f"{idx}_expval"] = ps.expectation_value(psi)
        # dense cross-check inside the generator: the two reference oracles must agree
        np.testing.assert_allclose(rec[f"{idx}_apply2d"], coeff * dense @ psi, atol=1e-12)
        names.append(s)
    rec["strings"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "pauli_string.npz"), **rec)


def gen_pauli_op() -> None:
    rec: dict[str, np.ndarray] = {}
    cases = [(2, 16, 5), (3, 64, 4), (4, 100, 7), (7, 128, 3), (8, 200, 2), (10, 64, 16)]
    for idx, (n, limit, B) in enumerate(cases):
        if n <= 4:
            strings = first_strings(n, limit)
        else:
            letters = np.array(list("IXYZ"))
            strings = ["".join(letters[rng.integers(0, 4, size=n)]) for _ in range(limit)]
        coeffs = rand_c(len(strings))
        psi = rand_c(2**n, B)
        op = pp.PauliOp(coeffs, strings)
        rec[f"{idx}_strings"] = np.array(strings)
        rec[f"{idx}_coeffs"] = coeffs
        rec[f"{idx}_states"] = psi
        rec[f"{idx}_apply2d"] = op.apply(psi)
        rec[f"{idx}_apply1d"] = op.apply(psi[:, 0].copy())
        rec[f"{idx}_expval"] = op.expectation_value(psi)
        if n <= 8:
            dense = naive_pauli_operator(list(coeffs), strings)
            np.testing.assert_allclose(rec[f"{idx}_apply2d"], dense @ psi, atol=1e-10)
    rec["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "pauli_op.npz"), **rec)


def gen_summed_pauli_op() -> None:
    rec: dict[str, np.ndarray] = {}
    cases = [(1, 1, 1), (2, 10, 10), (6, 10, 7), (6, 3, 33), (5, 4, 8)]  # (n_qubits, n_operators, n_states)
    for idx, (n, K, B) in enumerate(cases):
        strings = weight_le2_strings(n)
        S = len(strings)
        coeffs = rand_c(S, K)
        psi = rand_c(2**n, B)
        data = rng.random((K, B))
        apply = np.zeros_like(psi)
        weighted = np.zeros_like(psi)
        expv = np.zeros((K, B), dtype=np.complex128)
        for k in range(K):
            a_k = pp.PauliOp(coeffs[:, k].copy(), strings)
            y = a_k.apply(psi)
            apply += y
            weighted += y * data[k]
            dense = naive_pauli_operator(list(coeffs[:, k]), strings)
            expv[k] = np.einsum("it,ij,jt->t", psi.conj(), dense, psi)
        rec[f"{idx}_strings"] = np.array(strings)
        rec[f"{idx}_coeffs"] = coeffs
        rec[f"{idx}_states"] = psi
        rec[f"{idx}_data"] = data
        rec[f"{idx}_apply"] = apply
        rec[f"{idx}_apply_weighted"] = weighted
        rec[f"{idx}_expval"] = expv
    rec["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "summed_pauli_op.npz"), **rec)


if __name__ == "__main__":
    gen_pauli_string()
    gen_pauli_op()
    gen_summed_pauli_op()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
