"""CPU tests of the Python API surface below the hot path (no GPU): value types, operator algebra, dense exports,
generators, pickling -- mirrored from the reference bindings and checked against dense numpy algebra built here
(np.kron of the 2x2 matrices) and, for the sparse representation, against golden vectors from the reference."""
from __future__ import annotations

import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN
from __graft_entry__ import load_package

fp = load_package()

P2 = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0, -1.0])}


def kron(string: str) -> np.ndarray:
    m = np.array([[1.0 + 0j]])
    for ch in string:
        m = np.kron(m, P2[ch])
    return m


def test_pauli_value_type():
    for a in "IXYZ":
        for b in "IXYZ":
            phase, p = fp.Pauli(a) @ fp.Pauli(b)
            np.testing.assert_allclose(phase * P2[str(p)], P2[a] @ P2[b])
    assert str(fp.Pauli(2)) == "Y" and fp.Pauli("Z").code == 3 and fp.Pauli().code == 0
    np.testing.assert_allclose(fp.Pauli("Y").to_tensor(), P2["Y"])
    with pytest.raises(ValueError):
        fp.Pauli(4)
    with pytest.raises(ValueError):
        fp.Pauli("Q")
    assert pickle.loads(pickle.dumps(fp.Pauli("X"))) == fp.Pauli("X")


def test_pauli_string_algebra_and_dense():
    rng = np.random.default_rng(0)
    for _ in range(30):
        n = int(rng.integers(1, 6))
        a = "".join(rng.choice(list("IXYZ"), n))
        b = "".join(rng.choice(list("IXYZ"), n))
        np.testing.assert_allclose(fp.PauliString(a).to_tensor(), kron(a))
        phase, prod = fp.PauliString(a) @ fp.PauliString(b)
        np.testing.assert_allclose(phase * kron(str(prod)), kron(a) @ kron(b), atol=1e-14)
        np.testing.assert_allclose((fp.PauliString(a) + fp.PauliString(b)).to_tensor(), kron(a) + kron(b))
        np.testing.assert_allclose((fp.PauliString(a) - fp.PauliString(b)).to_tensor(), kron(a) - kron(b), atol=1e-14)
    assert fp.PauliString([fp.Pauli("X"), fp.Pauli(3)]).string == "XZ"
    with pytest.raises(ValueError):
        fp.PauliString("XX") @ fp.PauliString("XXX")
    ps = pickle.loads(pickle.dumps(fp.PauliString("XYZI")))
    assert ps == fp.PauliString("XYZI") and ps.weight == 3


def test_sparse_repr_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "pauli_string.npz"))
    for idx, s in enumerate(g["strings"]):
        k, m = fp.helpers.pauli_string_sparse_repr(str(s))
        np.testing.assert_array_equal(k.astype(np.int64), g[f"{idx}_k"])
        np.testing.assert_array_equal(m, g[f"{idx}_m"])


def test_pauli_op_algebra():
    rng = np.random.default_rng(1)
    n = 3
    sa = ["".join(rng.choice(list("IXYZ"), n)) for _ in range(5)]
    sb = ["".join(rng.choice(list("IXYZ"), n)) for _ in range(4)]
    ca = rng.random(5) + 1j * rng.random(5)
    cb = rng.random(4) + 1j * rng.random(4)
    A, B = fp.PauliOp(ca, sa), fp.PauliOp(cb, sb)
    dA = sum(c * kron(s) for c, s in zip(ca, sa))
    dB = sum(c * kron(s) for c, s in zip(cb, sb))
    np.testing.assert_allclose(A.to_tensor(), dA)
    np.testing.assert_allclose((A @ B).to_tensor(), dA @ dB, atol=1e-13)
    assert (A @ B).n_pauli_strings <= 20
    ps = fp.PauliString("XYZ")
    np.testing.assert_allclose((A @ ps).to_tensor(), dA @ kron("XYZ"), atol=1e-13)
    np.testing.assert_allclose((ps @ A).to_tensor(), kron("XYZ") @ dA, atol=1e-13)
    np.testing.assert_allclose((A + B).to_tensor(), dA + dB)
    np.testing.assert_allclose((A - B).to_tensor(), dA - dB, atol=1e-14)
    np.testing.assert_allclose((A + ps).to_tensor(), dA + kron("XYZ"))
    np.testing.assert_allclose((ps + A).to_tensor(), dA + kron("XYZ"))
    np.testing.assert_allclose((ps - A).to_tensor(), kron("XYZ") - dA, atol=1e-14)
    np.testing.assert_allclose((A * 2j).to_tensor(), 2j * dA)
    np.testing.assert_allclose((0.5 * A).to_tensor(), 0.5 * dA)
    C = A.clone()
    C += B
    C -= ps
    C *= 3.0
    np.testing.assert_allclose(C.to_tensor(), 3 * (dA + dB - kron("XYZ")), atol=1e-13)
    D = fp.PauliOp([1.0], ["XYZ"])
    D.extend(fp.PauliString("XYZ"), 2.0, dedupe=True)
    assert D.n_pauli_strings == 1 and D.coeffs[0] == 3.0
    D.extend(fp.PauliString("XYZ"), 2.0, dedupe=False)
    assert D.n_pauli_strings == 2
    with pytest.raises(ValueError):
        A @ fp.PauliOp([1.0], ["XX"])
    E = pickle.loads(pickle.dumps(A))
    np.testing.assert_allclose(E.to_tensor(), dA)
    assert fp.PauliOp(["XX", "YY"]).coeffs.tolist() == [1, 1]  # strings-only constructor (PO:59-80)


def test_helpers_enumeration_order():
    h = fp.helpers
    assert h.get_nontrivial_paulis(0) == []
    assert h.get_nontrivial_paulis(2) == ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
    assert [str(p) for p in h.calculate_pauli_strings(2, 1)] == ["XI", "IX", "YI", "IY", "ZI", "IZ"]
    assert [str(p) for p in h.calculate_pauli_strings(3, 0)] == ["III"]
    allp = h.calculate_pauli_strings_max_weight(3, 3)
    assert len(allp) == 64 and len({str(p) for p in allp}) == 64
    assert len(h.calculate_pauli_strings_max_weight(4, 2)) == 1 + 12 + 54


def test_summed_pauli_op_host_helpers():
    rng = np.random.default_rng(2)
    strings = [str(p) for p in fp.helpers.calculate_pauli_strings_max_weight(3, 1)]
    coeffs = rng.random((len(strings), 4)) + 1j * rng.random((len(strings), 4))
    sop = fp.SummedPauliOp(strings, coeffs)
    dense = sop.to_tensor()
    for k, op in enumerate(sop.split()):
        np.testing.assert_allclose(op.to_tensor(), dense[k])
    # square() contracts on the GPU (fp_sop_square): covered by tests/test_gpu_parity.py::test_summed_pauli_op_square
    assert sop.pauli_strings_as_str == strings
    cl = pickle.loads(pickle.dumps(sop))
    np.testing.assert_allclose(cl.to_tensor(), dense)
    assert sop.clone().coeffs.shape == (4, len(strings))


def test_import_fast_pauli_alias_package():
    """`import fast_pauli` keeps working for code written against the reference package (fast_pauli/__init__.py:18-25)
    once fast-pauli_b200/compat is on the path."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import fast_pauli as fp, pickle\n"
        "from fast_pauli.helpers import calculate_pauli_strings_max_weight\n"
        "assert fp.PauliString('XYZ').dim == 8 and len(calculate_pauli_strings_max_weight(3, 2)) == 37\n"
        "assert fp.PauliOp([1, 2], ['XI', 'IZ']).n_pauli_strings == 2\n"
        "assert pickle.loads(pickle.dumps(fp.PauliString('XY'))) == fp.PauliString('XY')\n"
        "try:\n    fp.from_qiskit(None)\nexcept NotImplementedError:\n    print('ok')\n"
    )
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "fast-pauli_b200", "compat"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]
