"""The reference's Python test cases for the hot path (PY_PS = tests/fast_pauli/test_pauli_string.py,
PY_PO = test_pauli_op.py, PY_SPO = test_summed_pauli_op.py, PY_P = test_pauli.py), restated against this package and
checked against dense ``np.kron`` algebra built here.

Every case runs three times:

* ``backend = "gpu"`` (``-m gpu``): the ctypes front-end over the real ``libfastpauli_b200.so`` on the B200 -- the
  parity test proper;
* ``backend = "native"`` (``-m gpu``): the same cases through the pybind11 module ``_fast_pauli`` over the C++ classes
  (``fast-pauli_b200/cpp/src/bindings.cpp``);
* ``backend = "mock"`` (``-m "not gpu"``): the ctypes front-end over ``tests/mock_abi.MockABI`` (the C ABI answered
  by the CPU oracle), which checks the HOST logic only -- dispatch on 1-D / 2-D input, implicit dtype conversion,
  output shapes, shape errors raised before a device is needed, coefficient orientation, plan invalidation.

``tests/run_reference_pytests.py`` runs the reference's own, unmodified pytest files against this package where
``/root/reference`` exists (202 passed, 6 skipped over the mock; see README.md).
"""
from __future__ import annotations

import itertools as it
import os
import pickle
import sys

import numpy as np
import pytest

import mock_abi
from conftest import ROOT
from __graft_entry__ import load_package

fp = load_package()

P2 = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
      "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0 + 0j, -1.0])}
TOL = dict(rtol=1e-12, atol=1e-12)


def kron(string: str) -> np.ndarray:
    m = np.ones((1, 1), dtype=complex)
    for ch in string:
        m = np.kron(m, P2[ch])
    return m


def dense_op(coeffs, strings) -> np.ndarray:
    return sum(c * kron(s) for c, s in zip(coeffs, strings))


def lexicographic(size: int, limit: int) -> list[str]:
    """First ``limit`` strings of ``size`` qubits in IXYZ product order (reference fixture conftest.py:57-69)."""
    return ["".join(s) for s in it.islice(it.product("IXYZ", repeat=size), limit)]


def sample_strings() -> list[str]:
    """All 1-3 qubit strings plus the five longer ones of the reference fixture (conftest.py:43-52)."""
    out = ["".join(s) for k in (1, 2, 3) for s in it.product("IXYZ", repeat=k)]
    return out + ["XYZXYZ", "ZZZIII", "XYIZXYZ", "XXIYYIZZ", "ZIXIZYXX"]


def load_native():
    """The pybind11 module, loaded from its in-tree location (built by `make -C fast-pauli_b200`)."""
    import glob
    import importlib.util

    hits = glob.glob(os.path.join(ROOT, "fast-pauli_b200", "_fast_pauli*.so"))
    if not hits:
        return None
    if "_fast_pauli" in sys.modules:
        return sys.modules["_fast_pauli"]
    spec = importlib.util.spec_from_file_location("_fast_pauli", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["_fast_pauli"] = mod
    return mod


@pytest.fixture(params=["mock", pytest.param("gpu", marks=pytest.mark.gpu),
                        pytest.param("native", marks=pytest.mark.gpu)])
def backend(request):
    """The module under test: every test binds it to the local name ``fp``."""
    if request.param == "mock":
        with mock_abi.installed(fp):
            yield fp
    elif request.param == "native":
        native = load_native()
        if native is None:
            pytest.skip("_fast_pauli extension not built")
        yield native
    else:
        yield fp


@pytest.fixture(params=["ctypes", "native"])
def host_backend(request):
    """Both front-ends with NO device and no mock: what they must get right before the GPU is touched."""
    if request.param == "native":
        native = load_native()
        if native is None:
            pytest.skip("_fast_pauli extension not built")
        return native
    return fp


@pytest.fixture
def rnd():
    rng = np.random.default_rng(321)  # the reference fixture's seed (conftest.py:87-90)
    return lambda *shape: rng.random(shape) + 1j * rng.random(shape)


# ----------------------------------------------------------------------------------------------- PauliString
def test_string_apply_1d(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PS:178-209: KATs, implicit conversion of float / int input, every sample string on a random state
    np.testing.assert_allclose(fp.PauliString("III").apply(np.arange(8)), np.arange(8), **TOL)  # int64 input
    np.testing.assert_allclose(fp.PauliString("ZYX").apply(np.ones(8)), kron("ZYX").sum(1), **TOL)  # float64 input
    r = fp.PauliString("IIII").apply(np.zeros(16))
    assert r.dtype == np.complex128 and r.shape == (16,) and not r.any()
    for s in sample_strings():
        psi = rnd(2 ** len(s))
        np.testing.assert_allclose(fp.PauliString(s).apply(psi), kron(s) @ psi, **TOL)
    # the binding drops `coeff` for a 1-D state (B_PS:142): same here
    psi = rnd(8)
    np.testing.assert_allclose(fp.PauliString("XYZ").apply(psi, 2.0), kron("XYZ") @ psi, **TOL)


def test_string_apply_batch(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PS:216-261: eye(8) -> the dense matrix; 42 states; 7 states with a complex coefficient
    np.testing.assert_allclose(fp.PauliString("ZYX").apply(np.eye(8)), kron("ZYX"), **TOL)
    for s in sample_strings():
        psis = rnd(2 ** len(s), 42)
        np.testing.assert_allclose(fp.PauliString(s).apply(psis), kron(s) @ psis, **TOL)
        coeff = complex(rnd(1)[0])
        psis = rnd(2 ** len(s), 7)
        np.testing.assert_allclose(fp.PauliString(s).apply(psis, coeff), coeff * (kron(s) @ psis), **TOL)
        np.testing.assert_allclose(fp.PauliString(s).apply(states=psis, coeff=coeff), coeff * (kron(s) @ psis), **TOL)


def test_string_expectation_value(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PS:268-315: III on arange(8) = sum k^2; 1-D -> shape (1,); 21 states with a coefficient
    ev = fp.PauliString("III").expectation_value(np.arange(8))
    assert ev.shape == (1,)
    np.testing.assert_allclose(ev, [np.sum(np.arange(8) ** 2)], **TOL)
    for s in sample_strings():
        D = kron(s)
        psi = rnd(2 ** len(s))
        np.testing.assert_allclose(fp.PauliString(s).expectation_value(psi), [np.vdot(psi, D @ psi)], **TOL)
        psis = rnd(2 ** len(s), 21)
        coeff = complex(rnd(1)[0])
        got = fp.PauliString(s).expectation_value(psis, coeff)
        assert got.shape == (21,)
        np.testing.assert_allclose(got, coeff * np.einsum("it,ij,jt->t", psis.conj(), D, psis), **TOL)


def test_string_exceptions_need_no_device(host_backend):
    fp = host_backend  # noqa: F841 - the module under test
    # PY_PS:388-406; shape errors are ValueErrors whether or not a device exists
    with pytest.raises(ValueError):
        fp.PauliString("ABC")
    with pytest.raises(ValueError):
        fp.PauliString("II").apply(np.array([0.1, 0.2, 0.3]))
    with pytest.raises(ValueError):
        fp.PauliString("XYZ").apply(np.eye(4))
    with pytest.raises(ValueError):
        fp.PauliString("XYZ").expectation_value(np.ones((4, 4)))
    with pytest.raises(ValueError):
        fp.PauliString("XYZ").apply(np.ones((8, 2, 2)))
    with pytest.raises(ValueError):
        fp.PauliString("XY").apply(np.ones((4, 6), dtype=complex)[:, ::2])  # not row-major contiguous (NB:65-73)
    with pytest.raises(AttributeError):
        fp.PauliString("XYZ").dim = 99
    with pytest.raises(AttributeError):
        fp.PauliString("XYZ").weight = 99


# ----------------------------------------------------------------------------------------------- PauliOp
OP_STRING_SETS = [lexicographic(2, 16), lexicographic(3, 64), lexicographic(4, 256)[::3], lexicographic(7, 128),
                  lexicographic(8, 200)[:100], lexicographic(8, 200)[100:], lexicographic(10, 64)]


def test_op_apply_1d(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PO:175-215
    np.testing.assert_allclose(fp.PauliOp([0.5, 0.5], ["III", "III"]).apply(np.arange(8)), np.arange(8), **TOL)
    for strings in OP_STRING_SETS:
        coeffs = rnd(len(strings))
        psi = rnd(2 ** len(strings[0]))
        np.testing.assert_allclose(fp.PauliOp(coeffs, strings).apply(psi), dense_op(coeffs, strings) @ psi, **TOL)


def test_op_apply_batch(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PO:222-262: a random number (< 100) of states per string set
    np.testing.assert_allclose(fp.PauliOp([0.5, 0.5], ["III", "III"]).apply(np.eye(8)), np.eye(8), **TOL)
    for strings in OP_STRING_SETS:
        coeffs = rnd(len(strings))
        n_states = 1 + int(99 * rnd(1)[0].real)
        psis = rnd(2 ** len(strings[0]), n_states)
        got = fp.PauliOp(coeffs, strings).apply(psis)
        assert got.shape == psis.shape and got.dtype == np.complex128
        np.testing.assert_allclose(got, dense_op(coeffs, strings) @ psis, **TOL)


def test_op_expectation_value(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # PY_PO:269-328: two identities on arange(8) = 2 sum k^2; 1-D -> (1,), batch -> (n_states,)
    ev = fp.PauliOp([1, 1], ["III", "III"]).expectation_value(np.arange(8))
    np.testing.assert_allclose(ev, [2 * np.sum(np.arange(8) ** 2)], **TOL)
    for strings in OP_STRING_SETS:
        coeffs = rnd(len(strings))
        D = dense_op(coeffs, strings)
        op = fp.PauliOp(coeffs, strings)
        psi = rnd(D.shape[0])
        np.testing.assert_allclose(op.expectation_value(psi), [np.vdot(psi, D @ psi)], **TOL)
        psis = rnd(D.shape[0], 1 + int(99 * rnd(1)[0].real))
        np.testing.assert_allclose(op.expectation_value(psis), np.einsum("it,ij,jt->t", psis.conj(), D, psis), **TOL)


def test_op_ctor_forms_and_plan_invalidation(backend, rnd):
    fp = backend  # noqa: F841 - the module under test
    # numpy arrays of coefficients / strings (PY_PO:549), PauliString items, strings-only ctor (PO:59-80);
    # scale / extend must rebuild the device plan (the plan caches the coefficients)
    strings = lexicographic(3, 20)
    coeffs = rnd(20)
    psi = rnd(8, 5)
    want = dense_op(coeffs, strings) @ psi
    for op in (fp.PauliOp(coeffs, np.array(strings)), fp.PauliOp(list(coeffs), [fp.PauliString(s) for s in strings]),
               pickle.loads(pickle.dumps(fp.PauliOp(coeffs, strings))), fp.PauliOp(coeffs, strings).clone()):
        np.testing.assert_allclose(op.apply(psi), want, **TOL)
    np.testing.assert_allclose(fp.PauliOp(strings).apply(psi), dense_op(np.ones(20), strings) @ psi, **TOL)
    op = fp.PauliOp(coeffs, strings)
    op.apply(psi)
    op.scale(2.0)
    np.testing.assert_allclose(op.apply(psi), 2 * want, **TOL)
    op.extend(fp.PauliString("ZZZ"), 0.5j, dedupe=False)
    np.testing.assert_allclose(op.apply(psi), 2 * want + 0.5j * (kron("ZZZ") @ psi), **TOL)
    prod = fp.PauliOp(coeffs, strings) @ fp.PauliOp(coeffs[:7], strings[:7])
    np.testing.assert_allclose(prod.apply(psi), dense_op(coeffs, strings) @ (dense_op(coeffs[:7], strings[:7]) @ psi),
                               rtol=1e-11, atol=1e-11)


def test_op_exceptions_need_no_device(host_backend):
    fp = host_backend  # noqa: F841 - the module under test
    # PY_PO:886-946
    with pytest.raises(ValueError):
        fp.PauliOp([1, 2], ["XYZ"])
    with pytest.raises(ValueError):
        fp.PauliOp([1, 2], ["XYZ", "XY"])
    with pytest.raises(ValueError):
        fp.PauliOp([1, 1, 1], ["X", "Y", "Z"]).apply(np.ones(3))
    with pytest.raises(ValueError):
        fp.PauliOp([1, 1, 1], ["X", "Y", "Z"]).apply(np.eye(4))
    with pytest.raises(ValueError):
        fp.PauliOp([1, 1], ["XYZ", "ZYX"]).expectation_value(np.ones(32))
    with pytest.raises(ValueError):
        fp.PauliOp([1, 1], ["XYZ", "ZYX"]).expectation_value(np.eye(16))
    with pytest.raises(ValueError):
        fp.PauliOp([-1, 1], ["IX", "YZ"]) - fp.PauliOp([1], ["X"])
    with pytest.raises(ValueError):
        fp.PauliOp([-1, 1], ["IX", "YZ"]).extend(fp.PauliString("XYZ"), 1j, dedupe=False)


# ----------------------------------------------------------------------------------------------- SummedPauliOp
SOP_SHAPES = list(it.product([1, 10, 1000], [1, 10, 100], [1, 2, 6]))  # n_states, n_operators, n_qubits (PY_SPO:50-53)


def _sop_case(n_states, n_operators, n_qubits, rng):
    strings = [str(s) for s in fp.helpers.calculate_pauli_strings_max_weight(n_qubits, 2)]
    coeffs = rng.random((len(strings), n_operators)) + 1j * rng.random((len(strings), n_operators))
    dense = np.stack([kron(s) for s in strings])
    return strings, coeffs, dense


@pytest.mark.parametrize("n_states,n_operators,n_qubits", SOP_SHAPES)
def test_sop_apply_weighted_expval(backend, n_states, n_operators, n_qubits):
    fp = backend  # noqa: F841 - the module under test
    # PY_SPO:54-180: apply, apply_weighted and expectation_value on the same operator; n_states == 1 goes in 1-D
    rng = np.random.default_rng(n_states * 1000 + n_operators * 10 + n_qubits)
    strings, coeffs, dense = _sop_case(n_states, n_operators, n_qubits, rng)
    op = fp.SummedPauliOp(strings, coeffs)
    dim = 2**n_qubits
    psi = rng.random((dim, n_states)).astype(np.complex128)
    data = rng.random((n_operators, n_states))
    one_d = n_states == 1
    arg, darg = (psi[:, 0].copy(), data[:, 0].copy()) if one_d else (psi, data)

    A = np.einsum("sk,sij->kij", coeffs, dense)  # (K, dim, dim)
    want_apply = np.einsum("kij,jt->it", A, psi)
    want_weighted = np.einsum("kij,kt,jt->it", A, data, psi)
    want_ev = np.einsum("it,kij,jt->kt", psi.conj(), A, psi)

    got = op.apply(arg)
    assert got.shape == arg.shape
    np.testing.assert_allclose(got.reshape(dim, n_states), want_apply, rtol=1e-11, atol=1e-11)
    got = op.apply_weighted(arg, darg)
    assert got.shape == arg.shape
    np.testing.assert_allclose(got.reshape(dim, n_states), want_weighted, rtol=1e-11, atol=1e-11)
    got = op.expectation_value(arg)
    assert got.shape == ((n_operators,) if one_d else (n_operators, n_states))
    np.testing.assert_allclose(got.reshape(n_operators, n_states), want_ev, rtol=1e-11, atol=1e-11)


def test_sop_coeffs_orientation_setter_and_square(backend):
    fp = backend  # noqa: F841 - the module under test
    # PY_SPO:227-241, 270-300: the getter / setter use (n_operators, n_pauli_strings), the ctor (n_strings, n_operators);
    # assigning coefficients must invalidate the device plan; square() against the dense square
    rng = np.random.default_rng(5)
    strings, coeffs, dense = _sop_case(4, 3, 3, rng)
    op = fp.SummedPauliOp(strings, coeffs)
    assert op.coeffs.shape == (3, len(strings))
    np.testing.assert_array_equal(op.coeffs, coeffs.T)
    psi = rng.random((8, 4)) + 1j * rng.random((8, 4))
    op.apply(psi)  # builds the plan
    new = rng.random((3, len(strings))) + 0j
    op.coeffs = new
    np.testing.assert_array_equal(op.coeffs, new)
    np.testing.assert_allclose(op.apply(psi), np.einsum("ks,sij,jt->it", new, dense, psi), rtol=1e-11, atol=1e-11)
    with pytest.raises(ValueError):
        op.coeffs = new.T
    with pytest.raises(ValueError):
        op.apply_weighted(psi, np.ones((2, 4)))
    with pytest.raises(ValueError):
        op.apply(np.ones((4, 4)))
    sq = op.square()
    A = np.einsum("ks,sij->kij", new, dense)
    np.testing.assert_allclose(sq.to_tensor(), A @ A, rtol=1e-11, atol=1e-11)
    clone = pickle.loads(pickle.dumps(op))
    np.testing.assert_allclose(clone.expectation_value(psi), op.expectation_value(psi), **TOL)


def test_native_module_host_surface(host_backend):
    """Value types, operator algebra, dense exports, generators and pickling of both front-ends against dense
    algebra (no device needed): PY_P:28-100, PY_PS:62-172, 321-383, PY_PO:59-170, 337-880, PY_SPO:28-45, 190-300."""
    fp = host_backend  # noqa: F841 - the module under test
    for a, b in it.product("IXYZ", repeat=2):
        phase, p = fp.Pauli(a) @ fp.Pauli(b)
        np.testing.assert_allclose(phase * np.asarray(p.to_tensor()), P2[a] @ P2[b], atol=1e-15)
    assert str(fp.Pauli(code=2)) == "Y" and str(fp.Pauli(symbol="Z")) == "Z" and str(fp.Pauli()) == "I"
    with pytest.raises(TypeError):
        fp.Pauli("II")
    for bad in ("A", -1, 5):
        with pytest.raises(ValueError):
            fp.Pauli(bad)
    assert str(pickle.loads(pickle.dumps(fp.Pauli("X")))) == "X"
    empty = fp.PauliString()
    assert empty.n_qubits == 0 and empty.dim == 0 and empty.weight == 0
    ps = fp.PauliString([fp.Pauli("X"), fp.Pauli("I"), fp.Pauli("Y")])
    assert str(ps) == "XIY" and ps.weight == 2 and ps.dim == 8 and str(ps.clone()) == "XIY"
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5):
        a, b = ("".join(rng.choice(list("IXYZ"), n)) for _ in range(2))
        np.testing.assert_allclose(fp.PauliString(a).to_tensor(), kron(a), atol=1e-15)
        phase, prod = fp.PauliString(a) @ fp.PauliString(b)
        np.testing.assert_allclose(phase * kron(str(prod)), kron(a) @ kron(b), atol=1e-14)
        np.testing.assert_allclose((fp.PauliString(a) + fp.PauliString(b)).to_tensor(), kron(a) + kron(b), atol=1e-14)
        np.testing.assert_allclose((fp.PauliString(a) - fp.PauliString(b)).to_tensor(), kron(a) - kron(b), atol=1e-14)
    assert str(pickle.loads(pickle.dumps(fp.PauliString("XYZ")))) == "XYZ"
    assert fp.PauliOp().dim == 0 and fp.PauliOp().n_pauli_strings == 0
    sa, sb = lexicographic(3, 20), lexicographic(3, 64)[30:45]
    ca, cb = rng.random(20) + 1j * rng.random(20), rng.random(15) + 1j * rng.random(15)
    A, B = fp.PauliOp(ca, sa), fp.PauliOp(list(cb), [fp.PauliString(s) for s in sb])
    DA, DB = dense_op(ca, sa), dense_op(cb, sb)
    assert A.dim == 8 and A.n_qubits == 3 and A.n_pauli_strings == 20 and A.pauli_strings_as_str == sa
    np.testing.assert_allclose(np.asarray(A.coeffs), ca)
    np.testing.assert_allclose(A.to_tensor(), DA, atol=1e-13)
    np.testing.assert_allclose((A @ B).to_tensor(), DA @ DB, atol=1e-12)
    np.testing.assert_allclose((A @ fp.PauliString("XYZ")).to_tensor(), DA @ kron("XYZ"), atol=1e-13)
    np.testing.assert_allclose((fp.PauliString("XYZ") @ A).to_tensor(), kron("XYZ") @ DA, atol=1e-13)
    np.testing.assert_allclose((A + B).to_tensor(), DA + DB, atol=1e-13)
    np.testing.assert_allclose((A - B).to_tensor(), DA - DB, atol=1e-13)
    np.testing.assert_allclose((A + fp.PauliString("ZZZ")).to_tensor(), DA + kron("ZZZ"), atol=1e-13)
    np.testing.assert_allclose((fp.PauliString("ZZZ") - A).to_tensor(), kron("ZZZ") - DA, atol=1e-13)
    np.testing.assert_allclose((A * 2j).to_tensor(), 2j * DA, atol=1e-13)
    np.testing.assert_allclose((0.5 * A).to_tensor(), 0.5 * DA, atol=1e-13)
    C = A.clone()
    C += B
    C -= fp.PauliString("III")
    C *= 3.0
    np.testing.assert_allclose(C.to_tensor(), 3 * (DA + DB - np.eye(8)), atol=1e-12)
    C = A.clone()
    C.scale(np.arange(20, dtype=complex))
    np.testing.assert_allclose(C.to_tensor(), dense_op(ca * np.arange(20), sa), atol=1e-12)
    C.extend(fp.PauliString("III"), 2.0, dedupe=True)  # "III" is the first string: merged, not appended
    assert C.n_pauli_strings == 20
    C.extend(B)
    assert C.n_pauli_strings == 35
    np.testing.assert_allclose(pickle.loads(pickle.dumps(A)).to_tensor(), DA, atol=1e-13)
    strings = [str(s) for s in fp.helpers.calculate_pauli_strings_max_weight(3, 2)]
    assert len(strings) == 37 and strings[0] == "III" and len(fp.helpers.calculate_pauli_strings(4, 2)) == 54
    assert list(fp.helpers.get_nontrivial_paulis(2)) == ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
    k, mvals = fp.helpers.pauli_string_sparse_repr([fp.Pauli(c) for c in "XYZ"])
    dense = np.zeros((8, 8), dtype=complex)
    dense[np.arange(8), np.asarray(k)] = np.asarray(mvals)
    np.testing.assert_allclose(dense, kron("XYZ"), atol=1e-15)
    hk = rng.random((37, 3)) + 1j * rng.random((37, 3))
    sop = fp.SummedPauliOp(strings, hk)
    assert (sop.dim, sop.n_operators, sop.n_pauli_strings) == (8, 3, 37) and sop.pauli_strings_as_str == strings
    np.testing.assert_array_equal(sop.coeffs, hk.T)
    want = np.einsum("sk,sij->kij", hk, np.stack([kron(s) for s in strings]))
    np.testing.assert_allclose(sop.to_tensor(), want, atol=1e-12)
    for k_op, part in enumerate(sop.split()):
        np.testing.assert_allclose(part.to_tensor(), want[k_op], atol=1e-12)
    np.testing.assert_allclose(pickle.loads(pickle.dumps(sop)).to_tensor(), want, atol=1e-12)
    np.testing.assert_allclose(sop.clone().to_tensor(), want, atol=1e-12)
    sop2 = fp.SummedPauliOp([fp.PauliString(s) for s in strings], hk)
    np.testing.assert_array_equal(sop2.coeffs, hk.T)
    with pytest.raises(ValueError):
        fp.SummedPauliOp(strings, hk[:5])
    with pytest.raises(ValueError):
        sop.coeffs = hk  # wrong orientation


def test_front_ends_fail_loudly_without_a_device(host_backend):
    """No CPU fallback in either front-end: a hot-path call on a GPU-less host raises RuntimeError."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is for GPU-less hosts")
    fp = host_backend  # noqa: F841
    psi = np.ones((4, 3), dtype=np.complex128)
    for call in (lambda: fp.PauliString("XY").apply(psi), lambda: fp.PauliString("XY").expectation_value(psi),
                 lambda: fp.PauliOp([1, 2], ["XY", "ZZ"]).apply(psi),
                 lambda: fp.PauliOp([1, 2], ["XY", "ZZ"]).expectation_value(psi),
                 lambda: fp.SummedPauliOp(["XY", "ZZ"], np.ones((2, 2))).apply(psi),
                 lambda: fp.SummedPauliOp(["XY", "ZZ"], np.ones((2, 2))).apply_weighted(psi, np.ones((2, 3))),
                 lambda: fp.SummedPauliOp(["XY", "ZZ"], np.ones((2, 2))).expectation_value(psi),
                 lambda: fp.SummedPauliOp(["XY", "ZZ"], np.ones((2, 2))).square()):
        with pytest.raises(RuntimeError):
            call()


def test_mock_is_not_the_product():
    """Outside the fixture the package talks to the real shared library again."""
    import ctypes

    assert isinstance(fp.lib, ctypes.CDLL)
