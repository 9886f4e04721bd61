"""The reference's Python test cases for the hot path (PY_PS = tests/fast_pauli/test_pauli_string.py,
PY_PO = test_pauli_op.py, PY_SPO = test_summed_pauli_op.py, PY_P = test_pauli.py), restated against this package and
checked against dense ``np.kron`` algebra built here.

Every case runs twice:

* ``backend = "gpu"`` (``-m gpu``): the real ``libfastpauli_b200.so`` on the B200 -- the parity test proper;
* ``backend = "mock"`` (``-m "not gpu"``): the Python front-end over ``tests/mock_abi.MockABI`` (the C ABI answered
  by the CPU oracle), which checks the HOST logic only -- dispatch on 1-D / 2-D input, implicit dtype conversion,
  output shapes, shape errors raised before a device is needed, coefficient orientation, plan invalidation.

``scripts/run_reference_pytests.py`` runs the reference's own, unmodified pytest files against this package where
``/root/reference`` exists (202 passed, 6 skipped over the mock; see README.md).
"""
from __future__ import annotations

import itertools as it
import pickle

import numpy as np
import pytest

import mock_abi
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


@pytest.fixture(params=["mock", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "mock":
        with mock_abi.installed(fp) as mock:
            yield mock
    else:
        yield None


@pytest.fixture
def rnd():
    rng = np.random.default_rng(321)  # the reference fixture's seed (conftest.py:87-90)
    return lambda *shape: rng.random(shape) + 1j * rng.random(shape)


# ----------------------------------------------------------------------------------------------- PauliString
def test_string_apply_1d(backend, rnd):
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


def test_string_exceptions_need_no_device(backend):
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
    # PY_PO:175-215
    np.testing.assert_allclose(fp.PauliOp([0.5, 0.5], ["III", "III"]).apply(np.arange(8)), np.arange(8), **TOL)
    for strings in OP_STRING_SETS:
        coeffs = rnd(len(strings))
        psi = rnd(2 ** len(strings[0]))
        np.testing.assert_allclose(fp.PauliOp(coeffs, strings).apply(psi), dense_op(coeffs, strings) @ psi, **TOL)


def test_op_apply_batch(backend, rnd):
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


def test_op_exceptions_need_no_device(backend):
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


def test_mock_is_not_the_product():
    """Outside the fixture the package talks to the real shared library again."""
    import ctypes

    assert isinstance(fp.lib, ctypes.CDLL)
