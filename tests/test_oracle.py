"""CPU tests (no GPU): pin the oracle.

The plain-C restatement (oracle/pauli_oracle.c) is checked against
  * the unmodified reference compiled from /root/reference (oracle/_ref), seq and par,
  * golden vectors produced by the reference's numpy implementation (tests/golden/*.npz),
  * the reference tests' known-answer cases.
"""
from __future__ import annotations

import os

import numpy as np
import pytest

from conftest import GOLDEN, TOL, rand_states, rand_strings, rel_err
from oracle import oracle as orc

PORT = orc.port()
REF = orc.reference()
BACKENDS = [pytest.param(PORT, id="port")] + ([pytest.param(REF, id="reference")] if REF else [])
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference here)")


# ------------------------------------------------------------------ golden vectors (numpy reference)
@pytest.mark.parametrize("be", BACKENDS)
def test_golden_pauli_string(be):
    g = np.load(os.path.join(GOLDEN, "pauli_string.npz"))
    for idx, s in enumerate(g["strings"]):
        s = str(s)
        psi = g[f"{idx}_states"]
        c = complex(g[f"{idx}_coeff"])
        k, m = orc.np_sparse(s)
        np.testing.assert_array_equal(k, g[f"{idx}_k"])  # PY_PS:447-453 pins (k, m)
        np.testing.assert_array_equal(m, g[f"{idx}_m"])
        assert rel_err(be.string_apply(s, psi, c), g[f"{idx}_apply2d"]) < 1e-14
        assert rel_err(be.string_apply(s, psi[:, 0].copy()), g[f"{idx}_apply1d"]) < 1e-14
        assert rel_err(be.string_expval(s, psi), g[f"{idx}_expval"]) < 1e-13


@pytest.mark.parametrize("be", BACKENDS)
def test_golden_pauli_op(be):
    g = np.load(os.path.join(GOLDEN, "pauli_op.npz"))
    for idx in range(int(g["n_cases"])):
        strings = [str(s) for s in g[f"{idx}_strings"]]
        h, psi = g[f"{idx}_coeffs"], g[f"{idx}_states"]
        assert rel_err(be.op_apply(strings, h, psi), g[f"{idx}_apply2d"]) < 1e-13
        assert rel_err(be.op_apply(strings, h, psi[:, 0].copy()), g[f"{idx}_apply1d"]) < 1e-13
        assert rel_err(be.op_expval(strings, h, psi), g[f"{idx}_expval"]) < 1e-13


@pytest.mark.parametrize("be", BACKENDS)
def test_golden_summed_pauli_op(be):
    g = np.load(os.path.join(GOLDEN, "summed_pauli_op.npz"))
    for idx in range(int(g["n_cases"])):
        strings = [str(s) for s in g[f"{idx}_strings"]]
        h, psi, data = g[f"{idx}_coeffs"], g[f"{idx}_states"], g[f"{idx}_data"]
        assert rel_err(be.sop_apply(strings, h, psi), g[f"{idx}_apply"]) < 1e-13
        assert rel_err(be.sop_apply_weighted(strings, h, psi, data), g[f"{idx}_apply_weighted"]) < 1e-13
        assert rel_err(be.sop_expval(strings, h, psi), g[f"{idx}_expval"]) < 1e-13


# ------------------------------------------------------------------ known-answer tests of the reference
@pytest.mark.parametrize("be", BACKENDS)
def test_kat_identity_and_ixi(be):
    # T_PS:217-229 "IIII" on all-ones; T_PS:231-247 "IXI" on e6+e7
    ones = np.ones(16, dtype=np.complex128)
    np.testing.assert_array_equal(be.string_apply("IIII", ones), ones)
    st = np.zeros(8, dtype=np.complex128)
    st[6] = st[7] = 1
    exp = np.zeros(8, dtype=np.complex128)
    exp[4] = exp[5] = 1
    np.testing.assert_array_equal(be.string_apply("IXI", st), exp)


@pytest.mark.parametrize("be", BACKENDS)
def test_kat_python_cases(be):
    # PY_PS:184-200: III on arange -> identity; ZYX on ones -> dense row sums; ZYX on eye -> dense
    np.testing.assert_array_equal(be.string_apply("III", np.arange(8).astype(np.complex128)), np.arange(8))
    dense = np.zeros((8, 8), dtype=np.complex128)
    k, m = orc.np_sparse("ZYX")
    dense[np.arange(8), k] = m
    np.testing.assert_allclose(be.string_apply("ZYX", np.ones(8, dtype=np.complex128)), dense.sum(axis=1))
    np.testing.assert_allclose(be.string_apply("ZYX", np.eye(8, dtype=np.complex128)), dense)
    # PY_PS:285-295 expectation of III on arange(8) = sum k^2; PY_PO:286-297 doubled for two identities
    ar = np.arange(8).astype(np.complex128)
    assert be.string_expval("III", ar)[0] == pytest.approx(140.0)
    assert be.op_expval(["III", "III"], [1, 1], ar)[0] == pytest.approx(280.0)
    # PY_PO:186-190 [0.5,0.5] x ["III","III"] is the identity
    np.testing.assert_allclose(be.op_apply(["III", "III"], [0.5, 0.5], ar), ar)


@pytest.mark.parametrize("be", BACKENDS)
def test_kat_sixteen_identities(be):
    # T_PO:228-267: 16 identical IIII strings with coeff 1/16 -> identity on random states
    rng = np.random.default_rng(18)
    psi = rand_states(rng, 16, 10)
    out = be.op_apply(["IIII"] * 16, [1 / 16] * 16, psi)
    np.testing.assert_allclose(out, psi, atol=1e-15)


@pytest.mark.parametrize("be", BACKENDS)
def test_kat_ixyz_literal_state(be):
    # T_PO:159-162: PauliOp({1}, {"IXYZ"}) equals PauliString("IXYZ") on a fixed 16-amplitude state
    st = (np.arange(16) * 0.25 + 1j * (np.arange(16) % 5) * 0.5).astype(np.complex128)
    np.testing.assert_array_equal(be.op_apply(["IXYZ"], [1.0], st), be.string_apply("IXYZ", st))


@pytest.mark.parametrize("be", BACKENDS)
def test_accumulate_semantics(be):
    # the C++ methods do `+=` into the caller's buffer (PS:419,432; PS:523,534)
    rng = np.random.default_rng(3)
    psi = rand_states(rng, 8, 4)
    base = rand_states(rng, 8, 4)
    out = be.string_apply("XYZ", psi, 0.5 - 2j, out=base.copy())
    np.testing.assert_allclose(out - base, be.string_apply("XYZ", psi, 0.5 - 2j), atol=1e-15)
    e0 = rand_states(rng, 4, None)
    e = be.string_expval("XYZ", psi, 1.0, out=e0.copy())
    np.testing.assert_allclose(e - e0, be.string_expval("XYZ", psi), atol=1e-14)


@pytest.mark.parametrize("be", BACKENDS)
def test_error_paths(be):
    # PS:275-278, PS:347-353, PO:343-346: wrong leading dimension -> invalid_argument -> ValueError
    with pytest.raises(ValueError):
        be.string_apply("XYZ", np.zeros(4, dtype=np.complex128))
    with pytest.raises(ValueError):
        be.string_apply("XYZ", np.zeros((4, 2), dtype=np.complex128))
    with pytest.raises(ValueError):
        be.string_expval("XYZ", np.zeros((16, 2), dtype=np.complex128))
    with pytest.raises(ValueError):
        be.op_apply(["XYZ", "III"], [1, 1], np.zeros((4, 2), dtype=np.complex128))
    with pytest.raises(ValueError):
        be.op_expval(["XYZ", "III"], [1, 1], np.zeros((4, 2), dtype=np.complex128))
    with pytest.raises(ValueError):
        orc.encode_strings(["XAZ"])  # PS:194 bad character


# ------------------------------------------------------------------ port vs the compiled reference
@needs_ref
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("par", [False, True])
def test_port_matches_reference_all_entry_points(dtype, par):
    rng = np.random.default_rng(18)
    tol = 1e-13 if dtype == np.complex128 else 2e-5
    for n, S, B, K in [(1, 3, 2, 2), (4, 20, 5, 3), (7, 40, 9, 4), (10, 64, 16, 5)]:
        strings = rand_strings(rng, n, S)
        psi = rand_states(rng, 2**n, B, dtype)
        h = rand_states(rng, S, None, dtype) * 2 - (1 + 1j)
        hk = (rand_states(rng, S, K, dtype) * 2 - (1 + 1j)).astype(dtype)
        data = rng.random((K, B)).astype(np.float64 if dtype == np.complex128 else np.float32)
        c = 0.3 - 1.7j
        s0 = strings[0]
        assert rel_err(PORT.string_apply(s0, psi, c), REF.string_apply(s0, psi, c, par=par)) < tol
        assert rel_err(PORT.string_apply(s0, psi[:, 0].copy(), c), REF.string_apply(s0, psi[:, 0].copy(), c, par=par)) < tol
        assert rel_err(PORT.string_expval(s0, psi, c), REF.string_expval(s0, psi, c, par=par)) < tol
        assert rel_err(PORT.op_apply(strings, h, psi), REF.op_apply(strings, h, psi, par=par)) < tol
        assert rel_err(PORT.op_apply(strings, h, psi[:, 0].copy()), REF.op_apply(strings, h, psi[:, 0].copy(), par=par)) < tol
        assert rel_err(PORT.op_expval(strings, h, psi), REF.op_expval(strings, h, psi, par=par)) < tol
        assert rel_err(PORT.sop_apply(strings, hk, psi), REF.sop_apply(strings, hk, psi, par=par)) < tol
        assert rel_err(PORT.sop_apply_weighted(strings, hk, psi, data), REF.sop_apply_weighted(strings, hk, psi, data, par=par)) < tol
        assert rel_err(PORT.sop_expval(strings, hk, psi), REF.sop_expval(strings, hk, psi, par=par)) < tol


@needs_ref
def test_port_bitexact_vs_reference_seq_config1():
    # BASELINE config 1: PauliOp.apply, 10 qubits, 64 random strings, batch 16, complex128.
    # the closed form summed in string order is bit-identical to the reference seq path.
    rng = np.random.default_rng(18)
    strings = rand_strings(rng, 10, 64)
    h = rand_states(rng, 64, None) * 2 - (1 + 1j)
    psi = rand_states(rng, 1024, 16)
    a = PORT.op_apply(strings, h, psi)
    b = REF.op_apply(strings, h, psi, par=False)
    assert rel_err(a, b) < 1e-15


@needs_ref
def test_reference_rejects_mixed_weight_dtype():
    # SPO:484 does not instantiate for data_dtype != T; the wrapper reports that instead of faking it
    psi = np.zeros((2, 1), dtype=np.complex64)
    with pytest.raises(NotImplementedError):
        REF.sop_apply_weighted(["X"], np.ones((1, 1), np.complex64), psi, np.ones((1, 1), np.float64))
    # the port (like the GPU library) accepts float32 or float64 weights for either state type
    PORT.sop_apply_weighted(["X"], np.ones((1, 1), np.complex64), psi, np.ones((1, 1), np.float64))
