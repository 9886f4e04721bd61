"""TEST INFRASTRUCTURE ONLY -- a stand-in for ``libfastpauli_b200.so`` so the *host logic* of the Python front-end
(1-D / 2-D dispatch, dtype coercion, zeroed outputs, shape checks, plan caching, coefficient orientation) can be
exercised by ``-m "not gpu"`` tests on a machine without a CUDA device.

``MockABI`` answers the same entry points as ``include/fastpauli_b200.h`` for HOST pointers, computing with the CPU
oracle (``oracle/``: the checker, never the product).  It is installed only by the ``mock_abi`` fixture of
``tests/test_reference_python_cases.py`` (and by ``tests/run_reference_pytests.py --mock``); nothing under
``fast-pauli_b200/`` knows it exists, and the product still fails with ``RuntimeError`` without a GPU.
"""
from __future__ import annotations

import ctypes as C
import gc

import numpy as np

from oracle import oracle as orc

_P2 = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
       "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0 + 0j, -1.0])}


def _kron(string: str) -> np.ndarray:
    m = np.ones((1, 1), dtype=complex)
    for ch in string:
        m = np.kron(m, _P2[ch])
    return m


def _v(x):
    return x.value if hasattr(x, "value") else x


def _arr(ptr, shape, dtype) -> np.ndarray:
    """numpy view of caller memory behind a ctypes pointer argument."""
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    addr = _v(ptr)
    if not isinstance(addr, int):
        addr = C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def _strings(codes: np.ndarray) -> list[str]:
    return ["".join("IXYZ"[c] for c in row) for row in codes]


class MockABI:
    """Every ``fp_*`` symbol the Python front-end calls; unknown ``fp_*`` names succeed as no-ops."""

    def __init__(self):
        self.err = b""
        self.plans: dict[int, tuple] = {}
        self.next_handle = 1
        self.be = orc.port()
        self.calls: list[str] = []

    def __getattr__(self, name):
        if name.startswith("fp_"):
            return lambda *a: 0
        raise AttributeError(name)

    def fp_last_error(self):
        return self.err

    def _fail(self, msg: str) -> int:
        self.err = msg.encode()
        return 1  # FP_INVALID -> ValueError

    @staticmethod
    def _cdt(d):
        return np.complex64 if _v(d) == 0 else np.complex128

    def fp_ctx_create(self, dev, href):
        href._obj.value = 0xC0FFEE
        return 0

    # ---- PauliString
    def _string_call(self, kind, dt, n, codes, c, out, inp, N, B, acc):
        self.calls.append(kind)
        n, N, B, dt = _v(n), _v(N), _v(B), self._cdt(dt)
        if N != (1 << n if n else 0):
            return self._fail("states shape must match the dimension of the operators")
        s = _strings(_arr(codes, (1, n), np.uint8))[0]
        cc = _arr(c, (1,), dt)[0]
        x = _arr(inp, (N, B), dt)
        if kind == "fp_string_apply":
            o, r = _arr(out, (N, B), dt), self.be.string_apply(s, x, cc)
        else:
            o, r = _arr(out, (B,), dt), self.be.string_expval(s, x, cc)
        o[...] = (o if _v(acc) else 0) + r
        return 0

    def fp_string_apply(self, ctx, *a):
        return self._string_call("fp_string_apply", *a)

    def fp_string_expval(self, ctx, *a):
        return self._string_call("fp_string_expval", *a)

    # ---- PauliOp
    def fp_op_create(self, ctx, dt, n, S, codes, coeffs, href):
        n, S, dt = _v(n), _v(S), self._cdt(dt)
        h, self.next_handle = self.next_handle, self.next_handle + 1
        self.plans[h] = (dt, n, _strings(_arr(codes, (S, n), np.uint8)), _arr(coeffs, (S,), dt).copy())
        href._obj.value = h
        self.calls.append("fp_op_create")
        return 0

    def fp_op_destroy(self, h):
        self.plans.pop(_v(h), None)
        return 0

    def _op_call(self, kind, h, out, inp, N, B, acc):
        self.calls.append(kind)
        dt, n, strings, co = self.plans[_v(h)]
        N, B = _v(N), _v(B)
        if N != (1 << n if n and strings else 0):
            return self._fail("[PauliOp] states shape must match the dimension of the operators")
        x = _arr(inp, (N, B), dt)
        if kind == "fp_op_apply":
            o, r = _arr(out, (N, B), dt), self.be.op_apply(strings, co, x)
        else:
            o, r = _arr(out, (B,), dt), self.be.op_expval(strings, co, x)
        o[...] = (o if _v(acc) else 0) + r
        return 0

    def fp_op_apply(self, ctx, *a):
        return self._op_call("fp_op_apply", *a)

    def fp_op_expval(self, ctx, *a):
        return self._op_call("fp_op_expval", *a)

    # ---- SummedPauliOp
    def fp_sop_create(self, ctx, dt, n, S, codes, K, coeffs, href):
        n, S, K, dt = _v(n), _v(S), _v(K), self._cdt(dt)
        h, self.next_handle = self.next_handle, self.next_handle + 1
        self.plans[h] = (dt, n, _strings(_arr(codes, (S, n), np.uint8)), _arr(coeffs, (S, K), dt).copy())
        href._obj.value = h
        self.calls.append("fp_sop_create")
        return 0

    def fp_sop_destroy(self, h):
        self.plans.pop(_v(h), None)
        return 0

    def _sop_states(self, h, inp, N, B):
        dt, n, strings, co = self.plans[_v(h)]
        N, B = _v(N), _v(B)
        if N != (1 << n):
            return None
        return dt, strings, co, _arr(inp, (N, B), dt), N, B

    def fp_sop_apply(self, ctx, h, out, inp, N, B, acc):
        self.calls.append("fp_sop_apply")
        st = self._sop_states(h, inp, N, B)
        if st is None:
            return self._fail("[SummedPauliOp] states shape must match the dimension of the operators")
        dt, strings, co, x, N, B = st
        o = _arr(out, (N, B), dt)
        o[...] = (o if _v(acc) else 0) + self.be.sop_apply(strings, co, x)
        return 0

    def fp_sop_apply_weighted(self, ctx, h, out, inp, data, is64, N, B, acc):
        self.calls.append("fp_sop_apply_weighted")
        st = self._sop_states(h, inp, N, B)
        if st is None:
            return self._fail("[SummedPauliOp] states shape must match the dimension of the operators")
        dt, strings, co, x, N, B = st
        d = _arr(data, (co.shape[1], B), np.float64 if _v(is64) else np.float32)
        o = _arr(out, (N, B), dt)
        real = np.float64 if dt == np.complex128 else np.float32
        o[...] = (o if _v(acc) else 0) + self.be.sop_apply_weighted(strings, co, x, d.astype(real))
        return 0

    def fp_sop_expval(self, ctx, h, out, inp, N, B, acc):
        self.calls.append("fp_sop_expval")
        st = self._sop_states(h, inp, N, B)
        if st is None:
            return self._fail("[SummedPauliOp] states shape must match the dimension of the operators")
        dt, strings, co, x, N, B = st
        o = _arr(out, (co.shape[1], B), dt)
        o[...] = (o if _v(acc) else 0) + self.be.sop_expval(strings, co, x)
        return 0

    def fp_sop_square(self, ctx, dt, n, S, codes, K, coeffs, S2, codes2, out):
        """Dense restatement for small registers: coefficient of P_c in A_k^2 is tr(P_c A_k^2) / dim."""
        self.calls.append("fp_sop_square")
        n, S, K, S2, dt = _v(n), _v(S), _v(K), _v(S2), self._cdt(dt)
        assert n <= 6, "the mock squares densely: small registers only"
        strings = _strings(_arr(codes, (S, n), np.uint8))
        co = _arr(coeffs, (S, K), dt)
        dense = np.stack([_kron(s) for s in strings])
        o = _arr(out, (S2, K), dt)
        sq_dense = [_kron(s) for s in _strings(_arr(codes2, (S2, n), np.uint8))]
        for k in range(K):
            A = np.tensordot(co[:, k], dense, axes=1)
            A2 = A @ A
            for c, Pc in enumerate(sq_dense):
                o[c, k] = np.trace(Pc @ A2) / (1 << n)
        return 0


class installed:
    """Context manager: swap the package's ``lib`` (and its cached default context) for a ``MockABI``."""

    def __init__(self, fp_module):
        self.fp = fp_module

    def __enter__(self) -> MockABI:
        self.saved = (self.fp.lib, self.fp._default_ctx)
        self.fp._default_ctx = None
        self.fp.lib = self.mock = MockABI()
        return self.mock

    def __exit__(self, *exc):
        # every object created under the mock holds mock handles: make sure they are finalised against the mock
        self.fp._default_ctx = None
        gc.collect()
        self.fp.lib, self.fp._default_ctx = self.saved
        return False
