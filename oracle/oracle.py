"""TEST INFRASTRUCTURE ONLY -- ctypes driver for the CPU oracles.

Two backends with the same raw C signatures (see oracle/pauli_oracle.c and
oracle/ref_wrapper.cpp):

* ``port()``       -> oracle/liboracle.so            (plain-C closed-form restatement, prefix ``orc_``)
* ``reference()``  -> oracle/_ref/libfastpauli_ref.so (the UNMODIFIED reference headers from
                      /root/reference compiled by oracle/Makefile, prefix ``ref_``); ``None`` when the
                      prebuilt library is absent.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(fast-pauli_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CODE = {"I": 0, "X": 1, "Y": 2, "Z": 3}


def encode_strings(strings: Sequence[str] | np.ndarray, n_qubits: int | None = None) -> tuple[np.ndarray, int]:
    """Pauli strings -> (S, n) uint8 code matrix (0:I 1:X 2:Y 3:Z); left-most char first (PS:172-197)."""
    if isinstance(strings, np.ndarray):
        codes = np.ascontiguousarray(strings, dtype=np.uint8)
        if codes.ndim == 1:
            codes = codes[None, :]
        return codes, codes.shape[1]
    strings = list(strings)
    n = len(strings[0]) if strings else (n_qubits or 0)
    codes = np.zeros((len(strings), n), dtype=np.uint8)
    for s, st in enumerate(strings):
        if len(st) != n:
            raise ValueError("All PauliStrings must have the same size")
        for q, ch in enumerate(st):
            if ch not in _CODE:
                raise ValueError(f"Invalid Pauli character {ch}")
            codes[s, q] = _CODE[ch]
    return codes, n


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/ (only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or (
        os.path.isdir("/root/reference/fast_pauli/cpp/include")
        and not os.path.exists(os.path.join(HERE, "_ref", "libfastpauli_ref.so"))
    ):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


class OracleError(ValueError):
    """Raised where the reference throws std::invalid_argument (-> ValueError in its bindings)."""


class Backend:
    """Uniform numpy front-end over one raw C library (prefix ``orc_`` or ``ref_``)."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = kind  # "port" | "reference"
        getattr(self.lib, prefix + "last_error").restype = C.c_char_p

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _sfx(dtype) -> tuple[str, type]:
        dtype = np.dtype(dtype)
        if dtype == np.complex128:
            return "c128", np.float64
        if dtype == np.complex64:
            return "c64", np.float32
        raise TypeError(f"unsupported dtype {dtype}")

    def _call(self, name: str, *args) -> None:
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc == 3:
            raise NotImplementedError(getattr(self.lib, self.prefix + "last_error")().decode())
        if rc != 0:
            raise OracleError(getattr(self.lib, self.prefix + "last_error")().decode())

    @staticmethod
    def _p(a: np.ndarray):
        return a.ctypes.data_as(C.c_void_p)

    @staticmethod
    def _coef(c, real) -> np.ndarray:
        c = complex(c)
        return np.array([c.real, c.imag], dtype=real)

    def use_all_threads(self, cap: int | None = None) -> int:
        """Let the OpenMP reference use every core this process may run on (torchrun exports OMP_NUM_THREADS=1);
        `cap` bounds the count where the reference's par path allocates per-thread copies of the batch (PO:427)."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        if cap:
            n = min(n, cap)
        if self.prefix == "ref_":
            self.lib.ref_set_threads(C.c_int(n))
        else:
            try:  # the port uses the same libgomp runtime
                C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(n))
            except OSError:
                pass
        return n

    def max_threads(self) -> int:
        if self.prefix == "ref_":
            self.lib.ref_max_threads.restype = C.c_int
            return int(self.lib.ref_max_threads())
        return os.cpu_count() or 1

    # -- PauliString ---------------------------------------------------------------------------
    def string_apply(self, string, states: np.ndarray, coeff=1.0, out: np.ndarray | None = None, par: bool = False):
        """PauliString.apply (1-D, PS:296-341) / apply_batch (2-D, PS:377-436); accumulates into ``out``."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings([string] if isinstance(string, str) else string)
        states = np.ascontiguousarray(states)
        out = np.zeros_like(states) if out is None else out
        c = self._coef(coeff, real)
        if states.ndim == 1:
            self._call("string_apply1d_" + sfx, C.c_int(n), self._p(codes), self._p(c), self._p(out), self._p(states),
                       C.c_size_t(states.shape[0]), C.c_int(par))
        else:
            self._call("string_apply_" + sfx, C.c_int(n), self._p(codes), self._p(c), self._p(out), self._p(states),
                       C.c_size_t(states.shape[0]), C.c_size_t(states.shape[1]), C.c_int(par))
        return out

    def string_expval(self, string, states: np.ndarray, coeff=1.0, out: np.ndarray | None = None, par: bool = False):
        """PauliString.expectation_value (PS:470-538); states (dim, B) -> (B,)."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings([string] if isinstance(string, str) else string)
        states = np.ascontiguousarray(states)
        s2 = states.reshape(states.shape[0], -1)
        out = np.zeros(s2.shape[1], dtype=states.dtype) if out is None else out
        c = self._coef(coeff, real)
        self._call("string_expval_" + sfx, C.c_int(n), self._p(codes), self._p(c), self._p(out), self._p(s2),
                   C.c_size_t(s2.shape[0]), C.c_size_t(s2.shape[1]), C.c_int(par))
        return out

    # -- PauliOp -------------------------------------------------------------------------------
    def op_apply(self, strings, coeffs, states: np.ndarray, out: np.ndarray | None = None, par: bool = False):
        """PauliOp.apply 1-D (PO:362-383) / 2-D (PO:399-468); accumulates into ``out``."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings(strings)
        h = np.ascontiguousarray(coeffs, dtype=states.dtype)
        states = np.ascontiguousarray(states)
        out = np.zeros_like(states) if out is None else out
        if states.ndim == 1:
            self._call("op_apply1d_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes), self._p(h),
                       self._p(out), self._p(states), C.c_size_t(states.shape[0]), C.c_int(par))
        else:
            self._call("op_apply_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes), self._p(h),
                       self._p(out), self._p(states), C.c_size_t(states.shape[0]), C.c_size_t(states.shape[1]),
                       C.c_int(par))
        return out

    def op_expval(self, strings, coeffs, states: np.ndarray, out: np.ndarray | None = None, par: bool = False):
        """PauliOp.expectation_value (PO:482-549)."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings(strings)
        h = np.ascontiguousarray(coeffs, dtype=states.dtype)
        states = np.ascontiguousarray(states)
        s2 = states.reshape(states.shape[0], -1)
        out = np.zeros(s2.shape[1], dtype=states.dtype) if out is None else out
        self._call("op_expval_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes), self._p(h),
                   self._p(out), self._p(s2), C.c_size_t(s2.shape[0]), C.c_size_t(s2.shape[1]), C.c_int(par))
        return out

    # -- SummedPauliOp (coeffs is (n_strings, n_operators), SPO:45) ---------------------------------
    def sop_apply(self, strings, coeffs, states: np.ndarray, out: np.ndarray | None = None, par: bool = False):
        """SummedPauliOp.apply (SPO:277-349)."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings(strings)
        h = np.ascontiguousarray(coeffs, dtype=states.dtype)
        states = np.ascontiguousarray(states)
        out = np.zeros_like(states) if out is None else out
        self._call("sop_apply_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes),
                   C.c_size_t(h.shape[1]), self._p(h), self._p(out), self._p(states), C.c_size_t(states.shape[0]),
                   C.c_size_t(states.shape[1]), C.c_int(par))
        return out

    def sop_apply_weighted(self, strings, coeffs, states: np.ndarray, data: np.ndarray,
                           out: np.ndarray | None = None, par: bool = False):
        """SummedPauliOp.apply_weighted (SPO:364-503); data (n_operators, n_states) float32/float64."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings(strings)
        h = np.ascontiguousarray(coeffs, dtype=states.dtype)
        states = np.ascontiguousarray(states)
        data = np.ascontiguousarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        out = np.zeros_like(states) if out is None else out
        self._call("sop_apply_weighted_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes),
                   C.c_size_t(h.shape[1]), self._p(h), self._p(out), self._p(states), self._p(data),
                   C.c_int(data.dtype == np.float64), C.c_size_t(states.shape[0]), C.c_size_t(states.shape[1]),
                   C.c_int(par))
        return out

    def sop_expval(self, strings, coeffs, states: np.ndarray, out: np.ndarray | None = None, par: bool = False):
        """SummedPauliOp.expectation_value (SPO:520-614); -> (n_operators, n_states)."""
        sfx, real = self._sfx(states.dtype)
        codes, n = encode_strings(strings)
        h = np.ascontiguousarray(coeffs, dtype=states.dtype)
        states = np.ascontiguousarray(states)
        out = np.zeros((h.shape[1], states.shape[1]), dtype=states.dtype) if out is None else out
        self._call("sop_expval_" + sfx, C.c_int(n), C.c_size_t(codes.shape[0]), self._p(codes),
                   C.c_size_t(h.shape[1]), self._p(h), self._p(out), self._p(states), C.c_size_t(states.shape[0]),
                   C.c_size_t(states.shape[1]), C.c_int(par))
        return out


_port: Backend | None = None
_ref: Backend | None | bool = False


def ref_sop_square(strings: Sequence[str], coeffs: np.ndarray) -> tuple[list[str], np.ndarray] | None:
    """SummedPauliOp::square() of the compiled reference (SPO:197-268): (output strings, coeffs (n_sq, K)); None when
    the reference library is not available (the plain-C port has no square)."""
    be = reference()
    if be is None or not hasattr(be.lib, "ref_sop_square_c128"):
        return None
    coeffs = np.ascontiguousarray(coeffs)
    sfx, real = Backend._sfx(coeffs.dtype)
    codes, n = encode_strings(strings)
    S, K = coeffs.shape
    max_w = max(sum(ch != "I" for ch in s) for s in strings)
    from math import comb

    cap = sum(comb(n, w) * 3**w for w in range(min(n, 2 * max_w) + 1))
    sq_codes = np.zeros((cap, n), dtype=np.uint8)
    out = np.zeros((cap, K), dtype=coeffs.dtype)
    n_sq = C.c_size_t(0)
    be._call("sop_square_" + sfx, C.c_int(n), C.c_size_t(S), Backend._p(codes), C.c_size_t(K), Backend._p(coeffs),
             C.c_size_t(cap), C.byref(n_sq), Backend._p(sq_codes), Backend._p(out))
    letters = np.array(list("IXYZ"))
    strs = ["".join(letters[row]) for row in sq_codes[: n_sq.value]]
    return strs, out[: n_sq.value]


def port() -> Backend:
    """The plain-C restatement (always available; compiled on demand)."""
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _port = Backend(path, "orc_", "port")
    return _port


def reference() -> Backend | None:
    """The unmodified reference compiled from /root/reference, or None when not prebuilt."""
    global _ref
    if _ref is False:
        path = os.path.join(HERE, "_ref", "libfastpauli_ref.so")
        if not os.path.exists(path) and os.path.isdir("/root/reference/fast_pauli/cpp/include"):
            build()
        _ref = Backend(path, "ref_", "reference") if os.path.exists(path) else None
    return _ref


def best() -> Backend:
    """The reference when available, else the port."""
    return reference() or port()


# ---------------------------------------------------------------------------------------------
# Independent numpy closed form (third implementation; small sizes only)
# ---------------------------------------------------------------------------------------------
def masks(string: str) -> tuple[int, int, int]:
    """(x, z, nY mod 4) of a string; bit n-1-q <-> string[q] (PS:52-54, 94-98)."""
    n = len(string)
    x = z = ny = 0
    for q, ch in enumerate(string):
        bit = 1 << (n - 1 - q)
        if ch in "XY":
            x |= bit
        if ch in "YZ":
            z |= bit
        ny += ch == "Y"
    return x, z, ny & 3


def np_sparse(string: str) -> tuple[np.ndarray, np.ndarray]:
    """(k, m) of get_sparse_repr (PS:49-118) from the closed form."""
    x, z, ny = masks(string)
    dim = 1 << len(string) if string else 0
    i = np.arange(dim, dtype=np.int64)
    par = np.zeros(dim, dtype=np.int64)
    zz = i & z
    while np.any(zz):
        par ^= zz & 1
        zz >>= 1
    m = np.array([1, -1j, -1, 1j])[ny] * (1 - 2 * par)
    return i ^ x, m.astype(np.complex128)


def np_op_apply(strings: Sequence[str], coeffs, states: np.ndarray) -> np.ndarray:
    out = np.zeros_like(states)
    for s, h in zip(strings, coeffs):
        k, m = np_sparse(s)
        mm = (h * m).astype(states.dtype)
        out += (mm[:, None] * states[k]) if states.ndim == 2 else mm * states[k]
    return out


# ---------------------------------------------------------------------------------------------
# Like-ordered expectation values: numpy's pairwise summation, in the INPUT precision.
# The reference accumulates expectation values sequentially (PS:534), so its own rounding error grows like the number of
# terms; a tree-ordered sum in the same precision is the fair comparison for the GPU's tree reductions.  Used by the
# tests before any appeal to a higher-precision arbiter.
# ---------------------------------------------------------------------------------------------
def np_string_expvals(strings: Sequence[str], states: np.ndarray) -> np.ndarray:
    """E(s, t) = <psi_t| P_s |psi_t> with unit coefficients, in states.dtype, pairwise-summed over the rows."""
    states = np.ascontiguousarray(states)
    s2 = states.reshape(states.shape[0], -1)
    out = np.zeros((len(strings), s2.shape[1]), dtype=states.dtype)
    for k, st in enumerate(strings):
        idx, m = np_sparse(st)
        prod = np.conj(s2) * (m.astype(states.dtype)[:, None] * s2[idx])
        out[k] = prod.sum(axis=0, dtype=states.dtype)
    return out


def np_expval_pairwise(kind: str, *args) -> np.ndarray:
    """kind in {"string_expval", "op_expval", "sop_expval"} with the same argument lists as the Backend methods."""
    if kind == "string_expval":
        string, states = args[0], args[1]
        c = complex(args[2]) if len(args) > 2 else 1.0
        st = string if isinstance(string, str) else string[0]
        return (np_string_expvals([st], states)[0] * np.asarray(c, dtype=states.dtype)).astype(states.dtype)
    strings, coeffs, states = args[0], np.asarray(args[1]), args[2]
    E = np_string_expvals(list(strings), states)
    if kind == "op_expval":
        return (coeffs.astype(states.dtype) @ E).astype(states.dtype)
    if kind == "sop_expval":
        return (coeffs.astype(states.dtype).T @ E).astype(states.dtype)
    raise ValueError(kind)
