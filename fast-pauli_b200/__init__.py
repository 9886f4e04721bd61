"""fast_pauli_b200 -- Python face of the B200-native fast-pauli hot path.

This module is the ctypes binding a fast-pauli maintainer would add over the C ABI in
``include/fastpauli_b200.h``.  It mirrors the *signatures and behaviour* of the
reference's nanobind classes for the hot path only:

=====================================  ==========================================================
reference (fast_pauli/cpp/src/include)  here
=====================================  ==========================================================
``PauliString.apply(states, coeff)``    ``__pauli_string_bindings.hpp:131-159``
``PauliString.expectation_value``       ``__pauli_string_bindings.hpp:182-211``
``PauliOp.apply / expectation_value``   ``__pauli_op_bindings.hpp:489-567``
``SummedPauliOp.apply / apply_weighted  ``__summed_pauli_op_bindings.hpp:156-274``
/ expectation_value``
=====================================  ==========================================================

All compute happens in hand-written sm_100a kernels inside ``lib/libfastpauli_b200.so``.  There is **no CPU
fallback**: importing works anywhere (so symbols can be checked without a GPU) but the first compute call
without a CUDA device raises ``RuntimeError``; a missing shared library raises ``ImportError`` at import time.

Array arguments may be numpy arrays (host; staged through the GPU inside the call, result returned as a new
numpy array) or device-resident: a :class:`DeviceArray` or any object exposing ``__cuda_array_interface__``
(e.g. a CUDA ``torch.Tensor``), in which case the result is a :class:`DeviceArray` on the same GPU.
complex64 inputs stay complex64 (the reference's C++ templates support it, its Python bindings do not);
every other dtype is converted to complex128 like nanobind does for the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Iterable, Sequence

import numpy as np

__all__ = ["Pauli", "PauliString", "PauliOp", "SummedPauliOp", "DeviceArray", "Context", "lib", "default_context", "helpers"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libfastpauli_b200.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C fast-pauli_b200` (there is no Python/CPU fallback)"
    )
lib = C.CDLL(_LIB_PATH)
lib.fp_last_error.restype = C.c_char_p

FP_C64, FP_C128 = 0, 1
_STATUS_EXC = {1: ValueError, 2: RuntimeError, 3: RuntimeError, 4: MemoryError, 5: NotImplementedError}
_CODE = {"I": 0, "X": 1, "Y": 2, "Z": 3}
_LETTER = "IXYZ"


def _check(rc: int) -> None:
    if rc != 0:
        raise _STATUS_EXC.get(rc, RuntimeError)(lib.fp_last_error().decode())


def _dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.complex64:
        return FP_C64
    if dt == np.complex128:
        return FP_C128
    raise TypeError(f"fast_pauli_b200 handles complex64 / complex128 states, not {dt}")


def _encode(strings: Iterable[str]) -> tuple[np.ndarray, int]:
    strings = [str(s) for s in strings]
    n = len(strings[0]) if strings else 0
    codes = np.zeros((len(strings), n), dtype=np.uint8)
    for s, st in enumerate(strings):
        if len(st) != n:
            raise ValueError("All PauliStrings must have the same size")  # PO:588
        for q, ch in enumerate(st):
            if ch not in _CODE:
                raise ValueError(f"Invalid Pauli character {ch}")  # PS:194
            codes[s, q] = _CODE[ch]
    return codes, n


# --------------------------------------------------------------------------------------------- context
class Context:
    """One GPU + stream + scratch (``fp_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._pool: dict[int, list[int]] = {}
        self._pooled = 0
        _check(lib.fp_ctx_create(C.c_int(device), C.byref(self._h)))
        self.device = device

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                self.trim()
            except Exception:
                pass
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.fp_ctx_destroy(h)

    def sync(self) -> None:
        _check(lib.fp_ctx_sync(self._h))

    # -- tiny exact-size free list for result arrays: cudaMalloc / cudaFree per call costs milliseconds and
    #    synchronises the device, which dominates small device-resident calls
    _POOL_LIMIT = 8 << 30

    def _take(self, nbytes: int) -> int:
        pool = self.__dict__.setdefault("_pool", {})
        lst = pool.get(nbytes)
        if lst:
            self._pooled -= nbytes
            return lst.pop()
        p = C.c_void_p()
        rc = lib.fp_device_malloc(self._h, C.c_size_t(nbytes), C.byref(p))
        if rc == 4 and pool:  # out of memory: drop the cache and retry once
            self.trim()
            rc = lib.fp_device_malloc(self._h, C.c_size_t(nbytes), C.byref(p))
        _check(rc)
        return int(p.value)

    def _give(self, ptr: int, nbytes: int) -> None:
        pool = self.__dict__.setdefault("_pool", {})
        if self.__dict__.get("_pooled", 0) + nbytes > self._POOL_LIMIT:
            lib.fp_device_free(self._h, C.c_void_p(ptr))
            return
        self._pooled = self.__dict__.get("_pooled", 0) + nbytes
        pool.setdefault(nbytes, []).append(ptr)

    def trim(self) -> None:
        """Return every cached device buffer to the driver."""
        for lst in self.__dict__.get("_pool", {}).values():
            for ptr in lst:
                lib.fp_device_free(self._h, C.c_void_p(ptr))
        self._pool = {}
        self._pooled = 0

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run on a caller-owned ``cudaStream_t`` handle (``0`` = the legacy default stream, which is what
        ``torch.cuda.current_stream().cuda_stream`` returns by default); ``None`` restores the context's own stream."""
        if cuda_stream is None:
            _check(lib.fp_ctx_set_stream(self._h, C.c_void_p(0), C.c_int(0)))
        else:
            _check(lib.fp_ctx_set_stream(self._h, C.c_void_p(int(cuda_stream)), C.c_int(1)))

    def set_async(self, flag: bool) -> None:
        _check(lib.fp_ctx_set_async(self._h, C.c_int(bool(flag))))

    def set_zero_copy(self, flag: bool) -> None:
        _check(lib.fp_ctx_set_zero_copy(self._h, C.c_int(bool(flag))))

    def set_tensor_core(self, flag: bool) -> None:
        _check(lib.fp_ctx_set_tensor_core(self._h, C.c_int(bool(flag))))

    def set_coset(self, mode: int = 1, log_twc: int = -1, log_nt: int = 0) -> None:
        """Coset-blocked kernels: mode 0 never / 1 heuristic / 2 whenever applicable; log_twc / log_nt force the tile
        shape (2^log_twc vectors per row segment, 2^log_nt threads per CTA)."""
        _check(lib.fp_ctx_set_coset(self._h, C.c_int(mode), C.c_int(log_twc), C.c_int(log_nt)))

    def bind_host_to_gpu_numa(self) -> int | None:
        """Pin this process (and therefore the pinned host buffers it allocates from now on) to the NUMA node of the
        context's GPU; returns the node, or None when the topology is unknown.  With one process per GPU this keeps
        host<->device copies off the inter-socket link."""
        buf = C.create_string_buffer(32)
        _check(lib.fp_ctx_pci_bus_id(self._h, buf, C.c_int(32)))
        bus = buf.value.decode().lower()
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node < 0:
                return None
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus: set[int] = set()
                for part in f.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & set(os.sched_getaffinity(0))
            if not allowed:
                return None
            os.sched_setaffinity(0, allowed)
            return node
        except (OSError, ValueError):
            return None

    def set_pipeline(self, enable: bool = True, min_bytes: int = 0, chunk_bytes: int = 0) -> None:
        """Chunked upload / kernel / download pipeline of PauliString.apply on host arrays (0 keeps a size)."""
        _check(lib.fp_ctx_set_pipeline(self._h, C.c_int(bool(enable)), C.c_size_t(min_bytes), C.c_size_t(chunk_bytes)))

    def set_coset_few(self, mode: int = 1, column_tiles_per_cta: int = 0) -> None:
        """Few-mask coset passes (K3e / K3f / K3i / K3j): 0 off, 1 automatic, 2 without the TMA-fed kernels, 3 without
        K3i and K3j, 4 without K3j, 5 automatic with K3j on single-string masks as well (default: K3i)."""
        _check(lib.fp_ctx_set_coset_few(self._h, C.c_int(mode), C.c_int(column_tiles_per_cta)))

    def coset_kernels_used(self, reset: bool = True) -> int:
        """Bit mask of the coset-family kernels launched since the last reset: 1 K3b, 2 K3e, 4 K3f, 8 K3g, 16 K3i, 32 K3j."""
        m = C.c_uint32()
        _check(lib.fp_ctx_coset_kernels_used(self._h, C.byref(m), C.c_int(bool(reset))))
        return int(m.value)

    def set_rcoset(self, mode: int = 1, log_nt: int = 0) -> None:
        """Register-resident coset kernels (x-mask rank <= 4): mode 0 never / 1 automatic / 2 whenever applicable."""
        _check(lib.fp_ctx_set_rcoset(self._h, C.c_int(mode), C.c_int(log_nt)))

    def set_l2_budget(self, nbytes: int) -> None:
        _check(lib.fp_ctx_set_l2_budget(self._h, C.c_size_t(nbytes)))

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(lib.fp_ctx_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def mem_info(self) -> tuple[int, int]:
        f, t = C.c_size_t(), C.c_size_t()
        _check(lib.fp_device_mem_info(self._h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    # -- arrays
    def empty(self, shape, dtype=np.complex128) -> "DeviceArray":
        return DeviceArray(self, shape, dtype)

    def zeros(self, shape, dtype=np.complex128) -> "DeviceArray":
        a = DeviceArray(self, shape, dtype)
        _check(lib.fp_memset(self._h, C.c_void_p(a.ptr), C.c_int(0), C.c_size_t(a.nbytes)))
        return a

    def to_device(self, host: np.ndarray) -> "DeviceArray":
        host = np.ascontiguousarray(host)
        a = DeviceArray(self, host.shape, host.dtype)
        _check(lib.fp_memcpy(self._h, C.c_void_p(a.ptr), host.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)))
        return a

    def uniform(self, shape, dtype=np.complex128, seed: int = 18, first: int = 0) -> "DeviceArray":
        """Counter-based U[0,1)+iU[0,1) amplitudes generated on the device (``fp_fill_uniform``)."""
        a = DeviceArray(self, shape, dtype)
        _check(lib.fp_fill_uniform(self._h, C.c_int(_dtype_code(dtype)), C.c_void_p(a.ptr),
                                   C.c_uint64(a.size), C.c_uint64(first), C.c_uint64(seed)))
        return a

    def pinned_empty(self, shape, dtype=np.complex128) -> np.ndarray:
        """numpy array backed by pinned (page-locked) host memory."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        p = C.c_void_p()
        _check(lib.fp_host_malloc(self._h, C.c_size_t(max(n * dtype.itemsize, 16)), C.byref(p)))
        buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        _PINNED[arr.ctypes.data] = (self, p.value)
        return arr

    def pinned_free(self, arr: np.ndarray) -> None:
        ent = _PINNED.pop(arr.ctypes.data, None)
        if ent:
            lib.fp_host_free(self._h, C.c_void_p(ent[1]))


_PINNED: dict[int, tuple] = {}
_default_ctx: Context | None = None
_default_lock = threading.Lock()


def default_context() -> Context:
    """Process-wide context on ``FASTPAULI_DEVICE`` (default: ``LOCAL_RANK`` or 0)."""
    global _default_ctx
    with _default_lock:
        if _default_ctx is None:
            dev = int(os.environ.get("FASTPAULI_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            _default_ctx = Context(dev)
        return _default_ctx


class DeviceArray:
    """A C-contiguous array in HBM owned by a :class:`Context` (or a zero-copy view of foreign device memory)."""

    def __init__(self, ctx: Context, shape, dtype=np.complex128, ptr: int | None = None, owner=None):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._owner = owner
        if ptr is None:
            self._cap = max(self.nbytes, 16)
            self.ptr = ctx._take(self._cap)
            self._owned = True
        else:
            self.ptr = int(ptr)
            self._owned = False

    def __del__(self):
        if getattr(self, "_owned", False) and self.ctx._h:
            self.ctx._give(self.ptr, self._cap)
            self._owned = False

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def __cuda_array_interface__(self) -> dict:
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3,
                "strides": None}

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        _check(lib.fp_memcpy(self.ctx._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr),
                             C.c_size_t(self.nbytes)))
        return out

    def get_rows(self, r0: int, r1: int) -> np.ndarray:
        """Rows [r0, r1) of a 2-D (or 1-D) array as a host copy."""
        row_elems = int(np.prod(self.shape[1:])) if self.ndim > 1 else 1
        out = np.empty((r1 - r0,) + self.shape[1:], dtype=self.dtype)
        _check(lib.fp_memcpy(self.ctx._h, out.ctypes.data_as(C.c_void_p),
                             C.c_void_p(self.ptr + r0 * row_elems * self.dtype.itemsize), C.c_size_t(out.nbytes)))
        return out

    def set(self, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.size
        _check(lib.fp_memcpy(self.ctx._h, C.c_void_p(self.ptr), host.ctypes.data_as(C.c_void_p),
                             C.c_size_t(self.nbytes)))


# --------------------------------------------------------------------------------------------- argument plumbing
class _Arg:
    """Normalised array argument: pointer + shape + dtype, host or device."""

    __slots__ = ("ptr", "shape", "dtype", "on_device", "keep", "ctx")

    def __init__(self, obj, ctx: Context, want_real: bool = False):
        self.ctx = ctx
        if isinstance(obj, DeviceArray):
            self.ptr, self.shape, self.dtype, self.on_device, self.keep = obj.ptr, obj.shape, obj.dtype, True, obj
            return
        cai = getattr(obj, "__cuda_array_interface__", None)
        if cai is not None and not isinstance(obj, np.ndarray):
            if cai.get("strides") not in (None,) and not _is_c_contiguous(cai):
                raise ValueError("array must be C-contiguous (row-major)")  # NB:55-74
            self.ptr, self.shape = int(cai["data"][0]), tuple(cai["shape"])
            self.dtype, self.on_device, self.keep = np.dtype(cai["typestr"]), True, obj
            # device arrays are used in place: there is no implicit conversion as for host arrays, so the element
            # type must already be what the kernels read (a float64 / int buffer read as complex128 is out of bounds)
            ok = (np.float32, np.float64) if want_real else (np.complex64, np.complex128)
            if self.dtype not in ok:
                raise TypeError(f"device array of dtype {self.dtype}: expected "
                                f"{' or '.join(np.dtype(t).name for t in ok)} (device arrays are not converted)")
            return
        arr = np.asarray(obj)
        if want_real:
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
        elif arr.dtype not in (np.complex64, np.complex128):
            arr = arr.astype(np.complex128)  # nanobind converts implicitly (PY_PS:185,191 pass float/int arrays)
        if not arr.flags.c_contiguous:
            raise ValueError("array must be C-contiguous (row-major)")  # NB:55-74
        self.ptr, self.shape, self.dtype, self.on_device, self.keep = arr.ctypes.data, arr.shape, arr.dtype, False, arr

    @property
    def ndim(self) -> int:
        return len(self.shape)


def _is_c_contiguous(cai: dict) -> bool:
    shape, strides = cai["shape"], cai["strides"]
    item = np.dtype(cai["typestr"]).itemsize
    expect = item
    for s, st in zip(reversed(shape), reversed(strides)):
        if s != 1 and st != expect:
            return False
        expect *= s
    return True


def _alloc_like(arg: _Arg, shape, dtype):
    """Zeroed output like the reference's owning_ndarray_from_shape (NB:149-172): host -> numpy, device -> DeviceArray."""
    if arg.on_device:
        return arg.ctx.empty(shape, dtype)  # the kernels overwrite (accumulate=0): no need to zero
    return np.empty(shape, dtype=dtype)


def _state_arg(op, states, what: str):
    """Normalise the states argument and run the reference's shape checks (PS:271-283, PO:343-350, SPO:290-293,
    B_PS:155, B_PO:512) BEFORE a device context is needed: a wrong shape is a ValueError on any machine."""
    a = _Arg(states, None)
    if a.ndim not in (1, 2):
        raise ValueError(f"{what}: expected 1 or 2 dimensions, got {a.ndim}")
    if a.shape[0] != op.dim:
        raise ValueError(f"[{type(op).__name__}] states shape ({a.shape[0]}) must match the dimension of the "
                         f"operators ({op.dim})")
    a.ctx = ctx = op._context()
    return ctx, a


def _ptr(o) -> C.c_void_p:
    return C.c_void_p(o.ptr if isinstance(o, DeviceArray) else o.ctypes.data)


def _coef_buf(c, dtype) -> np.ndarray:
    return np.array([complex(c)], dtype=dtype)


# --------------------------------------------------------------------------------------------- host-side algebra
# Everything below the hot path (products, sums, dense exports, generators) is small host code on the (x, z) bit
# encoding  I=(0,0) X=(1,0) Y=(1,1) Z=(0,1):  sigma(x,z) = i^(x.z) X^x Z^z.
def _masks(string: str) -> tuple[int, int]:
    n = len(string)
    x = z = 0
    for q, ch in enumerate(string):
        bit = 1 << (n - 1 - q)
        if ch in "XY":
            x |= bit
        if ch in "YZ":
            z |= bit
    return x, z


def _string_from_masks(x: int, z: int, n: int) -> str:
    return "".join("IXZY"[((x >> (n - 1 - q)) & 1) | (((z >> (n - 1 - q)) & 1) << 1)] for q in range(n))


def _product(a: str, b: str) -> tuple[complex, str]:
    """(phase, string) of the matrix product of two Pauli strings (reference: PauliString operator*, PS:228-247)."""
    if len(a) != len(b):
        raise ValueError("PauliStrings must have the same size")
    xa, za = _masks(a)
    xb, zb = _masks(b)
    x, z = xa ^ xb, za ^ zb
    k = (bin(xa & za).count("1") + bin(xb & zb).count("1") + 2 * bin(za & xb).count("1") - bin(x & z).count("1")) & 3
    return (1, 1j, -1, -1j)[k], _string_from_masks(x, z, len(a))


def _dense(string: str) -> np.ndarray:
    """Dense matrix of a Pauli string from the closed form k[i] = i ^ x, m[i] = (-i)^nY (-1)^popc(i & z)."""
    n = len(string)
    dim = 1 << n if n else 0
    x, z = _masks(string)
    i = np.arange(dim, dtype=np.int64)
    par = np.zeros(dim, dtype=np.int64)
    zz = i & z
    while np.any(zz):
        par ^= zz & 1
        zz >>= 1
    m = np.array([1, -1j, -1, 1j])[string.count("Y") & 3] * (1 - 2 * par)
    out = np.zeros((dim, dim), dtype=np.complex128)
    out[i, i ^ x] = m
    return out


class Pauli:
    """One 2x2 Pauli matrix (reference binding: __pauli_bindings.hpp:36-106)."""

    __slots__ = ("code",)

    def __init__(self, code: "int | str | Pauli" = 0, *, symbol: "str | None" = None):
        # the reference binds three overloads: Pauli(), Pauli(code: int), Pauli(symbol: str) (B_P:41-58)
        if symbol is not None:
            code = symbol
        if isinstance(code, Pauli):
            code = code.code
        if isinstance(code, str):
            if len(code) != 1:
                raise TypeError("Pauli(symbol): expected a single character")  # no overload matches (PY_P:68-69)
            if code not in _CODE:
                raise ValueError("Invalid Pauli matrix symbol")
            code = _CODE[code]
        if not 0 <= int(code) <= 3:
            raise ValueError("Pauli code must be 0, 1, 2, or 3")
        self.code = int(code)

    def __matmul__(self, rhs: "Pauli") -> tuple[complex, "Pauli"]:
        phase, s = _product(_LETTER[self.code], _LETTER[rhs.code])
        return phase, Pauli(s)

    def to_tensor(self) -> np.ndarray:
        return _dense(_LETTER[self.code])

    def clone(self) -> "Pauli":
        return Pauli(self.code)

    def __str__(self) -> str:
        return _LETTER[self.code]

    def __repr__(self) -> str:
        return f"Pauli('{self}')"

    def __eq__(self, other) -> bool:
        return isinstance(other, Pauli) and other.code == self.code

    def __hash__(self) -> int:
        return hash(self.code)

    def __getstate__(self):
        return self.code

    def __setstate__(self, code):
        self.code = int(code)


# --------------------------------------------------------------------------------------------- PauliString
class PauliString:
    """Tensor product of Pauli matrices (reference: ``struct PauliString``, __pauli_string.hpp:126-206)."""

    def __init__(self, string: "str | PauliString | Sequence[Pauli]" = "", ctx: Context | None = None):
        if not isinstance(string, (str, PauliString)):
            string = "".join(str(Pauli(p)) for p in string)  # list of Pauli objects (B_PS: init from paulis)
        string = str(string)
        self._codes, self._n = _encode([string])
        self.string = string
        self._ctx = ctx

    # -- properties mirrored from the bindings (__pauli_string_bindings.hpp:120-129)
    @property
    def n_qubits(self) -> int:
        return self._n

    @property
    def dim(self) -> int:
        return (1 << self._n) if self._n else 0

    @property
    def weight(self) -> int:
        return int(np.count_nonzero(self._codes))

    def __str__(self) -> str:
        return self.string

    def __repr__(self) -> str:
        return f'PauliString("{self.string}")'

    def __eq__(self, other) -> bool:
        return isinstance(other, PauliString) and other.string == self.string

    def __hash__(self) -> int:
        return hash(self.string)

    def clone(self) -> "PauliString":
        return PauliString(self.string, self._ctx)

    # -- host-side algebra mirrored from the bindings (__pauli_string_bindings.hpp:61-121, 235-260)
    def __matmul__(self, rhs: "PauliString") -> tuple[complex, "PauliString"]:
        if not isinstance(rhs, PauliString):
            return NotImplemented  # PauliString @ PauliOp is PauliOp.__rmatmul__
        phase, prod = _product(self.string, rhs.string)
        return phase, PauliString(prod, self._ctx)

    def __add__(self, other: "PauliString") -> "PauliOp":
        if not isinstance(other, PauliString):
            return NotImplemented
        return PauliOp([1, 1], [self, other], self._ctx)

    def __sub__(self, other: "PauliString") -> "PauliOp":
        if not isinstance(other, PauliString):
            return NotImplemented
        return PauliOp([1, -1], [self, other], self._ctx)

    def to_tensor(self) -> np.ndarray:
        return _dense(self.string)

    def __getstate__(self):
        return self.string

    def __setstate__(self, string):
        self.__init__(string)

    def _context(self) -> Context:
        return self._ctx or default_context()

    def apply(self, states, coeff: complex = 1.0):
        """``coeff * P |psi_t>`` for a 1-D state or a (dim, n_states) batch (B_PS:131-159).

        Mirrors a quirk of the reference binding: for a 1-D state the ``coeff`` argument is NOT applied
        (``__pauli_string_bindings.hpp:142`` calls ``apply`` without ``c``); 2-D honours it.
        """
        ctx, a = _state_arg(self, states, "apply")
        if a.ndim == 1:
            coeff = 1.0
        B = 1 if a.ndim == 1 else a.shape[1]
        out = _alloc_like(a, a.shape, a.dtype)
        c = _coef_buf(coeff, a.dtype)
        _check(lib.fp_string_apply(ctx._h, C.c_int(_dtype_code(a.dtype)), C.c_int(self._n),
                                   self._codes.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), _ptr(out),
                                   C.c_void_p(a.ptr), C.c_size_t(a.shape[0]), C.c_size_t(B), C.c_int(0)))
        return out

    def expectation_value(self, states, coeff: complex = 1.0):
        """``<psi_t| coeff P |psi_t>``: 1-D state -> shape (1,), batch -> (n_states,) (B_PS:182-211)."""
        ctx, a = _state_arg(self, states, "expectation_value")
        B = 1 if a.ndim == 1 else a.shape[1]
        out = _alloc_like(a, (B,), a.dtype)
        c = _coef_buf(coeff, a.dtype)
        _check(lib.fp_string_expval(ctx._h, C.c_int(_dtype_code(a.dtype)), C.c_int(self._n),
                                    self._codes.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), _ptr(out),
                                    C.c_void_p(a.ptr), C.c_size_t(a.shape[0]), C.c_size_t(B), C.c_int(0)))
        return out


# --------------------------------------------------------------------------------------------- PauliOp
class PauliOp:
    """Weighted sum of Pauli strings ``sum_i h_i P_i`` (reference: ``struct PauliOp``, __pauli_op.hpp:38-97).

    The packed, x-mask-grouped device plan (``fp_op``) is built lazily per dtype and reused across calls.
    """

    def __init__(self, coeffs: Sequence[complex] | np.ndarray | None = None,
                 strings: Sequence["str | PauliString"] | None = None, ctx: Context | None = None):
        if strings is None and coeffs is not None and len(coeffs) and isinstance(coeffs[0], (str, PauliString)):
            strings, coeffs = coeffs, None  # PauliOp(strings): coefficients default to one (PO:59-80)
        strings = [str(s) for s in ([] if strings is None else strings)]
        self._coeffs = np.ones(len(strings), np.complex128) if coeffs is None else \
            np.array(coeffs, dtype=np.complex128).reshape(-1)
        if len(self._coeffs) != len(strings):
            raise ValueError("coeffs and pauli_strings must have the same size")  # PO:92-95
        self._codes, self._n = _encode(strings)
        self._strings = strings
        self._ctx = ctx
        self._plans: dict[int, C.c_void_p] = {}

    def __del__(self):
        self._drop_plans()

    def _drop_plans(self) -> None:
        for h in getattr(self, "_plans", {}).values():
            lib.fp_op_destroy(h)
        self._plans = {}

    # -- properties (__pauli_op_bindings.hpp)
    @property
    def dim(self) -> int:
        return (1 << self._n) if (self._strings and self._n) else 0

    @property
    def n_qubits(self) -> int:
        return self._n if self._strings else 0

    @property
    def n_pauli_strings(self) -> int:
        return len(self._strings)

    @property
    def coeffs(self) -> np.ndarray:
        return self._coeffs.copy()

    @property
    def pauli_strings(self) -> list[PauliString]:
        return [PauliString(s) for s in self._strings]

    @property
    def pauli_strings_as_str(self) -> list[str]:
        return list(self._strings)

    def scale(self, factors) -> None:
        """Scale each term (PO:142-158); drops the cached device plans."""
        f = np.asarray(factors, dtype=np.complex128)
        if f.ndim and f.shape != self._coeffs.shape:
            raise ValueError("factors must have the same length as the number of PauliStrings")
        self._coeffs = self._coeffs * f
        self._drop_plans()

    def clone(self) -> "PauliOp":
        return PauliOp(self._coeffs.copy(), list(self._strings), self._ctx)

    # -- host-side operator algebra mirrored from the bindings (__pauli_op_bindings.hpp:113-440, 589-617)
    def _check_dim(self, other) -> None:
        if other.dim != self.dim:
            raise ValueError("Mismatched dimensions for provided PauliOp / PauliString")

    def __matmul__(self, rhs: "PauliOp | PauliString") -> "PauliOp":
        self._check_dim(rhs)
        if isinstance(rhs, PauliString):
            pairs = [_product(s, rhs.string) for s in self._strings]
            return PauliOp([c * ph for c, (ph, _) in zip(self._coeffs, pairs)], [p for _, p in pairs], self._ctx)
        merged: dict[str, complex] = {}
        for ca, sa in zip(self._coeffs, self._strings):  # identical product strings are merged (PO:237-281)
            for cb, sb in zip(rhs._coeffs, rhs._strings):
                ph, prod = _product(sa, sb)
                merged[prod] = merged.get(prod, 0) + ph * ca * cb
        return PauliOp(list(merged.values()), list(merged.keys()), self._ctx)

    def __rmatmul__(self, lhs: PauliString) -> "PauliOp":
        self._check_dim(lhs)
        pairs = [_product(lhs.string, s) for s in self._strings]
        return PauliOp([c * ph for c, (ph, _) in zip(self._coeffs, pairs)], [p for _, p in pairs], self._ctx)

    def __mul__(self, factor: complex) -> "PauliOp":
        return PauliOp(self._coeffs * complex(factor), list(self._strings), self._ctx)

    __rmul__ = __mul__

    def __imul__(self, factor: complex) -> "PauliOp":
        self.scale(complex(factor))
        return self

    def extend(self, other: "PauliOp | PauliString", multiplier: complex = 1.0, dedupe: bool = True) -> None:
        """Append terms (PO:291-338): a PauliOp is appended as is, a PauliString is merged when ``dedupe``."""
        self._check_dim(other) if self._strings else None
        if isinstance(other, PauliString):
            if dedupe and other.string in self._strings:
                self._coeffs[self._strings.index(other.string)] += complex(multiplier)
            else:
                self._strings.append(other.string)
                self._coeffs = np.append(self._coeffs, complex(multiplier))
        else:
            self._strings.extend(other._strings)
            self._coeffs = np.append(self._coeffs, other._coeffs * complex(multiplier))
        self._codes, self._n = _encode(self._strings)
        self._drop_plans()

    def _plus(self, other, sign: float) -> "PauliOp":
        out = self.clone()
        out.extend(other, sign, dedupe=True) if isinstance(other, PauliString) else out.extend(other, sign)
        return out

    def __add__(self, other):
        return self._plus(other, 1.0)

    def __radd__(self, other: PauliString):
        return self._plus(other, 1.0)

    def __iadd__(self, other):
        self.extend(other, 1.0)
        return self

    def __sub__(self, other):
        return self._plus(other, -1.0)

    def __rsub__(self, other: PauliString):
        out = self * -1.0
        out.extend(other, 1.0)
        return out

    def __isub__(self, other):
        self.extend(other, -1.0)
        return self

    def to_tensor(self) -> np.ndarray:
        out = np.zeros((self.dim, self.dim), dtype=np.complex128)
        for c, st in zip(self._coeffs, self._strings):
            out += c * _dense(st)
        return out

    def __getstate__(self):
        return (self._coeffs.copy(), list(self._strings))

    def __setstate__(self, state):
        self.__init__(state[0], state[1])

    def _context(self) -> Context:
        return self._ctx or default_context()

    def _plan(self, dtype) -> C.c_void_p:
        code = _dtype_code(dtype)
        if code not in self._plans:
            h = C.c_void_p()
            coeffs = np.ascontiguousarray(self._coeffs, dtype=dtype)
            _check(lib.fp_op_create(self._context()._h, C.c_int(code), C.c_int(self._n),
                                    C.c_size_t(len(self._strings)), self._codes.ctypes.data_as(C.c_void_p),
                                    coeffs.ctypes.data_as(C.c_void_p), C.byref(h)))
            self._plans[code] = h
        return self._plans[code]

    def plan_info(self, dtype=np.complex128) -> dict:
        d, n, s, p, g = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(lib.fp_op_info(self._plan(dtype), C.byref(d), C.byref(n), C.byref(s), C.byref(p), C.byref(g)))
        return {"n_qubits": n.value, "n_strings": s.value, "n_packed_strings": p.value, "n_x_groups": g.value}

    def apply(self, states):
        """``(sum_i h_i P_i) |psi_t>`` for a 1-D state or (dim, n_states) batch (B_PO:489-516)."""
        ctx, a = _state_arg(self, states, "apply")
        B = 1 if a.ndim == 1 else a.shape[1]
        out = _alloc_like(a, a.shape, a.dtype)
        _check(lib.fp_op_apply(ctx._h, self._plan(a.dtype), _ptr(out), C.c_void_p(a.ptr), C.c_size_t(a.shape[0]),
                               C.c_size_t(B), C.c_int(0)))
        return out

    def expectation_value(self, states):
        """``<psi_t| sum_i h_i P_i |psi_t>`` (B_PO:537-567)."""
        ctx, a = _state_arg(self, states, "expectation_value")
        B = 1 if a.ndim == 1 else a.shape[1]
        out = _alloc_like(a, (B,), a.dtype)
        _check(lib.fp_op_expval(ctx._h, self._plan(a.dtype), _ptr(out), C.c_void_p(a.ptr), C.c_size_t(a.shape[0]),
                                C.c_size_t(B), C.c_int(0)))
        return out


# --------------------------------------------------------------------------------------------- SummedPauliOp
class SummedPauliOp:
    """``A_k = sum_i h_ik P_i`` for k = 0..n_operators-1 (reference: __summed_pauli_op.hpp:37-145).

    ``coeffs`` is (n_pauli_strings, n_operators) like the reference constructor (B_SPO:60-84).
    """

    def __init__(self, strings: Sequence["str | PauliString"], coeffs: np.ndarray, ctx: Context | None = None):
        strings = [str(s) for s in strings]
        coeffs = np.array(coeffs, dtype=np.complex128)
        if coeffs.ndim == 1 and len(strings):
            coeffs = coeffs.reshape(len(strings), -1)  # flat (n_strings * n_operators) form, SPO:83-92
        if coeffs.ndim != 2 or coeffs.shape[0] != len(strings):
            raise ValueError("The number of PauliStrings must match the number of rows in the coeffs matrix")  # SPO:60-64
        if not strings:
            raise ValueError("SummedPauliOp needs at least one PauliString")
        self._codes, self._n = _encode(strings)
        self._strings = strings
        self._coeffs = np.ascontiguousarray(coeffs)
        self._ctx = ctx
        self._plans: dict[int, C.c_void_p] = {}

    def __del__(self):
        self._drop_plans()

    @property
    def dim(self) -> int:
        return 1 << self._n if self._n else 0

    @property
    def n_qubits(self) -> int:
        return self._n

    @property
    def n_operators(self) -> int:
        return self._coeffs.shape[1]

    @property
    def n_pauli_strings(self) -> int:
        return len(self._strings)

    @property
    def coeffs(self) -> np.ndarray:
        """(n_operators, n_pauli_strings): the reference getter returns the transpose (B_SPO:113-135)."""
        return self._coeffs.T.copy()

    @coeffs.setter
    def coeffs(self, coeffs_new) -> None:
        """Takes the getter's (n_operators, n_pauli_strings) orientation (B_SPO:125-135); drops the device plans."""
        c = np.asarray(coeffs_new, dtype=np.complex128)
        if c.ndim != 2 or c.shape != (self.n_operators, self.n_pauli_strings):
            raise ValueError("The shape of provided coeffs must match the number of operators and PauliStrings")
        self._coeffs = np.ascontiguousarray(c.T)
        self._drop_plans()

    def _drop_plans(self) -> None:
        for h in getattr(self, "_plans", {}).values():
            lib.fp_sop_destroy(h)
        self._plans = {}

    @property
    def pauli_strings(self) -> list[PauliString]:
        return [PauliString(s) for s in self._strings]

    def clone(self) -> "SummedPauliOp":
        return SummedPauliOp(list(self._strings), self._coeffs.copy(), self._ctx)

    @property
    def pauli_strings_as_str(self) -> list[str]:
        return list(self._strings)

    # -- host-side helpers mirrored from the bindings (__summed_pauli_op_bindings.hpp:291-340)
    def split(self) -> list[PauliOp]:
        """The K operators as separate PauliOps (SPO:621-637)."""
        return [PauliOp(self._coeffs[:, k].copy(), list(self._strings), self._ctx) for k in range(self.n_operators)]

    def to_tensor(self) -> np.ndarray:
        """Dense (n_operators, dim, dim) tensor (SPO:644-665)."""
        dense = np.stack([_dense(st) for st in self._strings])  # (S, dim, dim)
        return np.einsum("sk,sij->kij", self._coeffs, dense)

    def square(self) -> "SummedPauliOp":
        """A_k -> A_k^2 (SPO:197-268): coefficient of string c in operator k is sum over pairs (a, b) with
        P_a P_b ~ P_c of phase(a,b) h_ak h_bk; the output string set is every string up to weight
        min(n, 2 * max weight) in calculate_pauli_strings_max_weight order.  The contraction runs on the GPU
        (fp_sop_square)."""
        from . import helpers

        max_w = max(sum(ch != "I" for ch in st) for st in self._strings)
        sq = [str(p) for p in helpers.calculate_pauli_strings_max_weight(self._n, min(self._n, 2 * max_w))]
        sq_codes, _ = _encode(sq)
        ctx = self._ctx or default_context()
        coeffs = np.ascontiguousarray(self._coeffs, dtype=np.complex128)
        out = np.zeros((len(sq), self.n_operators), dtype=np.complex128)
        _check(lib.fp_sop_square(ctx._h, C.c_int(_dtype_code(np.complex128)), C.c_int(self._n),
                                 C.c_size_t(len(self._strings)), self._codes.ctypes.data_as(C.c_void_p),
                                 C.c_size_t(self.n_operators), coeffs.ctypes.data_as(C.c_void_p), C.c_size_t(len(sq)),
                                 sq_codes.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return SummedPauliOp(sq, out, self._ctx)

    def __getstate__(self):
        return (list(self._strings), self._coeffs.copy())

    def __setstate__(self, state):
        self.__init__(state[0], state[1])

    def _context(self) -> Context:
        return self._ctx or default_context()

    def _plan(self, dtype) -> C.c_void_p:
        code = _dtype_code(dtype)
        if code not in self._plans:
            h = C.c_void_p()
            coeffs = np.ascontiguousarray(self._coeffs, dtype=dtype)
            _check(lib.fp_sop_create(self._context()._h, C.c_int(code), C.c_int(self._n),
                                     C.c_size_t(len(self._strings)), self._codes.ctypes.data_as(C.c_void_p),
                                     C.c_size_t(coeffs.shape[1]), coeffs.ctypes.data_as(C.c_void_p), C.byref(h)))
            self._plans[code] = h
        return self._plans[code]

    def apply(self, states):
        """``sum_k A_k |psi_t>`` (B_SPO:156-182)."""
        ctx, a = _state_arg(self, states, "apply")
        B = 1 if a.ndim == 1 else a.shape[1]
        out = _alloc_like(a, a.shape, a.dtype)
        _check(lib.fp_sop_apply(ctx._h, self._plan(a.dtype), _ptr(out), C.c_void_p(a.ptr), C.c_size_t(a.shape[0]),
                                C.c_size_t(B), C.c_int(0)))
        return out

    def apply_weighted(self, states, data):
        """``sum_k x_kt A_k |psi_t>`` with real weights ``data`` of shape (n_operators, n_states) (B_SPO:199-228)."""
        ctx, a = _state_arg(self, states, "apply_weighted")
        d = _Arg(data, ctx, want_real=True)
        B = 1 if a.ndim == 1 else a.shape[1]
        dshape = d.shape if d.ndim == 2 else (d.shape[0], 1)
        if d.ndim not in (1, 2) or d.ndim != a.ndim or dshape != (self.n_operators, B):  # SPO:389-394
            raise ValueError("data(k,t) must have the same number of operators as the SummedPauliOp "
                             "and the same number of states as the input states")
        if d.dtype not in (np.float32, np.float64):
            raise ValueError("data must be float32 or float64")
        out = _alloc_like(a, a.shape, a.dtype)
        _check(lib.fp_sop_apply_weighted(ctx._h, self._plan(a.dtype), _ptr(out), C.c_void_p(a.ptr),
                                         C.c_void_p(d.ptr), C.c_int(d.dtype == np.float64),
                                         C.c_size_t(a.shape[0]), C.c_size_t(B), C.c_int(0)))
        return out

    def expectation_value(self, states):
        """``<psi_t| A_k |psi_t>``: 1-D state -> (n_operators,), batch -> (n_operators, n_states) (B_SPO:246-274)."""
        ctx, a = _state_arg(self, states, "expectation_value")
        B = 1 if a.ndim == 1 else a.shape[1]
        shape = (self.n_operators,) if a.ndim == 1 else (self.n_operators, B)
        out = _alloc_like(a, shape, a.dtype)
        _check(lib.fp_sop_expval(ctx._h, self._plan(a.dtype), _ptr(out), C.c_void_p(a.ptr), C.c_size_t(a.shape[0]),
                                 C.c_size_t(B), C.c_int(0)))
        return out


from . import helpers  # noqa: E402  (the reference exposes these as the `helpers` submodule, fast_pauli.cpp:44-104)
