"""Multi-GPU helpers of the hot path: one process per GPU.  Neither sharding exists in the reference.

The production path of the high-qubit sharding is the C ABI (``fp_comm_*`` / ``fp_sharded_op_*``, csrc/sharded.cpp,
front-end :mod:`fast_pauli_b200.sharded`): NCCL inside the library, no torch on the data path.  This module keeps
the host-side planning in Python form (``plan_high_qubit`` / ``ShardedStateOp`` with injected arithmetic and
exchange, exercised on CPU by the world-size 2 / 4 gloo tests), the batch-axis helpers, and the fused CUDA-IPC
peer-memory variant (``PeerShards`` + ``ShardedStateOp.apply_peer``).

1. **Batch-axis sharding** (BASELINE configs 2-4): every hot-path formula is independent per batch column, so rank
   ``r`` of ``g`` owns the column block ``shard_columns(B, g, r)`` of the ``(dim, B)`` batch and runs the ordinary
   single-GPU calls on it.  No data-path collective; expectation values are concatenated with one ``all_gather``.

2. **High-qubit sharding** (BASELINE config 5: one state too large for a GPU): rank ``r`` owns the rows whose top
   ``log2 g`` index bits equal ``r``.  For a string with x-mask ``x = (x_hi << n_local) | x_lo``

       out_r[i_lo] += h * (-i)^nY * (-1)^popc(r & z_hi) * (-1)^popc(i_lo & z_lo) * psi_{r ^ x_hi}[i_lo ^ x_lo]

   so strings are grouped by ``x_hi``: the ``x_hi = 0`` class is purely local, every other class needs ONE pairwise
   shard swap with peer ``r ^ x_hi`` (``isend``/``irecv``; NVSwitch makes every XOR partner one hop), after which
   the class is an ordinary local PauliOp on ``n_local`` qubits whose coefficients carry the rank-dependent sign.
   The swap for class ``k + 1`` is in flight while class ``k`` is being applied (two receive buffers; 180 GB of HBM
   hold state + output + both buffers for a 34-qubit complex128 state on 8 GPUs: 4 x 32 GiB per GPU).

The local arithmetic is injected (``local_ops``) so the host logic can be exercised on CPU with the gloo backend
(tests/test_distributed.py); on GPUs the default is the CUDA library.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

__all__ = ["shard_columns", "HighQubitPlan", "plan_high_qubit", "ShardedStateOp", "gather_columns", "PeerShards"]


# ------------------------------------------------------------------------------------------------ batch axis
def shard_columns(n_states: int, world: int, rank: int) -> tuple[int, int]:
    """Column block [start, stop) of rank ``rank``: contiguous, sizes differ by at most one, covers [0, n_states)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_states, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_columns(local: "np.ndarray", n_states: int, dist=None) -> "np.ndarray":
    """Concatenate per-rank results along the LAST axis (expectation values (B_local,) or (K, B_local))."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_columns(n_states, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    lead = local.shape[:-1]
    pad = np.zeros(lead + (width,), dtype=local.dtype)
    pad[..., : local.shape[-1]] = local
    t = torch.from_numpy(np.ascontiguousarray(pad).view(np.float64 if local.dtype == np.complex128 else np.float32))
    backend = dist.get_backend()
    if backend == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = []
    for r, (a, b) in enumerate(sizes):
        arr = outs[r].cpu().numpy().view(local.dtype).reshape(lead + (width,))
        parts.append(arr[..., : b - a])
    return np.concatenate(parts, axis=-1)


# ------------------------------------------------------------------------------------------------ high qubits
@dataclass
class HighQubitClass:
    x_hi: int                      # peer offset: this class reads the shard of rank r ^ x_hi
    strings: list[str]             # low n_local characters of every string in the class
    coeffs: np.ndarray             # h * (-i)^nY_hi  (rank independent part)
    z_hi: np.ndarray               # per string: high z bits, sign (-1)^popc(rank & z_hi) is applied per rank


@dataclass
class HighQubitPlan:
    n_qubits: int
    n_local: int
    world: int
    classes: list[HighQubitClass] = field(default_factory=list)

    def peer_offsets(self) -> list[int]:
        return [c.x_hi for c in self.classes if c.x_hi]


def plan_high_qubit(strings: Sequence[str], coeffs: Sequence[complex], world: int) -> HighQubitPlan:
    """Split every string into (high, low) characters and group by the high x-mask (host-side, rank independent)."""
    if world & (world - 1) or world <= 0:
        raise ValueError("world size must be a power of two")
    n = len(strings[0])
    n_hi = int(math.log2(world))
    if n_hi > n:
        raise ValueError("more ranks than rows")
    n_local = n - n_hi
    by_class: dict[int, HighQubitClass] = {}
    phases = [1, -1j, -1, 1j]
    for s, h in zip(strings, coeffs):
        if len(s) != n:
            raise ValueError("All PauliStrings must have the same size")
        hi, lo = s[:n_hi], s[n_hi:]
        x_hi = z_hi = 0
        for q, ch in enumerate(hi):  # left-most character is the most significant qubit (PS:52-54)
            bit = 1 << (n_hi - 1 - q)
            if ch in "XY":
                x_hi |= bit
            if ch in "YZ":
                z_hi |= bit
        c = complex(h) * phases[hi.count("Y") & 3]
        cls = by_class.setdefault(x_hi, HighQubitClass(x_hi, [], np.zeros(0, np.complex128), np.zeros(0, np.int64)))
        cls.strings.append(lo)
        cls.coeffs = np.append(cls.coeffs, c)
        cls.z_hi = np.append(cls.z_hi, z_hi)
    plan = HighQubitPlan(n, n_local, world)
    plan.classes = [by_class[k] for k in sorted(by_class)]  # x_hi = 0 (local) first, then the swaps
    return plan


def _rank_coeffs(cls: HighQubitClass, rank: int) -> np.ndarray:
    sign = np.array([-1.0 if bin(rank & int(z)).count("1") & 1 else 1.0 for z in cls.z_hi])
    return cls.coeffs * sign


class ShardedStateOp:
    """PauliOp.apply / expectation_value on a state sharded by its high index bits.

    ``make_local_op(strings, coeffs)`` returns an object with ``apply_into(out, src, accumulate)`` and
    ``expval(src_conj_side, src)``; ``exchange(send, recv, peer)`` swaps whole shards.  Defaults: CUDA + NCCL.
    """

    def __init__(self, strings: Sequence[str], coeffs: Sequence[complex], world: int, rank: int,
                 make_local_op: Callable | None = None, exchange: Callable | None = None):
        self.plan = plan_high_qubit(list(strings), list(coeffs), world)
        self.world, self.rank = world, rank
        self._make = make_local_op or _cuda_local_op
        self._exchange = exchange or _no_exchange
        self.local_ops = [self._make(c.strings, _rank_coeffs(c, rank), self.plan.n_local) for c in self.plan.classes]

    def apply(self, out, psi, recv_bufs, accumulate: bool = False):
        """out (+)= A psi for this rank's shard.  ``recv_bufs``: two shard-sized buffers for incoming peer shards."""
        classes = self.plan.classes
        swaps = [i for i, c in enumerate(classes) if c.x_hi]
        pending = {}
        if swaps:  # start the first swap before any arithmetic
            i0 = swaps[0]
            pending[i0] = self._exchange(psi, recv_bufs[0], self.rank ^ classes[i0].x_hi)
        first = not accumulate
        n_swapped = 0
        for i, (cls, op) in enumerate(zip(classes, self.local_ops)):
            if cls.x_hi == 0:
                op.apply_into(out, psi, accumulate=not first)
                first = False
                continue
            k = swaps.index(i)
            if k + 1 < len(swaps):  # next swap goes out while this class is applied
                nxt = swaps[k + 1]
                pending[nxt] = self._exchange(psi, recv_bufs[(k + 1) & 1], self.rank ^ classes[nxt].x_hi)
            pending.pop(i)()  # wait for this class' peer shard
            op.apply_into(out, recv_bufs[k & 1], accumulate=not first)
            first = False
            n_swapped += 1
        return n_swapped

    def apply_peer(self, out_ptr: int, shard_ptrs: Sequence[int], dtype, n_states: int = 1,
                   accumulate: bool = False) -> int:
        """Fused exchange + apply over NVLink peer memory: no receive buffers, no NCCL on the data path.

        ``shard_ptrs[r]`` is a device pointer to rank r's input shard as mapped into THIS process (its own shard for
        r == rank, CUDA-IPC mappings for the others: see :class:`PeerShards`).  Every class reads its source shard
        ``rank ^ x_hi`` straight through the kernel's gathers, so the transfer is overlapped with the arithmetic
        element by element and each remote amplitude crosses NVLink exactly once per class.
        The caller synchronises ranks around the call (all shards complete before, none modified until all done).
        """
        first = not accumulate
        n_remote = 0
        for cls, op in zip(self.plan.classes, self.local_ops):
            src = shard_ptrs[self.rank ^ cls.x_hi]
            op.apply_ptr(out_ptr, src, 1 << self.plan.n_local, n_states, dtype, accumulate=not first)
            first = False
            n_remote += cls.x_hi != 0
        return n_remote

    def expectation_value(self, psi, recv_bufs, all_reduce: Callable):
        """<psi|A|psi> per batch column: local partial sums (same swaps as apply) + one all-reduce."""
        classes = self.plan.classes
        total = None
        swaps = [i for i, c in enumerate(classes) if c.x_hi]
        pending = {}
        if swaps:
            pending[swaps[0]] = self._exchange(psi, recv_bufs[0], self.rank ^ classes[swaps[0]].x_hi)
        for i, (cls, op) in enumerate(zip(classes, self.local_ops)):
            if cls.x_hi == 0:
                part = op.expval(psi, psi)
            else:
                k = swaps.index(i)
                if k + 1 < len(swaps):
                    nxt = swaps[k + 1]
                    pending[nxt] = self._exchange(psi, recv_bufs[(k + 1) & 1], self.rank ^ classes[nxt].x_hi)
                pending.pop(i)()
                part = op.expval(psi, recv_bufs[k & 1])
            total = part if total is None else total + part
        return all_reduce(total)


# ------------------------------------------------------------------------------------------------ CUDA defaults
def _cuda_local_op(strings, coeffs, n_local):
    import fast_pauli_b200 as fp

    return _CudaLocalOp(fp, strings, coeffs)


class _CudaLocalOp:
    """A class' local operator on the GPU, addressed by raw device pointers (torch-free).  Used by the fused
    peer-memory path (``ShardedStateOp.apply_peer``); the exchange-based production path is the C ABI's
    ``fp_sharded_op_apply`` (:mod:`fast_pauli_b200.sharded`), which owns its context, streams and NCCL communicator."""

    def __init__(self, fp, strings, coeffs):
        self.fp = fp
        self.op = fp.PauliOp(coeffs, strings)

    def apply_ptr(self, out_ptr: int, src_ptr: int, dim: int, n_states: int, dtype, accumulate: bool):
        """``src_ptr`` may be a peer GPU's memory mapped through CUDA IPC."""
        import ctypes as C

        fp = self.fp
        ctx = self.op._context()
        fp._check(fp.lib.fp_op_apply(ctx._h, self.op._plan(np.dtype(dtype)), C.c_void_p(out_ptr), C.c_void_p(src_ptr),
                                     C.c_size_t(dim), C.c_size_t(n_states), C.c_int(int(accumulate))))

    def apply_into(self, out, src, accumulate: bool):
        raise NotImplementedError("exchange-based sharded apply on GPUs: use fast_pauli_b200.sharded.ShardedPauliOp "
                                  "(C ABI fp_sharded_op_apply, NCCL inside the library)")

    expval = apply_into


def _no_exchange(send, recv, peer):
    raise NotImplementedError("inject `exchange` (CPU tests) or use fast_pauli_b200.sharded.ShardedPauliOp on GPUs")


class PeerShards:
    """All ranks' input shards mapped into this process (CUDA IPC over NVLink / NVSwitch).

    ``local`` must be a :class:`fast_pauli_b200.DeviceArray` allocated by this library (a plain ``cudaMalloc``
    allocation, which is what ``cudaIpcGetMemHandle`` needs).  ``ptrs[r]`` is usable as the ``in`` pointer of any
    entry point; remote ones are read over NVLink by the kernels' ordinary 16-byte loads.
    """

    def __init__(self, local, dist=None):
        import ctypes as C

        import fast_pauli_b200 as fp

        if dist is None:
            import torch.distributed as dist
        self._fp, self._ctx = fp, local.ctx
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handle = (C.c_ubyte * 64)()
        fp._check(fp.lib.fp_ipc_export(self._ctx._h, C.c_void_p(local.ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        self.ptrs: list[int] = []
        self._opened: list[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(local.ptr)
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            fp._check(fp.lib.fp_ipc_open(self._ctx._h, buf, C.byref(p)))
            self.ptrs.append(int(p.value))
            self._opened.append(int(p.value))

    def close(self) -> None:
        import ctypes as C

        for p in self._opened:
            self._fp.lib.fp_ipc_close(self._ctx._h, C.c_void_p(p))
        self._opened = []
