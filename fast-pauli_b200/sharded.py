"""One state vector sharded by its high index bits over the GPUs of a box (BASELINE config 5) -- thin ctypes front-end
of the C ABI's ``fp_comm_*`` / ``fp_sharded_op_*`` entry points (csrc/sharded.cpp).

The data path is entirely inside the library: local fused kernels for the strings that leave the high bits alone,
chunked ``ncclSend`` / ``ncclRecv`` exchanges overlapped with the streaming single-string kernel for the rest.  This
module never imports torch; the only thing the ranks have to share beforehand is the 128-byte NCCL unique id, which
the caller distributes by any means (``bench.py`` uses torchrun's store, a file or a socket work just as well).

Neither sharding exists in the reference: its ``dim`` is an ``int`` shift (PS:57).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

import fast_pauli_b200 as fp

__all__ = ["unique_id", "Comm", "ShardedPauliOp", "bench_config5"]

ID_BYTES = 128
_lib = fp.lib


def unique_id() -> bytes:
    """A fresh NCCL unique id (call on ONE rank, ship the bytes to all ranks)."""
    buf = (C.c_ubyte * ID_BYTES)()
    fp._check(_lib.fp_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """NCCL communicator of the sharded calls: one per process / GPU (collective constructor)."""

    def __init__(self, ctx: "fp.Context", uid: bytes, world: int, rank: int):
        if len(uid) != ID_BYTES:
            raise ValueError(f"the NCCL unique id has {ID_BYTES} bytes")
        self.ctx, self.world, self.rank = ctx, world, rank
        self._h = C.c_void_p()
        buf = (C.c_ubyte * ID_BYTES).from_buffer_copy(uid)
        fp._check(_lib.fp_comm_create(ctx._h, buf, C.c_int(world), C.c_int(rank), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _lib.fp_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def barrier(self) -> None:
        fp._check(_lib.fp_comm_barrier(self._h))

    def allreduce(self, values: Sequence[float], op: str = "sum") -> list[float]:
        arr = (C.c_double * len(values))(*values)
        fp._check(_lib.fp_comm_allreduce_f64(self._h, arr, C.c_size_t(len(values)), C.c_int(1 if op == "max" else 0)))
        return list(arr)

    def nccl_version(self) -> int:
        v = C.c_int()
        fp._check(_lib.fp_comm_info(self._h, None, None, C.byref(v)))
        return int(v.value)

    def measure_p2p(self, nbytes: int = 1 << 30, iters: int = 5) -> float:
        """GB/s per direction per GPU of a pairwise ncclSend/ncclRecv swap with rank ^ 1 (the NVLink denominator)."""
        g = C.c_double()
        fp._check(_lib.fp_comm_measure_p2p(self._h, C.c_size_t(nbytes), C.c_int(iters), C.byref(g)))
        return float(g.value)


class ShardedPauliOp:
    """``PauliOp`` acting on a state whose rows are split over ``comm.world`` GPUs by the top index bits."""

    def __init__(self, comm: Comm, coeffs, strings: Sequence[str], dtype=np.complex128):
        self.comm = comm
        self.dtype = np.dtype(dtype)
        codes, n = fp._encode(list(strings))
        self.n_qubits = n
        self.n_local = n - int(np.log2(comm.world))
        h = np.ascontiguousarray(coeffs, dtype=self.dtype)
        if h.shape != (len(strings),):
            raise ValueError("coeffs and strings must have the same length")
        self._h = C.c_void_p()
        fp._check(_lib.fp_sharded_op_create(comm._h, C.c_int(fp._dtype_code(self.dtype)), C.c_int(n),
                                            C.c_size_t(len(strings)), C.c_void_p(codes.ctypes.data),
                                            C.c_void_p(h.ctypes.data), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _lib.fp_sharded_op_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def local_dim(self) -> int:
        return 1 << self.n_local

    def set_mode(self, mode: int) -> None:
        """0 automatic, 1 chunked streaming (two chunk buffers), 2 whole-shard receive + fused class operator."""
        fp._check(_lib.fp_sharded_op_set_mode(self._h, C.c_int(mode)))

    def set_chunk_bytes(self, nbytes: int) -> None:
        fp._check(_lib.fp_sharded_op_set_chunk_bytes(self._h, C.c_size_t(nbytes)))

    def apply(self, out: "fp.DeviceArray", psi: "fp.DeviceArray", accumulate: bool = False) -> None:
        """out_shard (+)= (A psi)_shard; device arrays of shape (local_dim,) or (local_dim, B). Collective."""
        B = 1 if psi.ndim == 1 else psi.shape[1]
        fp._check(_lib.fp_sharded_op_apply(self._h, C.c_void_p(out.ptr), C.c_void_p(psi.ptr),
                                           C.c_size_t(psi.shape[0]), C.c_size_t(B), C.c_int(int(accumulate))))

    def expectation_value(self, psi: "fp.DeviceArray", work: "fp.DeviceArray") -> np.ndarray:
        """<psi|A|psi> per column, identical on every rank; ``work`` is a shard-sized scratch array. Collective."""
        B = 1 if psi.ndim == 1 else psi.shape[1]
        out = np.zeros(B, dtype=self.dtype)
        fp._check(_lib.fp_sharded_op_expval(self._h, C.c_void_p(out.ctypes.data), C.c_void_p(psi.ptr),
                                            C.c_void_p(work.ptr), C.c_size_t(psi.shape[0]), C.c_size_t(B)))
        return out

    def info(self) -> dict:
        n = C.c_size_t()
        offs = (C.c_uint64 * max(self.comm.world, 1))()
        sent, chunks, kernels = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ms = C.c_float()
        mode = C.c_int()
        fp._check(_lib.fp_sharded_op_info(self._h, C.byref(n), offs, C.byref(sent), C.byref(chunks), C.byref(kernels)))
        fp._check(_lib.fp_sharded_op_last_ms(self._h, C.byref(ms)))
        fp._check(_lib.fp_sharded_op_last_mode(self._h, C.byref(mode)))
        return {"mode_last": {1: "chunked", 2: "whole-shard"}.get(int(mode.value), "none"),"n_remote_classes": int(n.value), "peer_offsets": [int(offs[i]) for i in range(n.value)],
                "bytes_sent_last": int(sent.value), "chunks_last": int(chunks.value), "kernels_last": int(kernels.value),
                "device_ms_last": float(ms.value)}


def _masks(string: str) -> tuple[int, int, int]:
    n = len(string)
    x = z = ny = 0
    for q, ch in enumerate(string):
        bit = 1 << (n - 1 - q)
        if ch in "XY":
            x |= bit
        if ch in "YZ":
            z |= bit
        ny += ch == "Y"
    return x, z, ny


def closed_form_rows(strings, coeffs, rows: np.ndarray, seed: int) -> np.ndarray:
    """(A psi)[rows] of the counter-generated single state (B = 1), numpy only: the sampled-row check of a state no
    host can hold."""
    from fast_pauli_b200.synth import uniform_complex_at

    rows = np.asarray(rows, dtype=np.uint64)
    out = np.zeros(len(rows), dtype=np.complex128)
    for s, h in zip(strings, coeffs):
        x, z, ny = _masks(s)
        par = np.array([bin(int(r) & z).count("1") & 1 for r in rows])
        out += h * ((-1j) ** ny) * (1 - 2 * par) * uniform_complex_at(rows ^ np.uint64(x), np.complex128, seed)
    return out


def bench_config5(fp_mod, ctx, dist, torch, rank: int, world: int, local_rank: int, seed: int = 18,
                  string_seed: int = 1234, n_qubits: int | None = None, n_strings: int = 16, mode: int = 0) -> dict:
    """BASELINE config 5 scaled to `world` ranks: PauliOp.apply on one complex128 state of 31 + log2(world) qubits
    (34 qubits = 256 GiB at 8 ranks), 32 GiB shard in + 32 GiB out per GPU.  Parity on sampled rows before timing;
    device time of the slowest rank; NVLink GB/s per direction against a pairwise ncclSend/Recv probe measured in the
    same run.  `dist` / `torch` are only used to broadcast the NCCL unique id."""
    from fast_pauli_b200.synth import random_strings

    n_hi = int(np.log2(world))
    n = n_qubits if n_qubits is not None else 31 + n_hi
    n_local = n - n_hi
    local_dim = 1 << n_local
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = Comm(ctx, box[0], world, rank)
    rng = np.random.default_rng(string_seed + 5)
    strings = random_strings(rng, n, n_strings)
    h = rng.uniform(-1, 1, n_strings) + 1j * rng.uniform(-1, 1, n_strings)
    op = ShardedPauliOp(comm, h, strings)
    op.set_mode(mode)
    psi = ctx.uniform((local_dim,), np.complex128, seed=seed, first=rank * local_dim)
    out = ctx.empty((local_dim,), np.complex128)
    try:
        op.apply(out, psi)  # warm-up + the result that is checked
        prng = np.random.default_rng(7 + rank)
        rows_local = np.unique(np.concatenate([prng.integers(0, local_dim, size=14), [0, local_dim - 1]]))
        got = np.array([out.get_rows(int(r), int(r) + 1)[0] for r in rows_local])
        exp = closed_form_rows(strings, h, rows_local.astype(np.uint64) + np.uint64(rank * local_dim), seed)
        err = float(np.max(np.abs(got - exp)) / max(np.max(np.abs(exp)), 1e-300))
        p2p = comm.measure_p2p(1 << 30, 5)
        iters = 3
        ms = []
        for _ in range(iters):
            comm.barrier()
            op.apply(out, psi)
            ms.append(op.info()["device_ms_last"])
        info = op.info()
        best = comm.allreduce([min(ms)], "max")[0]  # every rank's best, slowest rank
        err_all = comm.allreduce([err], "max")[0]
        sent = info["bytes_sent_last"]
        gbps = sent / (best * 1e-3) / 1e9
        res = {"workload": f"PauliOp.apply, one {n}-qubit complex128 state ({(16 << n) / 2**30:.0f} GiB) sharded by the "
                           f"top {n_hi} index bits over {world} GPUs, {n_strings} random strings",
               "n_qubits": n, "n_strings": n_strings, "n_gpus": world, "shard_GiB": (16 << n_local) / 2**30,
               "peer_offsets_in_use": info["peer_offsets"], "ms": best, "amp_strings_per_s": (1 << n) * n_strings / (best * 1e-3),
               "parity_max_rel_err": err_all, "parity_tol": 1e-12, "parity": "ok" if err_all < 1e-12 else "FAILED",
               "parity_check": "16 sampled rows per rank against the closed form on the regenerated state",
               "nvlink_bytes_sent_per_gpu": sent, "nvlink_GBps_per_direction": gbps,
               "nvlink_peak_GBps_measured": p2p, "nvlink_frac": gbps / p2p if p2p > 0 else None,
               "nvlink_peak_source": "pairwise ncclSend/ncclRecv swap of 1 GiB with rank^1, same run (fp_comm_measure_p2p)",
               "exchange_mode": info["mode_last"], "transfers": info["chunks_last"], "kernel_calls": info["kernels_last"],
               "exchange": "ncclSend/ncclRecv in ncclGroupStart/End on a communication stream, double-buffered against "
                           "the kernels: whole-shard mode = one fused PauliOp per peer offset on the received shard, "
                           "chunked mode = 256 MiB chunks + the streaming single-string kernel (C ABI: "
                           "fp_sharded_op_apply; no torch on the data path)",
               "nccl_version": comm.nccl_version(), "timing": "CUDA events inside the call, best of 3, max over ranks"}
    finally:
        op.close()
        comm.close()
    return res
