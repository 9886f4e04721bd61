"""K7 (BASELINE config 5) through the C ABI: one state sharded by its high index bits.

* `-m gpu`, one GPU: every rank of a world of 2 / 4 / 8 is played in turn by the single-process stand-in
  (fp_comm_create_emulated + fp_sharded_op_apply_emulated): classes, chunk schedule, block signs and kernels are the
  production code, only the NCCL transport is replaced by a device copy.  The gathered result must equal the CPU
  oracle's PauliOp.apply on the whole state (1e-12 / 1e-5).
* `-m gpu`, two or more GPUs: the real thing under torchrun (tests/sharded_worker.py): fp_comm_create,
  ncclSend/ncclRecv exchange, expectation value with the all-reduce.
* CPU: the entry points exist and refuse to run without a device (tests/test_abi.py covers the export list).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, TOL, rand_states, rand_strings, rel_err
from __graft_entry__ import load_package
from oracle import oracle as orc

fp = load_package()
ORC = orc.best()


def emulated_apply(strings, h, psi, world, dtype, chunk_bytes, accumulate=False, out0=None, mode=1):
    """Gather of the emulated per-rank results: (dim,) or (dim, B)."""
    ctx = fp.default_context()
    dt = np.dtype(dtype)
    dim = psi.shape[0]
    B = 1 if psi.ndim == 1 else psi.shape[1]
    local = dim // world
    codes, n = fp._encode(list(strings))
    hh = np.ascontiguousarray(h, dtype=dt)
    d_all = ctx.to_device(np.ascontiguousarray(psi, dtype=dt))
    result = np.zeros_like(np.ascontiguousarray(psi, dtype=dt))
    for rank in range(world):
        comm = C.c_void_p()
        fp._check(fp.lib.fp_comm_create_emulated(ctx._h, C.c_int(world), C.c_int(rank), C.byref(comm)))
        op = C.c_void_p()
        fp._check(fp.lib.fp_sharded_op_create(comm, C.c_int(fp._dtype_code(dt)), C.c_int(n), C.c_size_t(len(strings)),
                                              C.c_void_p(codes.ctypes.data), C.c_void_p(hh.ctypes.data), C.byref(op)))
        fp._check(fp.lib.fp_sharded_op_set_chunk_bytes(op, C.c_size_t(chunk_bytes)))
        fp._check(fp.lib.fp_sharded_op_set_mode(op, C.c_int(mode)))
        shape = (local,) if psi.ndim == 1 else (local, B)
        if out0 is not None:
            d_out = ctx.to_device(np.ascontiguousarray(out0[rank * local:(rank + 1) * local], dtype=dt))
        else:
            d_out = ctx.empty(shape, dt)
        fp._check(fp.lib.fp_sharded_op_apply_emulated(op, C.c_void_p(d_out.ptr), C.c_void_p(d_all.ptr), C.c_size_t(local),
                                                      C.c_size_t(B), C.c_int(int(accumulate))))
        result[rank * local:(rank + 1) * local] = d_out.get()
        fp.lib.fp_sharded_op_destroy(op)
        fp.lib.fp_comm_destroy(comm)
    return result


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("world,n,B,chunk", [(2, 10, None, 1 << 20), (4, 11, None, 4096), (8, 12, None, 2048),
                                             (4, 10, 3, 4096), (8, 13, 4, 16384), (2, 9, 16, 1024)])
def test_emulated_sharded_apply_matches_oracle(mode, dtype, world, n, B, chunk, rng):
    strings = rand_strings(rng, n, 24)
    strings += ["I" * n, "Z" * n, "X" + "I" * (n - 1), "Y" * n]  # identity, diagonal, pure high-bit flip, all-Y
    h = rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))
    psi = rand_states(rng, 1 << n, B, dtype)
    got = emulated_apply(strings, h, psi, world, dtype, chunk, mode=mode)
    ref = ORC.op_apply(strings, h.astype(dtype), psi)
    assert rel_err(got, ref) < TOL[np.dtype(dtype)]


@pytest.mark.gpu
def test_emulated_sharded_apply_accumulates(rng):
    n, world = 10, 4
    strings = rand_strings(rng, n, 9)
    h = rng.uniform(-1, 1, 9) + 1j * rng.uniform(-1, 1, 9)
    psi = rand_states(rng, 1 << n, None)
    out0 = rand_states(rng, 1 << n, None)
    ref = out0.copy()
    ORC.op_apply(strings, h, psi, out=ref)  # the C++ methods accumulate (PO:419,432)
    for mode in (1, 2):
        got = emulated_apply(strings, h, psi, world, np.complex128, 2048, accumulate=True, out0=out0, mode=mode)
        assert rel_err(got, ref) < 1e-12


@pytest.mark.gpu
def test_emulated_only_remote_classes(rng):
    # no string leaves the high bits alone: the output has to be zero-initialised by the call itself
    n, world = 9, 4
    strings = ["XX" + s for s in rand_strings(rng, n - 2, 5)] + ["YI" + s for s in rand_strings(rng, n - 2, 3)]
    h = rng.uniform(-1, 1, 8) + 1j * rng.uniform(-1, 1, 8)
    psi = rand_states(rng, 1 << n, 2)
    for mode in (1, 2):
        got = emulated_apply(strings, h, psi, world, np.complex128, 1024, mode=mode)
        assert rel_err(got, ORC.op_apply(strings, h, psi)) < 1e-12


@pytest.mark.gpu
def test_sharded_argument_errors():
    ctx = fp.default_context()
    comm = C.c_void_p()
    assert fp.lib.fp_comm_create_emulated(ctx._h, C.c_int(3), C.c_int(0), C.byref(comm)) == 1  # not a power of two
    fp._check(fp.lib.fp_comm_create_emulated(ctx._h, C.c_int(2), C.c_int(1), C.byref(comm)))
    codes, n = fp._encode(["XYZI"])
    h = np.array([1.0 + 0j])
    op = C.c_void_p()
    fp._check(fp.lib.fp_sharded_op_create(comm, C.c_int(fp.FP_C128), C.c_int(n), C.c_size_t(1),
                                          C.c_void_p(codes.ctypes.data), C.c_void_p(h.ctypes.data), C.byref(op)))
    d = ctx.empty((8,), np.complex128)
    # wrong shard size -> the reference's dimension error (PO:343-346), host pointer -> refused
    assert fp.lib.fp_sharded_op_apply_emulated(op, C.c_void_p(d.ptr), C.c_void_p(d.ptr), C.c_size_t(4), C.c_size_t(1),
                                               C.c_int(0)) == 1
    host = np.zeros(16, dtype=np.complex128)
    assert fp.lib.fp_sharded_op_apply_emulated(op, C.c_void_p(host.ctypes.data), C.c_void_p(host.ctypes.data),
                                               C.c_size_t(8), C.c_size_t(1), C.c_int(0)) == 1
    # the real entry point refuses an emulated communicator
    assert fp.lib.fp_sharded_op_apply(op, C.c_void_p(d.ptr), C.c_void_p(d.ptr), C.c_size_t(8), C.c_size_t(1), C.c_int(0)) == 1
    fp.lib.fp_sharded_op_destroy(op)
    fp.lib.fp_comm_destroy(comm)


def _n_gpus() -> int:
    n = C.c_int()
    return int(n.value) if fp.lib.fp_device_count(C.byref(n)) == 0 else 0


@pytest.mark.gpu
def test_two_gpus_nccl_exchange():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_sharded.py -m gpu)")
    world = 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29800 + (os.getpid() % 150)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded worker ok" in r.stdout


def test_sharded_entry_points_need_a_device():
    """CPU box: the symbols exist; without a GPU nothing computes (and nothing falls back)."""
    for name in ("fp_comm_unique_id", "fp_comm_create", "fp_comm_create_emulated", "fp_comm_destroy", "fp_comm_barrier",
                 "fp_comm_allreduce_f64", "fp_comm_measure_p2p", "fp_sharded_op_create", "fp_sharded_op_apply",
                 "fp_sharded_op_apply_emulated", "fp_sharded_op_expval", "fp_sharded_op_info", "fp_sharded_op_last_ms",
                 "fp_sharded_op_set_chunk_bytes", "fp_sharded_op_set_mode", "fp_sharded_op_last_mode",
                 "fp_sharded_op_destroy"):
        assert hasattr(fp.lib, name), name
    if _n_gpus() == 0:
        ctxp = C.c_void_p()
        assert fp.lib.fp_ctx_create(C.c_int(0), C.byref(ctxp)) == 3  # FP_NO_DEVICE
        comm = C.c_void_p()
        assert fp.lib.fp_comm_create_emulated(None, C.c_int(2), C.c_int(0), C.byref(comm)) == 1
