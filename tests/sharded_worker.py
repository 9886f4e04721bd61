"""Worker of tests/test_sharded.py::test_two_gpus_nccl_exchange (torchrun, one process per GPU).

torch.distributed (gloo) only broadcasts the NCCL unique id and gathers the result for the check; the sharded apply and
expectation value run through the C ABI (NCCL inside the library).  The checker is the CPU oracle.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

os.environ.setdefault("FASTPAULI_DEVICE", os.environ.get("LOCAL_RANK", "0"))
fp = load_package()
from fast_pauli_b200 import sharded  # noqa: E402
from fast_pauli_b200.synth import random_strings  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = fp.Context(int(os.environ.get("LOCAL_RANK", "0")))
    box = [sharded.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = sharded.Comm(ctx, box[0], world, rank)
    rng = np.random.default_rng(11)  # same stream on every rank: identical global problem
    for n, B, S, chunk, mode in ((12, None, 20, 4096, 1), (14, 3, 12, 1 << 16, 2), (16, None, 30, 1 << 18, 1),
                                 (16, None, 30, 1 << 18, 0), (13, 2, 9, 1 << 12, 2)):
        strings = random_strings(rng, n, S) + ["X" + "Z" * (n - 1), "I" * n]
        h = rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))
        shape = (1 << n,) if B is None else (1 << n, B)
        psi = rng.random(shape) + 1j * rng.random(shape)
        local = (1 << n) // world
        op = sharded.ShardedPauliOp(comm, h, strings)
        op.set_chunk_bytes(chunk)
        op.set_mode(mode)
        d_in = ctx.to_device(np.ascontiguousarray(psi[rank * local:(rank + 1) * local]))
        d_out = ctx.empty(d_in.shape, np.complex128)
        op.apply(d_out, d_in)
        mine = d_out.get()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        got = np.concatenate(parts, axis=0)
        ref = orc.best().op_apply(strings, h, psi)
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        assert err < 1e-12, f"sharded apply n={n} B={B}: {err:.3e}"
        info = op.info()
        assert info["n_remote_classes"] >= 1 and info["bytes_sent_last"] > 0
        assert info["mode_last"] == ("chunked" if mode == 1 else "whole-shard")
        ev = op.expectation_value(d_in, d_out)
        ev_ref = orc.best().op_expval(strings, h, psi if B is not None else psi[:, None])
        err_e = float(np.max(np.abs(ev - ev_ref)) / np.max(np.abs(ev_ref)))
        assert err_e < 1e-12, f"sharded expectation value n={n} B={B}: {err_e:.3e}"
        op.close()
    gbps = comm.measure_p2p(1 << 28, 3)
    assert gbps > 1.0
    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"sharded worker ok (p2p {gbps:.0f} GB/s per direction, NCCL {comm.world} ranks)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
