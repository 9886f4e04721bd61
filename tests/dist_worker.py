"""Worker for tests/test_distributed.py: world_size-2 (or more) run of the multi-GPU host logic on CPU (gloo).

The arithmetic stand-in is the CPU oracle (test infrastructure); what is under test is the sharding plan, the peer
exchange schedule, the rank-dependent signs and the gather / all-reduce of results.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from fast_pauli_b200 import distributed as fpd  # noqa: E402
from fast_pauli_b200.synth import random_strings  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ORC = orc.port()


class OracleLocalOp:
    def __init__(self, strings, coeffs, n_local):
        self.strings, self.coeffs = list(strings), np.asarray(coeffs, dtype=np.complex128)

    def apply_into(self, out, src, accumulate):
        o = out.numpy()
        if not accumulate:
            o[...] = 0
        ORC.op_apply(self.strings, self.coeffs, src.numpy(), out=o)

    def expval(self, bra, src):
        tmp = np.zeros_like(src.numpy())
        ORC.op_apply(self.strings, self.coeffs, src.numpy(), out=tmp)
        return torch.from_numpy((bra.numpy().conj() * tmp).sum(axis=0))


def cpu_exchange(send, recv, peer):
    reqs = [dist.isend(send, peer), dist.irecv(recv, peer)]

    def wait():
        for r in reqs:
            r.wait()

    return wait


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(7)  # same stream on every rank: identical global problem
    # ---- batch-axis sharding: disjoint column blocks, results gathered in order
    n, B, S = 6, 11, 20
    strings = random_strings(rng, n, S)
    h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
    psi = rng.random((2**n, B)) + 1j * rng.random((2**n, B))
    a, b = fpd.shard_columns(B, world, rank)
    local = np.ascontiguousarray(psi[:, a:b])
    ev_local = ORC.op_expval(strings, h, local) if b > a else np.zeros(0, np.complex128)
    ev = fpd.gather_columns(ev_local, B, dist)
    np.testing.assert_allclose(ev, ORC.op_expval(strings, h, psi), rtol=1e-13)
    covered = sorted(sum(([*range(*fpd.shard_columns(B, world, r))] for r in range(world)), []))
    assert covered == list(range(B))

    # ---- high-qubit sharding: rows split by the top log2(world) index bits, pairwise shard swaps
    for n, B, S in [(7, 3, 40), (5, 1, 12)]:
        strings = random_strings(rng, n, S)
        h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
        psi = rng.random((2**n, B)) + 1j * rng.random((2**n, B))
        expect = ORC.op_apply(strings, h, psi)
        expect_ev = ORC.op_expval(strings, h, psi)
        n_loc_rows = 2**n // world
        mine = torch.from_numpy(np.ascontiguousarray(psi[rank * n_loc_rows:(rank + 1) * n_loc_rows]))
        out = torch.zeros_like(mine)
        bufs = [torch.empty_like(mine), torch.empty_like(mine)]
        op = fpd.ShardedStateOp(strings, h, world, rank, make_local_op=OracleLocalOp, exchange=cpu_exchange)
        n_sw = op.apply(out, mine, bufs)
        assert n_sw == len(op.plan.peer_offsets()) <= world - 1  # each non-zero peer offset is exchanged exactly once
        np.testing.assert_allclose(out.numpy(), expect[rank * n_loc_rows:(rank + 1) * n_loc_rows], rtol=1e-12, atol=1e-12)
        # accumulate on top of existing output
        out2 = out.clone()
        op.apply(out2, mine, bufs, accumulate=True)
        np.testing.assert_allclose(out2.numpy(), 2 * out.numpy(), rtol=1e-12, atol=1e-12)

        def all_reduce(t):
            t = t.clone()
            dist.all_reduce(t)
            return t

        ev = op.expectation_value(mine, bufs, all_reduce)
        np.testing.assert_allclose(ev.numpy(), expect_ev, rtol=1e-12)
    dist.barrier()
    if rank == 0:
        print("dist worker ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
