"""Multi-process tests of the multi-GPU host logic on CPU: gloo backend, world_size 2 and 4, rendezvous on 127.0.0.1."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from __graft_entry__ import load_package

load_package()
from fast_pauli_b200 import distributed as fpd  # noqa: E402


def test_shard_columns_partition():
    for B in (1, 7, 16, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            blocks = [fpd.shard_columns(B, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_high_qubit_plan_groups_by_peer_offset():
    strings = ["XIZY", "ZIXX", "YYII", "IZZZ", "XXXX"]
    plan = fpd.plan_high_qubit(strings, [1, 2, 3, 4, 5], world=4)
    assert plan.n_local == 2
    offs = {c.x_hi: c for c in plan.classes}
    # top two characters decide the peer: XI -> 0b10, ZI -> 0, YY -> 0b11, IZ -> 0, XX -> 0b11
    assert sorted(offs) == [0, 2, 3]
    assert offs[0].strings == ["XX", "ZZ"] and offs[3].strings == ["II", "XX"] and offs[2].strings == ["ZY"]
    # (-i)^nY of the high characters is folded into the coefficient: "YY" -> (-i)^2 = -1
    assert offs[3].coeffs[0] == pytest.approx(-3)
    # rank-dependent sign (-1)^popc(rank & z_hi): "ZI" has z_hi = 0b10 -> flips on ranks 2, 3
    assert fpd._rank_coeffs(offs[0], 2)[0] == pytest.approx(-2) and fpd._rank_coeffs(offs[0], 1)[0] == pytest.approx(2)
    with pytest.raises(ValueError):
        fpd.plan_high_qubit(strings, [1] * 5, world=3)


@pytest.mark.parametrize("world", [2, 4])
def test_multi_process_gloo(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + world * 7 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist worker ok" in r.stdout
