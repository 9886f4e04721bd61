#!/usr/bin/env python
"""One small call of every kernel family added in the second session, for compute-sanitizer runs:

    compute-sanitizer --tool memcheck  python tests/sanitize_case.py
    compute-sanitizer --tool racecheck python tests/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # conftest helpers
from __graft_entry__ import load_package  # noqa: E402

fp = load_package()
from conftest import rand_states, rand_strings, rel_err  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ORC = orc.best()
rng = np.random.default_rng(3)


def span_strings(n, rank, S):
    gens = []
    while len(gens) < rank:
        m = int(rng.integers(1, 2**n))
        basis, ok = [], True
        for g in gens + [m]:
            v = g
            for b in basis:
                v = min(v, v ^ b)
            if v == 0:
                ok = False
                break
            basis.append(v)
        if ok:
            gens.append(m)
    out = []
    for k in range(S):
        sel = int(rng.integers(0, 2**rank)) if k >= rank else (1 << k)
        x = 0
        for j in range(rank):
            if (sel >> j) & 1:
                x ^= gens[j]
        z = int(rng.integers(0, 2**n))
        out.append("".join("IZXY"[2 * ((x >> (n - 1 - q)) & 1) + ((z >> (n - 1 - q)) & 1)] for q in range(n)))
    return out


ctx = fp.Context(0)
n, B = 9, 20
for rank in (2, 3, 4, 5, 8):  # K3c (2-4 c64 / 2-3 c128), K3d (4-5 c128), K3b (8)
    for dtype in (np.complex128, np.complex64):
        strings = span_strings(n, rank, 40)
        h = (rand_states(rng, 40, None, dtype) * 2 - (1 + 1j)).astype(dtype)
        psi = rand_states(rng, 2**n, B, dtype)
        op = fp.PauliOp(h, strings, ctx=ctx)
        t = 1e-12 if dtype == np.complex128 else 1e-5
        assert rel_err(op.apply(psi), ORC.op_apply(strings, h, psi, par=True)) < t
        assert rel_err(op.expectation_value(psi), ORC.op_expval(strings, h, psi, par=True)) < 10 * t
n = 11
strings = rand_strings(rng, n, 300)
for dtype in (np.complex64, np.complex128):  # K6b, K4c
    hk = (rand_states(rng, 300, 3, dtype) * 2 - (1 + 1j)).astype(dtype)
    psi = rand_states(rng, 2**n, 8, dtype)
    data = rng.random((3, 8)).astype(np.float32 if dtype == np.complex64 else np.float64)
    sop = fp.SummedPauliOp(strings, hk, ctx=ctx)
    up = (hk.astype(np.complex128), psi.astype(np.complex128), data.astype(np.float64))
    t = 1e-12 if dtype == np.complex128 else 1e-5
    assert rel_err(sop.apply_weighted(psi, data), ORC.sop_apply_weighted(strings, *up)) < t
    assert rel_err(sop.expectation_value(psi), ORC.sop_expval(strings, up[0], up[1])) < t
# K8
st = [str(p) for p in fp.helpers.calculate_pauli_strings_max_weight(4, 2)]
co = rng.uniform(-1, 1, (len(st), 3)) + 1j * rng.uniform(-1, 1, (len(st), 3))
s2 = fp.SummedPauliOp(st, co, ctx=ctx)
d = s2.to_tensor()
assert np.allclose(s2.square().to_tensor(), np.einsum("kij,kjl->kil", d, d), atol=1e-11)
# host pipeline
ctx.set_pipeline(True, 1, 2**9 * 20 * 16 // 8)
ps = fp.PauliString("XYZIXYZIX", ctx=ctx)
psi = rand_states(rng, 2**9, 20)
assert rel_err(ps.apply(psi, 0.5j), ORC.string_apply("XYZIXYZIX", psi, 0.5j)) < 1e-12
print("sanitize_case ok")
