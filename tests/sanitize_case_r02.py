#!/usr/bin/env python
"""One small call of every kernel family added in round 2, for compute-sanitizer runs:

    compute-sanitizer --tool memcheck  python tests/sanitize_case_r02.py
    compute-sanitizer --tool racecheck python tests/sanitize_case_r02.py

K3e (coset_few_kernel, overwrite + read-modify-write + expectation value), K3f (coset_few_tma_kernel), K3g
(coset_gen_tma_kernel), K3i (coset_dir_tma_kernel, overwrite + read-modify-write), K3j (coset_pair_tma_kernel, overwrite +
read-modify-write) in both precisions, and K3d on
complex64 batches; every result is checked against the oracle and the launch path is asserted.
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
rng = np.random.default_rng(5)
ctx = fp.Context(0)


def variants(xs, per):
    out = []
    for s in xs:
        for _ in range(per):
            t = list(s)
            for q in range(len(t)):
                if rng.random() < 0.5:
                    t[q] = {"X": "Y", "Y": "X", "I": "Z", "Z": "I"}[t[q]]
            out.append("".join(t))
    return out


def check(strings, n, B, dtype, want_mask, expval=False):
    h = (rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))).astype(dtype)
    psi = rand_states(rng, 2**n, B, dtype)
    d_psi = ctx.to_device(psi)
    op = fp.PauliOp(h, strings, ctx=ctx)
    t = 1e-12 if dtype == np.complex128 else 1e-5
    ctx.coset_kernels_used(reset=True)
    got = op.apply(d_psi).get()
    used = ctx.coset_kernels_used()
    assert used & want_mask, (used, want_mask)
    ref = ORC.op_apply(strings, h.astype(np.complex128), psi.astype(np.complex128), par=True)
    assert rel_err(got, ref) < t
    if expval:
        ev = op.expectation_value(d_psi).get()
        assert rel_err(ev, ORC.op_expval(strings, h.astype(np.complex128), psi.astype(np.complex128), par=True)) < 10 * t
    print(f"ok: n={n} B={B} {np.dtype(dtype).name} strings={len(strings)} kernels={used}")


for dtype in (np.complex128, np.complex64):
    wide = 1 if dtype == np.complex128 else 2
    # K3e: 8 masks x 4 strings on a small register (overwrite), 12 masks (second pass accumulates), expectation value
    ctx.set_coset(2, 4, 8)  # rank-8 tiles of 16 vectors per row: the shape the few-mask / TMA-fed kernels take
    check(variants(rand_strings(rng, 12, 8), 4), 12, 16 * wide, dtype, 2, expval=True)
    check(variants(rand_strings(rng, 12, 12), 3), 12, 16 * wide, dtype, 2)
    # K3f: 16 qubits, 8 masks, overwrite pass through the TMA
    check(variants(rand_strings(rng, 16, 8), 2), 16, 16 * wide, dtype, 4)
    # K3g: low-weight strings, more than 8 masks per pass
    check(rand_strings(rng, 13, 150, max_weight=3), 13, 32 * wide, dtype, 8)
    # K3i: random strings (one per mask), enough tiles for the persistent grid; 12 strings = overwrite + accumulate
    check(rand_strings(rng, 14, 12), 14, 256 * wide, dtype, 16)
    # K3j: eight independent masks x 3 strings (overwrite: table build + paired gathers), 16 masks x 2 strings (the
    # second pass is read-modify-write)
    check(variants(rand_strings(rng, 14, 8), 3), 14, 256 * wide, dtype, 32)
    check(variants(rand_strings(rng, 14, 16), 2), 14, 256 * wide, dtype, 32)
    ctx.set_coset(1)

# K3d on complex64: dense 5-local operator
os.environ["FASTPAULI_DCOSET"] = "2"
ctx2 = fp.Context(0)
del os.environ["FASTPAULI_DCOSET"]
n = 10
pos = sorted(int(p) for p in rng.choice(n, size=5, replace=False))
strings = []
for k in range(4**5):
    t = ["I"] * n
    for i, p_ in enumerate(pos):
        t[p_] = "IXYZ"[(k >> (2 * i)) & 3]
    strings.append("".join(t))
h = (rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))).astype(np.complex64)
psi = rand_states(rng, 2**n, 40, np.complex64)
op = fp.PauliOp(h, strings, ctx=ctx2)
l0 = ctx2.launch_count
got = op.apply(psi)
assert ctx2.launch_count - l0 == 1
up = (h.astype(np.complex128), psi.astype(np.complex128))
assert rel_err(got, ORC.op_apply(strings, *up, par=True)) < 1e-5
assert rel_err(op.expectation_value(psi), ORC.op_expval(strings, *up, par=True)) < 1e-4
print("ok: K3d complex64")
print("sanitize_case_r02: all checks passed")
