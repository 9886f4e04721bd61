#!/usr/bin/env python
"""Which kernel family should take a sparse operator of small x-mask rank?  20 qubits x 64 complex128 columns,
G distinct x-masks drawn from a random GF(2) span of rank r, `zv` z-variants per mask; times PauliOp.apply and
expectation_value under  default dispatch | FP64 tensor-core cosets off (register cosets, rank <= 4, else K3b) |
register cosets off as well (shared-memory cosets K3b).  One process per configuration (the knobs are read when the
context is created).   python scripts/dispatch_sweep.py            # driver: prints the table
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(r, g, zv) for r in (3, 4, 5, 6) for g in (r, 8, 12, 16, 24, 32) if r <= g <= (1 << r) for zv in (1, 4)]


def worker():
    from __graft_entry__ import load_package
    fp = load_package()
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from run_case import timed, vp, sz

    ctx = fp.Context(0)
    ctx.set_async(True)
    if os.environ.get("SWEEP_RCOSET") == "0":
        ctx.set_rcoset(0)
    n, B = 20, 64
    psi = ctx.uniform((1 << n, B), np.complex128)
    y = ctx.empty((1 << n, B), np.complex128)
    ev = ctx.empty((B,), np.complex128)
    for r, g, zv in CASES:
        rng = np.random.default_rng(100 * r + g)
        gens = [int(rng.integers(1, 1 << n)) for _ in range(r)]
        # the r generators themselves (so the rank is r) plus random further members of the span
        combos = [1 << j for j in range(r)]
        rest = [k for k in range(1, 1 << r) if k not in combos]
        rng.shuffle(rest)
        combos = (combos + rest)[:g]
        strings = []
        for k in combos:
            x = 0
            for j in range(r):
                if (k >> j) & 1:
                    x ^= gens[j]
            for _ in range(zv):
                z = int(rng.integers(0, 1 << n))
                strings.append("".join("IZXY"[2 * ((x >> (n - 1 - q)) & 1) + ((z >> (n - 1 - q)) & 1)] for q in range(n)))
        h = rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))
        op = fp.PauliOp(h, strings, ctx=ctx)
        plan = op._plan(np.complex128)
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), 5)
        ms2 = timed(ctx, lambda: fp.lib.fp_op_expval(ctx._h, plan, vp(ev.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), 5)
        print(f"{r} {g} {zv} {ms:.4f} {ms2:.4f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
        sys.exit(0)
    table = {}
    for tag, env in (("default", {}), ("no_dmma", {"FASTPAULI_DCOSET": "0"}),
                     ("k3b", {"FASTPAULI_DCOSET": "0", "SWEEP_RCOSET": "0"})):
        out = subprocess.run([sys.executable, __file__, "worker"], env=dict(os.environ, **env), capture_output=True,
                             text=True, timeout=600)
        if out.returncode != 0:
            print(tag, "FAILED", out.stderr[-1500:])
            continue
        for line in out.stdout.splitlines():
            p = line.split()
            if len(p) == 5:
                table.setdefault((int(p[0]), int(p[1]), int(p[2])), {})[tag] = (float(p[3]), float(p[4]))
    print("rank masks zvar | apply ms: default no_dmma k3b | expval ms: default no_dmma k3b")
    for key in sorted(table):
        row = table[key]
        a = " ".join(f"{row[t][0]:7.3f}" if t in row else "   -   " for t in ("default", "no_dmma", "k3b"))
        e = " ".join(f"{row[t][1]:7.3f}" if t in row else "   -   " for t in ("default", "no_dmma", "k3b"))
        print(f"{key[0]:4d} {key[1]:5d} {key[2]:4d} | {a} | {e}")
