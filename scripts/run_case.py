#!/usr/bin/env python
"""Run ONE device-resident workload a few times (for ncu captures and quick timing).

    python scripts/run_case.py few20|rand20|cfg3|cfg4w|cfg4e|str20 [--iters N] [--coset MODE] [--log-twc V]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

fp = load_package()
from fast_pauli_b200.synth import random_strings  # noqa: E402


def vp(p):
    return C.c_void_p(p)


def sz(v):
    return C.c_size_t(v)


def timed(ctx, fn, iters):
    fn()
    ctx.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    fp.lib.fp_event_create(C.byref(e0))
    fp.lib.fp_event_create(C.byref(e1))
    fp.lib.fp_event_record(ctx._h, e0)
    for _ in range(iters):
        fn()
    fp.lib.fp_event_record(ctx._h, e1)
    ms = C.c_float()
    fp.lib.fp_event_elapsed_ms(e0, e1, C.byref(ms))
    return ms.value / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--coset", type=int, default=1)
    ap.add_argument("--log-twc", type=int, default=-1)
    ap.add_argument("--log-nt", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--few", type=int, default=-1, help="fp_ctx_set_coset_few mode (0 general kernel, 1 default dispatch, 2 K3e only, 3 K3e/K3f)")
    ap.add_argument("--c128", action="store_true", help="cfg4w / cfg4e in complex128 instead of complex64")
    ap.add_argument("--c64", action="store_true", help="span* / local* in complex64 instead of complex128")
    a = ap.parse_args()
    ctx = fp.Context(0)
    ctx.set_coset(a.coset, a.log_twc, a.log_nt)
    ctx.set_async(True)
    if a.few >= 0:
        ctx.set_coset_few(a.few)
    rng = np.random.default_rng(1234)
    if a.case in ("few20", "rand20", "few20low", "few20one"):
        n, B = 20, a.batch or 64
        if a.case in ("few20", "few20low", "few20one"):
            xs = random_strings(rng, n, 8)
            if a.case == "few20low":  # x-masks confined to the 8 low qubits: cosets are contiguous 256-row blocks
                xs = ["".join("IZ"[int(rng.integers(0, 2))] for _ in range(n - 8)) + s[n - 8:] for s in xs]
            strings = []
            for s in xs:
                for _ in range(1 if a.case == "few20one" else 8):
                    t = list(s)
                    for q in range(n):
                        if rng.random() < 0.5:
                            t[q] = {"X": "Y", "Y": "X", "I": "Z", "Z": "I"}[t[q]]
                    strings.append("".join(t))
        else:
            strings = random_strings(rng, n, 64)
        cdt = np.complex64 if a.c64 else np.complex128
        h = (rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))).astype(cdt)
        psi = ctx.uniform((1 << n, B), cdt)
        op = fp.PauliOp(h, strings, ctx=ctx)
        y = ctx.empty((1 << n, B), cdt)
        plan = op._plan(cdt)
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        amps = (1 << n) * B
        ev = ctx.empty((B,), cdt)
        ms2 = timed(ctx, lambda: fp.lib.fp_op_expval(ctx._h, plan, vp(ev.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        bpa = 16 if a.c64 else 32
        print(f"{a.case}{' c64' if a.c64 else ''}: {ms:.3f} ms  {amps*bpa/ms/1e6:.0f} GB/s algorithmic  groups={op.plan_info()['n_x_groups']} | "
              f"expval {ms2:.3f} ms {amps*bpa/2/ms2/1e6:.0f} GB/s  kernels_used={ctx.coset_kernels_used()}")
    elif a.case in ("span1", "span2", "span3", "span4", "span5", "local2", "local3", "local4", "local5"):
        # x-masks confined to a GF(2) span of rank r (register-resident coset kernel): 64 strings over 2^r masks;
        # local3 = all 64 Pauli strings on 3 fixed qubits (8 x-masks x 8 z-masks)
        n, B = 20, a.batch or 64
        if a.case.startswith("local"):
            kq = int(a.case[-1])
            pos = sorted(int(p) for p in rng.choice(n, size=kq, replace=False))
            strings = []
            for k in range(4**kq):
                t = ["I"] * n
                for i, p_ in enumerate(pos):
                    t[p_] = "IXYZ"[(k >> (2 * i)) & 3]
                strings.append("".join(t))
        else:
            r = int(a.case[-1])
            gens = [int(rng.integers(1, 1 << n)) for _ in range(r)]
            strings = []
            for k in range(max(64, 4 << r)):
                x = 0
                for j in range(r):
                    if (k >> j) & 1:
                        x ^= gens[j]
                z = int(rng.integers(0, 1 << n))
                strings.append("".join("IZXY"[2 * ((x >> (n - 1 - q)) & 1) + ((z >> (n - 1 - q)) & 1)] for q in range(n)))
        cdt = np.complex64 if a.c64 else np.complex128
        h = (rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))).astype(cdt)
        psi = ctx.uniform((1 << n, B), cdt)
        op = fp.PauliOp(h, strings, ctx=ctx)
        y = ctx.empty((1 << n, B), cdt)
        ev = ctx.empty((B,), cdt)
        plan = op._plan(cdt)
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        ms2 = timed(ctx, lambda: fp.lib.fp_op_expval(ctx._h, plan, vp(ev.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        amps = (1 << n) * B
        bpa = 16 if a.c64 else 32
        print(f"{a.case}{' c64' if a.c64 else ''} B={B}: apply {ms:.3f} ms  {amps*bpa/ms/1e6:.0f} GB/s | expval {ms2:.3f} ms "
              f"{amps*bpa/2/ms2/1e6:.0f} GB/s groups={op.plan_info()['n_x_groups']}")
    elif a.case in ("heis20", "tfim20"):
        # nearest-neighbour chain Hamiltonians on 20 qubits: x-masks of rank ~20 but weight <= 2
        n, B = 20, a.batch or 64
        strings, h = [], []
        for i in range(n - 1):
            for pp in (("XX", "YY", "ZZ") if a.case == "heis20" else ("ZZ",)):
                t = ["I"] * n
                t[i], t[i + 1] = pp[0], pp[1]
                strings.append("".join(t))
                h.append(1.0)
        if a.case == "tfim20":
            for i in range(n):
                t = ["I"] * n
                t[i] = "X"
                strings.append("".join(t))
                h.append(0.7)
        h = np.array(h, dtype=np.complex128)
        psi = ctx.uniform((1 << n, B), np.complex128)
        op = fp.PauliOp(h, strings, ctx=ctx)
        y = ctx.empty((1 << n, B), np.complex128)
        ev = ctx.empty((B,), np.complex128)
        plan = op._plan(np.complex128)
        l0 = ctx.launch_count
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        nl = (ctx.launch_count - l0) // (a.iters + 1)
        ms2 = timed(ctx, lambda: fp.lib.fp_op_expval(ctx._h, plan, vp(ev.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        amps = (1 << n) * B
        print(f"{a.case}: apply {ms:.3f} ms ({nl} launches) {amps*32/ms/1e6:.0f} GB/s | expval {ms2:.3f} ms "
              f"groups={op.plan_info()['n_x_groups']} strings={len(strings)}")
    elif a.case == "cfg3":
        n, B, S = 16, a.batch or 1024, 2000
        strings = random_strings(rng, n, S, max_weight=4)
        h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
        psi = ctx.uniform((1 << n, B), np.complex128)
        op = fp.PauliOp(h, strings, ctx=ctx)
        y = ctx.empty((1 << n, B), np.complex128)
        plan = op._plan(np.complex128)
        l0 = ctx.launch_count
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0), a.iters)
        print(f"cfg3: {ms:.3f} ms  launches/call={(ctx.launch_count-l0)//(a.iters+1)} groups={op.plan_info()['n_x_groups']}")
    elif a.case in ("cfg4w", "cfg4e"):
        n, B, S, K = 12, a.batch or 4096, 10000, 64
        strings = random_strings(rng, n, S)
        cdt = np.complex128 if a.c128 else np.complex64
        hk = (rng.uniform(-1, 1, (S, K)) + 1j * rng.uniform(-1, 1, (S, K))).astype(cdt)
        psi = ctx.uniform((1 << n, B), cdt)
        data = ctx.to_device(rng.random((K, B)).astype(np.float64 if a.c128 else np.float32))
        sop = fp.SummedPauliOp(strings, hk, ctx=ctx)
        plan = sop._plan(cdt)
        y = ctx.empty((1 << n, B), cdt)
        ev = ctx.empty((K, B), cdt)
        if a.case == "cfg4w":
            ms = timed(ctx, lambda: fp.lib.fp_sop_apply_weighted(ctx._h, plan, vp(y.ptr), vp(psi.ptr), vp(data.ptr),
                                                                 1 if a.c128 else 0, sz(1 << n), sz(B), 0), a.iters)
        else:
            ms = timed(ctx, lambda: fp.lib.fp_sop_expval(ctx._h, plan, vp(ev.ptr), vp(psi.ptr), sz(1 << n), sz(B), 0),
                       a.iters)
        print(f"{a.case}: {ms:.3f} ms")
    elif a.case == "e2e":
        # host-pointer paths of the single-string entry points: zero-copy vs staged, apply and expval separately
        import time

        n, B = 20, a.batch or 256
        dim = 1 << n
        ctx.set_async(False)
        string = random_strings(rng, n, 1)[0]
        codes, _ = fp._encode([string])
        coeff = np.array([0.75 - 0.5j])
        psi = ctx.uniform((dim, B), np.complex128)
        h_in = ctx.pinned_empty((dim, B), np.complex128)
        h_out = ctx.pinned_empty((dim, B), np.complex128)
        h_ev = ctx.pinned_empty((B,), np.complex128)
        fp.lib.fp_memcpy(ctx._h, vp(h_in.ctypes.data), vp(psi.ptr), sz(h_in.nbytes))
        d_out = ctx.empty((dim, B), np.complex128)

        def t(fn, reps=3):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) / reps * 1e3

        def apply(i, o):
            return lambda: fp.lib.fp_string_apply(ctx._h, 1, n, vp(codes.ctypes.data), vp(coeff.ctypes.data), vp(o),
                                                  vp(i), sz(dim), sz(B), 0)

        def expval(i):
            return lambda: fp.lib.fp_string_expval(ctx._h, 1, n, vp(codes.ctypes.data), vp(coeff.ctypes.data),
                                                   vp(h_ev.ctypes.data), vp(i), sz(dim), sz(B), 0)

        for zc in (1, 0):
            ctx.set_zero_copy(bool(zc))
            print(f"zero_copy={zc}: apply host->host {t(apply(h_in.ctypes.data, h_out.ctypes.data)):.1f} ms, "
                  f"apply host->dev {t(apply(h_in.ctypes.data, d_out.ptr)):.1f} ms, "
                  f"apply dev->host {t(apply(psi.ptr, h_out.ctypes.data)):.1f} ms, "
                  f"expval host {t(expval(h_in.ctypes.data)):.1f} ms")
    elif a.case == "b1":
        # one big state, batch 1 (the local piece of config 5): coalescing must come from the row index
        n, S = a.batch or 28, 8
        strings = random_strings(rng, n, S)
        h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
        psi = ctx.uniform((1 << n,), np.complex128)
        y = ctx.empty((1 << n,), np.complex128)
        op = fp.PauliOp(h, strings, ctx=ctx)
        plan = op._plan(np.complex128)
        ms = timed(ctx, lambda: fp.lib.fp_op_apply(ctx._h, plan, vp(y.ptr), vp(psi.ptr), sz(1 << n), sz(1), 0), a.iters)
        print(f"b1 n={n}: {ms:.3f} ms  {(1 << n) * 32 / ms / 1e6:.0f} GB/s algorithmic  kernels_used={ctx.coset_kernels_used()} "
              f"launches/call={ctx.launch_count // (a.iters + 1)}")
    elif a.case == "str20":
        n, B = 20, a.batch or 256
        psi = ctx.uniform((1 << n, B), np.complex128)
        ps = fp.PauliString(random_strings(rng, n, 1)[0], ctx=ctx)
        ms = timed(ctx, lambda: ps.apply(psi), a.iters)
        print(f"str20 apply: {ms:.3f} ms")
    ctx.sync()


if __name__ == "__main__":
    main()
