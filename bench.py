#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native fast-pauli hot path.

Metric (BASELINE.json): PauliOp.apply amplitude*strings/s and HBM GB/s (% of roofline).

Workload at every N (weak scaling, one rank per GPU, no data-path collective): the north-star shape
    PauliOp.apply (2-D overload, reference __pauli_op.hpp:399-468), 20 qubits, complex128, batch 64 per GPU,
on the SURVEY.md 8(d) operator pair:
    (i)  "few_group": 64 strings over 8 x-masks (8 z-variants each)  -> HBM-bound
    (ii) "random":    64 i.i.d. uniform IXYZ strings (64 x-masks)     -> FP64-bound (8*64 flops per amplitude)
One "step" = one apply of (i) + one apply of (ii) over the resident batch = 2 * 64 * dim * B amplitude*strings/GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras] [--no-cpu-baseline]

For N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ...
Rank 0 prints exactly one JSON line on stdout.  At N > 1 the line also carries `config3_strong` (BASELINE config 3:
one global batch of 1024 columns split over the ranks) and `config5` (one state sharded by its high index bits,
pairwise exchange through the C ABI's fp_comm_* entry points).

--impl reference times the reference's own OpenMP CPU implementation (oracle/_ref, compiled from the unmodified
reference headers; falls back to the plain-C port) on the host cores on the SAME config: every step is the full
20-qubit x 64-column step.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QUBITS = 20
BATCH = 64
N_STRINGS = 64
DTYPE = np.complex128
SEED = 18
STRING_SEED = 1234
METRIC = "PauliOp.apply amplitude*strings/s (20 qubits, complex128, batch 64/GPU; 64 strings over 8 x-masks + 64 random strings)"
UNIT = "amplitude*strings/s"
EMIT = print


def synth():
    """fast-pauli_b200/synth.py (numpy only) loaded WITHOUT importing the product package: the reference arm must not
    map the product library."""
    mod = sys.modules.get("_fp_synth_standalone")
    if mod is None:
        spec = importlib.util.spec_from_file_location("_fp_synth_standalone",
                                                      os.path.join(ROOT, "fast-pauli_b200", "synth.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_fp_synth_standalone"] = mod
        spec.loader.exec_module(mod)
    return mod


def headline_operators(n: int = N_QUBITS) -> dict:
    """The SURVEY 8(d) operator pair, deterministic: name -> (strings, coefficients)."""
    rng = np.random.default_rng(STRING_SEED)
    rs = synth().random_strings
    few = []
    for s in rs(rng, n, 8):  # 8 z-variants per x-mask: swapping X<->Y and I<->Z keeps the x-mask
        for _ in range(8):
            t = list(s)
            for q in range(n):
                if rng.random() < 0.5:
                    t[q] = {"X": "Y", "Y": "X", "I": "Z", "Z": "I"}[t[q]]
            few.append("".join(t))
    rand = rs(rng, n, N_STRINGS)
    ops = {}
    for name, strings in (("few_group", few), ("random", rand)):
        h = rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))
        ops[name] = (strings, h.astype(DTYPE))
    return ops


def config_dict(world: int) -> dict:
    dim = 1 << N_QUBITS
    return {"workload": "PauliOp.apply (2-D), 20 qubits, complex128, batch 64 per GPU: step = apply(64 strings over 8 "
                        "x-masks) + apply(64 random strings); batch-axis sharded (no collective)",
            "n_qubits": N_QUBITS, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "n_strings": N_STRINGS,
            "operators": ["few_group: 64 strings / 8 x-masks", "random: 64 i.i.d. IXYZ strings"],
            "state_bytes_per_gpu": dim * BATCH * 16,
            "l2": "inputs (1 GiB in + 1 GiB out per GPU and operator) are 8x larger than L2; no flush needed",
            "input_generator": f"counter-based splitmix64 U[0,1)+iU[0,1), seed {SEED}+rank; strings seed {STRING_SEED}"}


def masks_of(string: str) -> tuple[int, int, int]:
    """(x, z, nY) with string[q] <-> bit n-1-q (reference __pauli_string.hpp:52-54, 94-98)."""
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


def closed_form_rows(strings, coeffs, rows: np.ndarray, B: int, seed: int, first: int = 0,
                     cols: np.ndarray | None = None, ld: int | None = None) -> np.ndarray:
    """out[rows, cols] of PauliOp.apply on the counter-generated batch, from the closed form
    out(i,t) = sum_s h_s (-i)^nY_s (-1)^popc(i & z_s) psi(i ^ x_s, t); numpy only (no oracle, no product code)."""
    sy = synth()
    ld = B if ld is None else ld
    cols = np.arange(B, dtype=np.uint64) if cols is None else np.asarray(cols, dtype=np.uint64)
    rows = np.asarray(rows, dtype=np.uint64)
    out = np.zeros((len(rows), len(cols)), dtype=np.complex128)
    for s, h in zip(strings, coeffs):
        x, z, ny = masks_of(s)
        src = rows ^ np.uint64(x)
        par = np.array([bin(int(r) & z).count("1") & 1 for r in rows])
        m = (1 - 2 * par) * ((-1j) ** ny)
        idx = src[:, None] * np.uint64(ld) + cols[None, :] + np.uint64(first)
        out += (h * m)[:, None] * sy.uniform_complex_at(idx, np.complex128, seed)
    return out


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = f"/tmp/fp_clocks_{os.getpid()}.csv"

    def start(self) -> None:
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.fh.close()
            sm, smax, reasons = [], [], set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower() == "active":
                        reasons.add(name)
            if sm:
                out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
            os.unlink(self.path)
        except Exception as e:  # clocks are evidence, never a reason to lose the measurement
            out["error"] = str(e)
        return out


# ------------------------------------------------------------------------------------------- CPU reference arm
class CpuStep:
    """The full benchmark step on the reference CPU path: PauliOp::apply (std::execution::par overload,
    __pauli_op.hpp:413-468) of both headline operators on the same (dim, 64) host batch the GPU arm uses."""

    def __init__(self, backend):
        self.be = backend
        self.ops = headline_operators()
        dim = 1 << N_QUBITS
        self.psi = synth().uniform_host((dim, BATCH), DTYPE, seed=SEED)
        self.out = np.zeros_like(self.psi)
        self.work = 2.0 * N_STRINGS * dim * BATCH

    def run(self, which=("few_group", "random")) -> float:
        t0 = time.perf_counter()
        for name in which:
            strings, h = self.ops[name]
            self.out[...] = 0
            self.be.op_apply(strings, h, self.psi, out=self.out, par=True)
        return time.perf_counter() - t0


def cpu_sample_text(be) -> str:
    impl = ("unmodified reference headers compiled by oracle/Makefile (x86-64-v3: AVX2+FMA, no AVX-512, so the "
            "prebuilt .so runs on any host), std::execution::par") if be.kind == "reference" else "plain-C port, OpenMP"
    return (f"the full step: both operators on all {BATCH} columns of one rank's 20-qubit batch (the host has one "
            f"memory system whatever N is); {impl}")


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU arm; the others exit 0 without work
    from oracle import oracle as orc

    be = orc.reference() or orc.port()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every host core, up to 64: its par path allocates
    # n_threads x (dim x B) private copies of the 1 GiB batch (PO:427-429) and zero-fills / reduces them every call
    be.use_all_threads(cap=64)
    threads = be.max_threads()
    step = CpuStep(be)
    for _ in range(args.warmup):
        step.run()
    times = [step.run() for _ in range(args.steps)]
    total = sum(times)
    value = step.work * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": be.kind,
                         "sample": cpu_sample_text(be), "statistic": "mean over the timed steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    EMIT(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def timed_ms(fp, ctx, fn, iters: int, warmup: int = 1) -> float:
    import ctypes as C

    for _ in range(warmup):
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
    fp.lib.fp_event_destroy(e0)
    fp.lib.fp_event_destroy(e1)
    return ms.value / iters


def run_extras(fp, ctx, hbm_peak: float, fp64_tflops: float = 37.2) -> dict:
    """Device-resident timings of the other BASELINE configs (reported beside the headline, never as `value`)."""
    from fast_pauli_b200.synth import random_strings as rand_strings

    out = {}
    rng = np.random.default_rng(STRING_SEED)

    def guard(name, f):
        try:
            out[name] = f()
        except Exception as e:  # an extra must never cost the headline line
            out[name] = {"error": f"{type(e).__name__}: {e}"}

    # PauliOp.apply, 20 qubits, batch 64, complex128: (i) 64 strings over 8 x-masks (HBM-bound), (ii) 64 random strings
    def op20():
        n, B = 20, 64
        psi = ctx.uniform((1 << n, B), DTYPE, seed=SEED)
        res = {}
        xs = rand_strings(rng, n, 8)
        few = []
        for s in xs:  # 8 z-variants per x-mask: swap X<->Y and I<->Z at random positions keeps the x-mask
            for _ in range(8):
                t = list(s)
                for q in range(n):
                    if rng.random() < 0.5:
                        t[q] = {"X": "Y", "Y": "X", "I": "Z", "Z": "I"}[t[q]]
                few.append("".join(t))
        def variants(xmasks, per_mask):
            out_s = []
            for s in xmasks:
                for _ in range(per_mask):
                    t = list(s)
                    for q in range(n):
                        if rng.random() < 0.5:
                            t[q] = {"X": "Y", "Y": "X", "I": "Z", "Z": "I"}[t[q]]
                    out_s.append("".join(t))
            return out_s

        def local_dense(k):  # every Pauli string on k fixed qubits: 4^k strings over 2^k x-masks (GF(2) rank k)
            pos = sorted(int(p) for p in rng.choice(n, size=k, replace=False))
            out_s = []
            for idx in range(4**k):
                t = ["I"] * n
                for i, p_ in enumerate(pos):
                    t[p_] = "IXYZ"[(idx >> (2 * i)) & 3]
                out_s.append("".join(t))
            return out_s

        def chain(kind):  # nearest-neighbour chain Hamiltonians: low-weight x-masks of full rank (multi-pass plans)
            out_s = []
            for i in range(n - 1):
                for pp in (("XX", "YY", "ZZ") if kind == "heisenberg" else ("ZZ",)):
                    t = ["I"] * n
                    t[i], t[i + 1] = pp[0], pp[1]
                    out_s.append("".join(t))
            if kind == "tfim":
                for i in range(n):
                    t = ["I"] * n
                    t[i] = "X"
                    out_s.append("".join(t))
            return out_s

        cases = [("few_group_64_strings_8_xmasks", few), ("random_64_strings", rand_strings(rng, n, 64)),
                 ("heisenberg_chain_57_strings", chain("heisenberg")), ("tfim_chain_39_strings", chain("tfim")),
                 # the same 64 strings / 8 x-masks / 8 z-variants shape when the masks are closed under XOR (all
                 # Paulis on 3 qubits): rank 3, register-resident coset kernel
                 ("dense_3local_64_strings_8_xmasks", local_dense(3)),
                 ("dense_2local_16_strings_4_xmasks", local_dense(2)),
                 ("dense_4local_256_strings_16_xmasks", local_dense(4)),
                 # rank 4 / 5: FP64 tensor-core dense-coset kernel (apply); expectation_value stays on SIMT / K3b
                 ("dense_5local_1024_strings_32_xmasks", local_dense(5))]
        # HBM fraction as a function of the number of distinct x-masks (64 strings each time)
        for g in (1, 2, 4, 16):
            cases.append((f"sweep_64_strings_{g}_xmasks", variants(rand_strings(rng, n, g), 64 // g)))
        for tag, strings in cases:
            h = rng.uniform(-1, 1, len(strings)) + 1j * rng.uniform(-1, 1, len(strings))
            op = fp.PauliOp(h, strings, ctx=ctx)
            info = op.plan_info()
            y = op.apply(psi)
            ms = timed_ms(fp, ctx, lambda: fp.lib.fp_op_apply(ctx._h, op._plan(DTYPE), _vp(y.ptr), _vp(psi.ptr),
                                                                _sz(1 << n), _sz(B), 0), 5)
            amps = (1 << n) * B
            t_hbm = amps * 32 / (hbm_peak * 1e9)
            t_fp = 8.0 * info["n_x_groups"] * amps / (fp64_tflops * 1e12)
            res[tag] = {"ms": ms, "amp_strings_per_s": amps * len(strings) / (ms * 1e-3),
                        "algorithmic_GBps": amps * 32 / (ms * 1e-3) / 1e9,
                        "hbm_frac": amps * 32 / (ms * 1e-3) / 1e9 / hbm_peak, "x_groups": info["n_x_groups"],
                        # a call is quoted against the slower of its HBM and FP64 floors
                        "bound": "hbm" if t_hbm >= t_fp else "fp64", "fp64_frac": t_fp / (ms * 1e-3),
                        "frac_of_bound": max(t_hbm, t_fp) / (ms * 1e-3)}
            if tag.startswith("dense_") or tag.startswith("few_group") or "chain" in tag:
                ev = ctx.empty((B,), DTYPE)
                ms_e = timed_ms(fp, ctx, lambda: fp.lib.fp_op_expval(ctx._h, op._plan(DTYPE), _vp(ev.ptr), _vp(psi.ptr),
                                                                      _sz(1 << n), _sz(B), 0), 5)
                res[tag]["expectation_value_ms"] = ms_e
                res[tag]["expectation_value_hbm_frac"] = amps * 16 / (ms_e * 1e-3) / 1e9 / hbm_peak
            del op
        return res

    # config 3: PauliOp.apply (batch), 16 qubits, 2000 strings weight <= 4, batch 1024, complex128
    def cfg3():
        n, B, S = 16, 1024, 2000
        strings = rand_strings(rng, n, S, max_weight=4)
        h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
        psi = ctx.uniform((1 << n, B), DTYPE, seed=SEED)
        op = fp.PauliOp(h, strings, ctx=ctx)
        info = op.plan_info()
        y = op.apply(psi)
        ms = timed_ms(fp, ctx, lambda: fp.lib.fp_op_apply(ctx._h, op._plan(DTYPE), _vp(y.ptr), _vp(psi.ptr),
                                                            _sz(1 << n), _sz(B), 0), 3)
        amps = (1 << n) * B
        return {"ms": ms, "amp_strings_per_s": amps * S / (ms * 1e-3), "x_groups": info["n_x_groups"],
                "packed_strings": info["n_packed_strings"], "algorithmic_GBps": amps * 32 / (ms * 1e-3) / 1e9,
                "fp64_TFLOPs": 8.0 * info["n_x_groups"] * amps / (ms * 1e-3) / 1e12}

    # config 4: SummedPauliOp.apply_weighted + expectation_value, 12 qubits, 10k strings x 64 ops, batch 4096, complex64
    def cfg4():
        n, B, S, K = 12, 4096, 10000, 64
        strings = rand_strings(rng, n, S)
        hk = (rng.uniform(-1, 1, (S, K)) + 1j * rng.uniform(-1, 1, (S, K))).astype(np.complex64)
        psi = ctx.uniform((1 << n, B), np.complex64, seed=SEED)
        data = ctx.to_device(rng.random((K, B)).astype(np.float32))
        sop = fp.SummedPauliOp(strings, hk, ctx=ctx)
        plan = sop._plan(np.complex64)
        y = ctx.empty((1 << n, B), np.complex64)
        ev = ctx.empty((K, B), np.complex64)
        ms_w = timed_ms(fp, ctx, lambda: fp.lib.fp_sop_apply_weighted(ctx._h, plan, _vp(y.ptr), _vp(psi.ptr),
                                                                      _vp(data.ptr), 0, _sz(1 << n), _sz(B), 0), 2)
        ms_e = timed_ms(fp, ctx, lambda: fp.lib.fp_sop_expval(ctx._h, plan, _vp(ev.ptr), _vp(psi.ptr), _sz(1 << n),
                                                              _sz(B), 0), 2)
        amps = (1 << n) * B
        return {"apply_weighted_ms": ms_w, "expectation_value_ms": ms_e,
                "apply_weighted_amp_strings_per_s": amps * S / (ms_w * 1e-3),
                "expectation_value_amp_strings_per_s": amps * S / (ms_e * 1e-3)}

    # the shapes the reference itself publishes (docs/benchmark_results/qiskit_adv.csv, measured
    # through its Python bindings on an i9-13950HX, complex128).  Here: the same calls through this repo's Python
    # classes with HOST numpy arrays (staging included) and device-resident.  Different hardware: context only.
    # SummedPauliOp.square() at the size of the reference's examples/05_summed_pauli_op_sq.cpp: 12 qubits, all strings
    # of weight <= 2 (631), 1000 operators, complex64 -> 46666 output strings
    def square():
        import itertools as it
        import time as _t

        n, K = 12, 1000
        strings = ["I" * n]
        for w in (1, 2):
            combos = list(it.combinations(range(n), w))
            for word in it.product("XYZ", repeat=w):
                for combo in combos:
                    t = ["I"] * n
                    for p_, ch in zip(combo, word):
                        t[p_] = ch
                    strings.append("".join(t))
        sq_strings = list(strings)
        for w in (3, 4):
            combos = list(it.combinations(range(n), w))
            for word in it.product("XYZ", repeat=w):
                for combo in combos:
                    t = ["I"] * n
                    for p_, ch in zip(combo, word):
                        t[p_] = ch
                    sq_strings.append("".join(t))
        coeffs = (rng.uniform(-1, 1, (len(strings), K)) + 1j * rng.uniform(-1, 1, (len(strings), K))).astype(np.complex64)
        codes, _ = fp._encode(strings)
        sq_codes, _ = fp._encode(sq_strings)
        out = ctx.pinned_empty((len(sq_strings), K), np.complex64)

        def call():
            rc = fp.lib.fp_sop_square(ctx._h, fp.FP_C64, n, _sz(len(strings)), _vp(codes.ctypes.data), _sz(K),
                                      _vp(coeffs.ctypes.data), _sz(len(sq_strings)), _vp(sq_codes.ctypes.data),
                                      _vp(out.ctypes.data))
            if rc:
                raise RuntimeError(fp.lib.fp_last_error().decode())

        call()
        t0 = _t.perf_counter()
        for _ in range(3):
            call()
        ours = (_t.perf_counter() - t0) / 3
        res = {"n_strings": len(strings), "n_output_strings": len(sq_strings), "n_operators": K,
               "ours_host_to_host_ms": 1e3 * ours}
        try:
            from oracle import oracle as orc

            be = orc.reference()
            if be is not None:
                be.use_all_threads()
                t0 = _t.perf_counter()
                ref = orc.ref_sop_square(strings, coeffs)
                res["reference_ms"] = 1e3 * (_t.perf_counter() - t0)
                res["reference_cores"] = be.max_threads()
                scale = float(np.max(np.abs(ref[1])))
                res["max_rel_err_vs_reference"] = float(np.max(np.abs(np.asarray(out) - ref[1]))) / scale
        except Exception as e:  # the CPU comparison is optional
            res["reference_error"] = f"{type(e).__name__}: {e}"
        ctx.pinned_free(out)
        return res

    def published():
        import time as _t

        res = {}
        prng = np.random.default_rng(18)

        def run(tag, make, published_ms):
            obj, host_in, call = make()
            call(obj, host_in)  # builds plans, warms up
            t0 = _t.perf_counter()
            reps = 3
            for _ in range(reps):
                call(obj, host_in)
            host_ms = (_t.perf_counter() - t0) / reps * 1e3
            dev_in = ctx.to_device(host_in)
            call(obj, dev_in)
            ctx.sync()
            t0 = _t.perf_counter()
            for _ in range(reps):
                call(obj, dev_in)
            ctx.sync()
            dev_ms = (_t.perf_counter() - t0) / reps * 1e3
            res[tag] = {"published_cpu_ms": published_ms, "host_arrays_ms": host_ms, "device_resident_ms": dev_ms}

        def op_apply(nq, S):
            def make():
                strings = rand_strings(prng, nq, S)
                return (fp.PauliOp(np.ones(S), strings, ctx=ctx), prng.random(1 << nq).astype(np.complex128),
                        lambda o, x: o.apply(x))
            return make

        def op_expval(nq, S, B):
            def make():
                strings = rand_strings(prng, nq, S)
                return (fp.PauliOp(np.ones(S), strings, ctx=ctx), prng.random((1 << nq, B)).astype(np.complex128),
                        lambda o, x: o.expectation_value(x))
            return make

        def str_apply(nq):
            def make():
                return (fp.PauliString(rand_strings(prng, nq, 1)[0], ctx=ctx), prng.random(1 << nq).astype(np.complex128),
                        lambda o, x: o.apply(x))
            return make

        run("PauliString.apply_20q_1state", str_apply(20), 23.3)
        run("PauliString.apply_24q_1state", str_apply(24), 401.0)
        run("PauliOp.apply_16q_1000strings_1state", op_apply(16, 1000), 121.5)
        run("PauliOp.apply_18q_1000strings_1state", op_apply(18, 1000), 1024.0)
        run("PauliOp.expectation_value_12q_1024strings_1000states", op_expval(12, 1024, 1000), 1420.0)
        run("PauliOp.expectation_value_16q_1024strings_1000states", op_expval(16, 1024, 1000), 52500.0)
        return res

    # BASELINE config 1 (the reference's own CPU-runnable case): PauliOp.apply, 10 qubits, 64 random strings, batch 16,
    # complex128 -- host arrays in, host arrays out, timed beside the compiled reference on this box's cores
    def cfg1():
        import time as _t

        prng = np.random.default_rng(18)
        strings = rand_strings(prng, 10, 64)
        h = prng.uniform(-1, 1, 64) + 1j * prng.uniform(-1, 1, 64)
        psi = prng.random((1024, 16)) + 1j * prng.random((1024, 16))
        op = fp.PauliOp(h, strings, ctx=ctx)
        op.apply(psi)
        reps = 200
        t0 = _t.perf_counter()
        for _ in range(reps):
            y = op.apply(psi)
        ours_us = (_t.perf_counter() - t0) / reps * 1e6
        res = {"ours_host_arrays_us": ours_us, "amp_strings_per_s": 1024 * 16 * 64 / (ours_us * 1e-6)}
        try:
            from oracle import oracle as orc

            be = orc.reference() or orc.port()
            be.use_all_threads()
            out = np.zeros_like(psi)
            for par, tag in ((True, "reference_par_us"), (False, "reference_seq_us")):
                be.op_apply(strings, h, psi, out=out, par=par)
                t0 = _t.perf_counter()
                for _ in range(20):
                    be.op_apply(strings, h, psi, out=out, par=par)
                res[tag] = (_t.perf_counter() - t0) / 20 * 1e6
            res["reference_cores"] = be.max_threads()
            ref = np.zeros_like(psi)
            be.op_apply(strings, h, psi, out=ref, par=False)
            res["max_rel_err_vs_reference"] = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
        except Exception as e:
            res["reference_error"] = str(e)
        return res

    # BASELINE config 2 (round 1's headline, kept as a secondary key): PauliString.apply_batch + expectation_value,
    # 20 qubits, batch 256, complex128, device resident
    def cfg2():
        import ctypes as C

        n, B = 20, 256
        dim = 1 << n
        string = synth().random_strings(np.random.default_rng(STRING_SEED), n, 1)[0]
        if "X" not in string and "Y" not in string:
            string = "X" + string[1:]
        codes, _ = fp._encode([string])
        coeff = np.array([0.75 - 0.5j], dtype=DTYPE)
        psi = ctx.uniform((dim, B), DTYPE, seed=SEED)
        y = ctx.empty((dim, B), DTYPE)
        ev = ctx.empty((B,), DTYPE)
        ms_a = timed_ms(fp, ctx, lambda: fp.lib.fp_string_apply(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data),
                                                                _vp(coeff.ctypes.data), _vp(y.ptr), _vp(psi.ptr),
                                                                _sz(dim), _sz(B), 0), 10, warmup=3)
        ms_e = timed_ms(fp, ctx, lambda: fp.lib.fp_string_expval(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data),
                                                                 _vp(coeff.ctypes.data), _vp(ev.ptr), _vp(psi.ptr),
                                                                 _sz(dim), _sz(B), 0), 10, warmup=3)
        return {"string": string, "apply_batch_ms": ms_a, "apply_batch_hbm_frac": dim * B * 32 / (ms_a * 1e-3) / 1e9 / hbm_peak,
                "expectation_value_ms": ms_e, "expectation_value_hbm_frac": dim * B * 16 / (ms_e * 1e-3) / 1e9 / hbm_peak,
                "amp_strings_per_s": 2.0 * dim * B / ((ms_a + ms_e) * 1e-3)}

    ctx.set_async(True)
    guard("config2_pauli_string_apply_batch_expval_20q_b256_c128", cfg2)
    ctx.sync()
    ctx.set_async(False)
    guard("config1_pauli_op_apply_10q_64strings_b16_c128", cfg1)
    guard("reference_published_shapes", published)
    guard("summed_pauli_op_square_12q_weight2_1000ops_c64", square)
    def op20_b256():
        n, B = 20, 256
        strings, h = headline_operators(n)["few_group"]
        psi = ctx.uniform((1 << n, B), DTYPE, seed=SEED)
        op = fp.PauliOp(h, strings, ctx=ctx)
        y = op.apply(psi)
        ms = timed_ms(fp, ctx, lambda: fp.lib.fp_op_apply(ctx._h, op._plan(DTYPE), _vp(y.ptr), _vp(psi.ptr),
                                                            _sz(1 << n), _sz(B), 0), 5)
        amps = (1 << n) * B
        return {"ms": ms, "hbm_frac": amps * 32 / (ms * 1e-3) / 1e9 / hbm_peak,
                "amp_strings_per_s": amps * len(strings) / (ms * 1e-3)}

    ctx.set_async(True)
    guard("pauli_op_apply_20q_b64_c128", op20)
    guard("pauli_op_apply_20q_b256_c128_few_group", op20_b256)
    guard("config3_pauli_op_apply_16q_2000strings_b1024_c128", cfg3)
    guard("config4_summed_12q_10k_strings_64ops_b4096_c64", cfg4)
    ctx.sync()
    ctx.set_async(False)
    return out


def _vp(p):
    import ctypes as C

    return C.c_void_p(p)


def _sz(v):
    import ctypes as C

    return C.c_size_t(v)


def run_ours(args) -> None:
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    os.environ.setdefault("FASTPAULI_DEVICE", str(local_rank))

    dist = None
    torch = None
    if world > 1:
        # NCCL writes its banner / debug lines to stdout by default; keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", f"/tmp/fp_nccl_debug_{os.getpid()}.log")
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from __graft_entry__ import load_package

    fp = load_package()  # raises ImportError if the CUDA extension is missing: no fallback
    ctx = fp.Context(local_rank)
    # one process per GPU: keep the process and its pinned buffers on the GPU's NUMA node (e2e leg)
    numa_node = ctx.bind_host_to_gpu_numa() if world > 1 else None
    pk = peaks()

    n, B = N_QUBITS, BATCH
    dim = 1 << n
    ops = headline_operators(n)
    names = ("few_group", "random")
    pauli_ops = {k: fp.PauliOp(ops[k][1], ops[k][0], ctx=ctx) for k in names}
    plans = {k: pauli_ops[k]._plan(DTYPE) for k in names}
    infos = {k: pauli_ops[k].plan_info() for k in names}
    # batch shard of this rank: columns [rank*B, (rank+1)*B) of the global (dim, B*world) batch -- the generator is
    # counter based, so give every rank a distinct stream
    psi = ctx.uniform((dim, B), DTYPE, seed=SEED + rank)
    outs = {k: ctx.empty((dim, B), DTYPE) for k in names}
    ctx.set_async(True)

    def apply_call(k, dst_ptr=None, src_ptr=None):
        rc = fp.lib.fp_op_apply(ctx._h, plans[k], _vp(outs[k].ptr if dst_ptr is None else dst_ptr),
                                _vp(psi.ptr if src_ptr is None else src_ptr), _sz(dim), _sz(B), 0)
        if rc:
            raise RuntimeError(fp.lib.fp_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        if torch is not None:
            torch.cuda.synchronize()

    # ---- parity gate before anything is timed: sampled rows of both results against the closed form on regenerated
    # inputs (numpy, no oracle).  A number from a wrong result is not reported.
    parity = {}
    prng = np.random.default_rng(99 + rank)
    rows = np.unique(np.concatenate([prng.integers(0, dim, size=24), [0, dim - 1]])).astype(np.uint64)
    for k in names:
        l0 = ctx.launch_count
        apply_call(k)
        ctx.sync()
        launches_call = ctx.launch_count - l0
        got = np.stack([outs[k].get_rows(int(r), int(r) + 1)[0] for r in rows])
        exp = closed_form_rows(ops[k][0], ops[k][1], rows, B, SEED + rank)
        err = float(np.max(np.abs(got - exp)) / np.max(np.abs(exp)))
        parity[k] = {"max_rel_err": err, "rows_checked": int(len(rows)), "tol": 1e-12, "launches_per_call": int(launches_call)}
        if not err < 1e-12:
            raise SystemExit(f"parity gate failed for operator {k}: {err:.3e}")

    for _ in range(max(args.warmup, 3)):
        for k in names:
            apply_call(k)
    barrier()

    K = args.steps
    evs = []
    for _ in range(2 * K + 1):
        e = C.c_void_p()
        fp.lib.fp_event_create(C.byref(e))
        evs.append(e)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    l0 = ctx.launch_count
    fp.lib.fp_event_record(ctx._h, evs[0])
    for k_ in range(K):
        apply_call("few_group")
        fp.lib.fp_event_record(ctx._h, evs[2 * k_ + 1])
        apply_call("random")
        fp.lib.fp_event_record(ctx._h, evs[2 * k_ + 2])
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else {}

    def el(a, b) -> float:
        ms = C.c_float()
        fp.lib.fp_event_elapsed_ms(evs[a], evs[b], C.byref(ms))
        return float(ms.value)

    total_ms = el(0, 2 * K)
    call_ms = {"few_group": sum(el(2 * k_, 2 * k_ + 1) for k_ in range(K)) / K,
               "random": sum(el(2 * k_ + 1, 2 * k_ + 2) for k_ in range(K)) / K}
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * 2.0 * N_STRINGS * dim * B / (ms_per_step * 1e-3)

    # ---- roofline per operator call: algorithmic bytes = dim*B*(16 B read + 16 B written) whatever the number of
    # strings; companion compute figure 8 flops per amplitude per distinct x-mask; the bound of a call
    # is the slower of the two and frac is quoted against it.
    fp64 = fp64_peak(fp, ctx, clocks)
    alg_bytes = dim * B * 32.0
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "r02c_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic_db = json.load(open(tpath))
        except Exception:
            traffic_db = {}

    def roof(k):
        G = infos[k]["n_x_groups"]
        t = call_ms[k] * 1e-3
        t_hbm = alg_bytes / (pk["hbm_gbs"] * 1e9)
        flops = 8.0 * G * dim * B
        t_fp = flops / (fp64["tflops"] * 1e12)
        tr = traffic_db.get(k, {})
        r = {"operator": k, "x_groups": G, "strings": N_STRINGS, "avg_call_ms": call_ms[k],
             "launches_per_call": parity[k]["launches_per_call"],
             "algorithmic_bytes_per_call": alg_bytes, "hbm_GBps": alg_bytes / t / 1e9,
             "hbm_frac": t_hbm / t, "fp64_TFLOPs": flops / t / 1e12, "fp64_frac": t_fp / t,
             "t_hbm_ms": 1e3 * t_hbm, "t_fp64_ms": 1e3 * t_fp,
             "amp_strings_per_s": N_STRINGS * dim * B / t,
             "traffic": tr.get("dram_bytes_per_call"), "traffic_source": tr.get("source"),
             "kernel": tr.get("kernel", "coset_pair_tma_kernel / coset_dir_tma_kernel (K3j / K3i, csrc/coset4.cuh, coset3.cuh)")}
        # third floor, reported beside the two the SURVEY names: every complex FMA of a multi-mask pass needs one 16-byte
        # shared-memory gather (no register-level reuse between independent masks); 128 B/clk/SM of LDS bandwidth
        t_lds = 16.0 * G * dim * B / (148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6) if G > 4 else 0.0
        r["t_smem_gather_ms"] = 1e3 * t_lds
        r["frac_of_max_hbm_fp64_smem"] = max(t_hbm, t_fp, t_lds) / t
        if r["traffic"]:
            # what the implementation really moves through HBM (multi-pass plans re-stream the batch): how busy HBM is
            r["traffic_GBps"] = r["traffic"] / t / 1e9
            r["traffic_frac_of_hbm_peak"] = r["traffic"] / t / 1e9 / pk["hbm_gbs"]
        if t_hbm >= t_fp:
            r.update(bound="hbm", achieved=alg_bytes / t / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=t_hbm / t)
        else:
            r.update(bound="fp64", achieved=flops / t / 1e12, peak=fp64["tflops"], unit="TFLOP/s", frac=t_fp / t)
        return r

    per_op = {k: roof(k) for k in names}
    dominant = max(names, key=lambda k: call_ms[k])
    roofline = dict(per_op[dominant])
    roofline.update({"peak_source": pk["source"] + "; fp64: " + fp64["source"],
                     "note": "dominant = the operator call with the largest share of the step; frac is quoted against "
                             "max(t_HBM, t_FP64) of that call (SURVEY 8d); `by_operator` has both calls",
                     "by_operator": per_op})

    # ---- end to end through the public C ABI with pinned HOST buffers (H2D + kernels + D2H inside the call)
    ctx.set_async(False)
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        h_in = ctx.pinned_empty((dim, B), DTYPE)
        h_out = {k: ctx.pinned_empty((dim, B), DTYPE) for k in names}
        fp.lib.fp_memcpy(ctx._h, _vp(h_in.ctypes.data), _vp(psi.ptr), _sz(h_in.nbytes))

        def e2e_step():
            for k in names:
                apply_call(k, h_out[k].ctypes.data, h_in.ctypes.data)

        Ke = max(2, min(K, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * 2.0 * N_STRINGS * dim * B * Ke / dt, "unit": UNIT,
               "h2d_bytes_per_step": 2 * h_in.nbytes, "d2h_bytes_per_step": sum(h.nbytes for h in h_out.values()),
               "steps": Ke, "ms_per_step": 1e3 * dt / Ke, "host_numa_node_rank0": numa_node,
               # attribution of the N > 1 behaviour: what all ranks together pull through the host per second, beside
               # a plain single-thread host memcpy of the same pinned buffers on rank 0
               "aggregate_host_traffic_GBps": world * (2 * h_in.nbytes + sum(h.nbytes for h in h_out.values())) * Ke / dt / 1e9,

               "path": "fp_op_apply with pinned host pointers for input and output, once per operator: the call streams "
                       "256-byte column blocks through the GPU (strided upload of block j+1 | kernels on block j | "
                       "download of block j-1, three device blocks per direction)"}
        # keep the device result honest: the host copy of the output must equal the device-resident one
        for k in names:
            chk = outs[k].get_rows(12345, 12346)
            if not np.array_equal(chk, h_out[k][12345:12346]):
                e2e["warning"] = "host-staged result differs from device-resident result"
        e2e["host_memcpy_GBps_rank0"] = host_memcpy_gbps(h_out["few_group"], h_in)  # after the check: it overwrites
        ctx.pinned_free(h_in)
        for h in h_out.values():
            ctx.pinned_free(h)
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "error": f"{type(ex).__name__}: {ex}", "h2d_bytes_per_step": None,
               "d2h_bytes_per_step": None}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own par path on the same full step
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as orc

            be = orc.reference() or orc.port()
            be.use_all_threads(cap=64)
            step = CpuStep(be)
            step.run(("few_group",))  # warm-up: first touch of the n_threads x dim x B private copies (PO:427)
            dt = step.run()
            cpu_baseline = {"value": step.work / dt, "unit": UNIT, "cores": be.max_threads(), "kind": be.kind,
                            "sample": cpu_sample_text(be) + f"; one timed step ({dt:.1f} s) after one warm-up apply",
                            "statistic": "one step (the --impl reference arm reports the mean over its steps)"}
            del step
        except Exception as ex:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    config3_strong = None
    if not args.no_extras:
        try:
            config3_strong = run_config3_strong(fp, ctx, dist, torch, rank, world, fp64)
        except Exception as ex:
            config3_strong = {"error": f"{type(ex).__name__}: {ex}"}
    config5 = None
    if world > 1 and not args.no_extras:
        try:
            config5 = run_config5(fp, ctx, dist, torch, rank, world, local_rank)
        except Exception as ex:
            config5 = {"error": f"{type(ex).__name__}: {ex}"}

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        outs.clear()
        extras = run_extras(fp, ctx, pk["hbm_gbs"], fp64["tflops"])

    if dist is not None:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic", "config": config_dict(world),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity, "config3_strong": config3_strong, "config5": config5, "extras": extras,
        }
        EMIT(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def host_memcpy_gbps(dst: np.ndarray, src: np.ndarray) -> float:
    """Single-thread host memcpy bandwidth (read + write bytes per second) between two pinned buffers."""
    np.copyto(dst, src)
    t0 = time.perf_counter()
    np.copyto(dst, src)
    return 2.0 * src.nbytes / (time.perf_counter() - t0) / 1e9


def fp64_peak(fp, ctx, clocks: dict) -> dict:
    """FP64 FMA peak of this GPU: measured in this run by the library's DFMA microkernel when available, else
    148 SMs x 64 FMA/clk x the maximum SM clock (round 1 measured 63.7 FMA/clk/SM, scripts/micro/dmma_rate.cu)."""
    import ctypes as C

    try:
        val = C.c_double()
        if hasattr(fp.lib, "fp_measure_fp64_tflops") and fp.lib.fp_measure_fp64_tflops(ctx._h, C.byref(val)) == 0 \
                and val.value > 1.0:
            return {"tflops": float(val.value), "source": "measured in this run (fp_measure_fp64_tflops: dependent DFMA chains, all SMs)"}
    except Exception:
        pass
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    return {"tflops": 148 * 64 * 2 * mhz * 1e6 / 1e12, "source": f"nominal 148 SM x 64 FMA/clk x {mhz:.0f} MHz"}


def run_config3_strong(fp, ctx, dist, torch, rank: int, world: int, fp64: dict) -> dict:
    """BASELINE config 3 as STRONG scaling: PauliOp.apply, 16 qubits, 2000 strings of weight <= 4, ONE global batch
    of 1024 complex128 columns split over the ranks by column blocks (no collective on the data path).  Every rank
    regenerates its block of the global counter-based batch, applies the operator, and checks one sampled column of
    its block against the closed form; rank 0 also times the undivided problem in the same run."""
    import ctypes as C

    n, Bg, S = 16, 1024, 2000
    dim = 1 << n
    rng = np.random.default_rng(STRING_SEED + 3)
    strings = synth().random_strings(rng, n, S, max_weight=4)
    h = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
    op = fp.PauliOp(h, strings, ctx=ctx)
    plan = op._plan(DTYPE)
    info = op.plan_info()
    Bl = Bg // world
    c0 = rank * Bl
    # this rank's column block of the global (dim, Bg) batch: element (i, t) has flat index i*Bg + t
    idx = (np.arange(dim, dtype=np.uint64)[:, None] * np.uint64(Bg) + np.arange(c0, c0 + Bl, dtype=np.uint64)[None, :])
    host = synth().uniform_complex_at(idx, DTYPE, SEED)
    psi = ctx.to_device(np.ascontiguousarray(host))
    out = ctx.empty((dim, Bl), DTYPE)
    del idx

    def call(o, p, b):
        rc = fp.lib.fp_op_apply(ctx._h, plan, _vp(o.ptr), _vp(p.ptr), _sz(dim), _sz(b), 0)
        if rc:
            raise RuntimeError(fp.lib.fp_last_error().decode())

    ctx.set_async(True)
    call(out, psi, Bl)
    ctx.sync()
    # parity: one sampled column of the block, all rows, closed form in numpy
    tcol = int(np.random.default_rng(5 + rank).integers(0, Bl))
    got = out.get()[:, tcol]
    exp = np.zeros(dim, dtype=np.complex128)
    i = np.arange(dim, dtype=np.uint64)
    col = host[:, tcol]
    for s_, h_ in zip(strings, h):
        x, z, ny = masks_of(s_)
        zz = i & np.uint64(z)
        par = np.zeros(dim, dtype=np.int64)
        while zz.any():
            par ^= (zz & np.uint64(1)).astype(np.int64)
            zz >>= np.uint64(1)
        exp += (h_ * (-1j) ** ny) * (1 - 2 * par) * col[(i ^ np.uint64(x)).astype(np.int64)]
    err = float(np.max(np.abs(got - exp)) / np.max(np.abs(exp)))
    del host
    iters = 5
    ms = timed_ms(fp, ctx, lambda: call(out, psi, Bl), iters, warmup=2)
    ms_all = ms
    err_all = err
    if dist is not None:
        t = torch.tensor([ms, err], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_all, err_all = float(t[0].item()), float(t[1].item())
    res = {"workload": "PauliOp.apply, 16 qubits, 2000 strings weight<=4, global batch 1024 complex128, column blocks",
           "n_gpus": world, "batch_per_gpu": Bl, "ms": ms_all, "x_groups": info["n_x_groups"],
           "amp_strings_per_s": dim * Bg * S / (ms_all * 1e-3), "parity_max_rel_err": err_all, "parity_tol": 1e-12,
           "parity_check": "one sampled column per rank, all rows, closed form",
           "fp64_TFLOPs_per_gpu": 8.0 * info["n_x_groups"] * dim * Bl / (ms_all * 1e-3) / 1e12,
           "fp64_frac_per_gpu": 8.0 * info["n_x_groups"] * dim * Bl / (ms_all * 1e-3) / 1e12 / fp64["tflops"],
           "scaling": "strong", "timing": "CUDA events per rank, max over ranks"}
    if not err_all < 1e-12:
        res["error"] = "parity gate failed"
    del psi, out
    if rank == 0 and world > 1:
        # the undivided problem on one GPU in the same run: the denominator of the speed-up
        psi1 = ctx.uniform((dim, Bg), DTYPE, seed=SEED)
        out1 = ctx.empty((dim, Bg), DTYPE)
        ms1 = timed_ms(fp, ctx, lambda: call(out1, psi1, Bg), 3, warmup=2)
        res["ms_1gpu_same_run"] = ms1
        res["speedup_vs_1gpu"] = ms1 / ms_all
        del psi1, out1
    ctx.sync()
    ctx.set_async(False)
    if dist is not None:
        dist.barrier()
    return res


def run_config5(fp, ctx, dist, torch, rank: int, world: int, local_rank: int) -> dict:
    """BASELINE config 5 (scaled to the number of ranks): one complex128 state sharded by its high index bits,
    PauliOp.apply with pairwise amplitude exchange through the C ABI (fp_comm_* / fp_sharded_op_apply, NCCL linked by
    the library; torch only hands out the ncclUniqueId bytes)."""
    from fast_pauli_b200 import sharded

    return sharded.bench_config5(fp, ctx, dist, torch, rank, world, local_rank, seed=SEED, string_seed=STRING_SEED)


def _claim_stdout():
    """Keep the real stdout for the single JSON line: everything else written to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr.  Returns a writer for the JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(text: str) -> None:
        os.write(real, (text + "\n").encode())

    return emit


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (ncu launch lists)")
    args = ap.parse_args()
    global EMIT
    EMIT = _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
