#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native fast-pauli hot path.

Metric (BASELINE.json): PauliOp.apply amplitude*strings/s and HBM GB/s (% of roofline).

Workload at every N (weak scaling, one rank per GPU, no data-path collective): BASELINE config 2,
    PauliString.apply_batch + PauliString.expectation_value, 20 qubits, batch 256 per GPU, complex128
(4 GiB in + 4 GiB out per GPU; a PauliString is the one-string PauliOp, and both calls run the same
kernels PauliOp.apply / expectation_value use).  One "step" = one apply_batch + one expectation_value
over the whole resident batch = 2 * dim * n_states amplitude*strings per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras] [--no-cpu-baseline]

For N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ...
Rank 0 prints exactly one JSON line on stdout.

--impl reference times the reference's own OpenMP CPU implementation (oracle/_ref, compiled from the unmodified
reference headers; falls back to the plain-C port) on the host cores for the same metric/config, each step a
bounded column sample of the workload.
"""
from __future__ import annotations

import argparse
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
BATCH = 256
DTYPE = np.complex128
SEED = 18
STRING_SEED = 1234
METRIC = "PauliOp.apply amplitude*strings/s (PauliString.apply_batch + expectation_value, 20 qubits, batch 256/GPU, complex128)"
UNIT = "amplitude*strings/s"
EMIT = print


def make_string(n: int, seed: int = STRING_SEED) -> str:
    """One i.i.d. uniform IXYZ string (tests/benchmarks/test_qiskit_adv.py:122-125 style), fixed seed."""
    rng = np.random.default_rng(seed)
    s = "".join(np.array(list("IXYZ"))[rng.integers(0, 4, size=n)])
    if "X" not in s and "Y" not in s:  # keep the gather non-trivial
        s = "X" + s[1:]
    return s


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
def cpu_reference_step(backend, string: str, n: int, cols: int, par: bool = True) -> tuple[float, float]:
    """One bounded step of the reference CPU path: apply_batch + expectation_value on `cols` columns.
    Returns (seconds, amplitude*strings processed)."""
    from __graft_entry__ import load_package

    load_package()  # only for the host twin of the input generator; no GPU work on this arm
    from fast_pauli_b200.synth import uniform_host

    dim = 1 << n
    psi = getattr(cpu_reference_step, "_psi", None)
    if psi is None or psi.shape != (dim, cols):
        psi = uniform_host((dim, cols), DTYPE, seed=SEED)
        cpu_reference_step._psi = psi
        cpu_reference_step._out = np.zeros_like(psi)
        cpu_reference_step._ev = np.zeros(cols, dtype=DTYPE)
    out, ev = cpu_reference_step._out, cpu_reference_step._ev
    t0 = time.perf_counter()
    backend.string_apply(string, psi, 0.75 - 0.5j, out=out, par=par)
    backend.string_expval(string, psi, 0.75 - 0.5j, out=ev, par=par)
    dt = time.perf_counter() - t0
    return dt, 2.0 * dim * cols


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU arm; the others exit 0 without work
    from oracle import oracle as orc

    be = orc.reference() or orc.port()
    be.use_all_threads()  # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every host core
    string = make_string(N_QUBITS)
    cols = 16  # bounded sample: 16 of the 256 columns per step (work is exactly linear in the batch)
    threads = be.max_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_step(be, string, N_QUBITS, cols)
    times, work = [], 0.0
    for _ in range(args.steps):
        dt, w = cpu_reference_step(be, string, N_QUBITS, cols)
        times.append(dt)
        work += w
    total = sum(times)
    value = work / total
    sample = (f"{cols} of {BATCH} batch columns per step (work is linear in the batch); "
              f"{'unmodified reference headers, std::execution::par' if be.kind == 'reference' else 'plain-C port, OpenMP'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: PauliString.apply_batch + expectation_value, 20 qubits, complex128",
                   "n_qubits": N_QUBITS, "batch_per_step": cols, "string": string},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": be.kind, "sample": sample},
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


def run_extras(fp, ctx, hbm_peak: float) -> dict:
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
            res[tag] = {"ms": ms, "amp_strings_per_s": amps * len(strings) / (ms * 1e-3),
                        "algorithmic_GBps": amps * 32 / (ms * 1e-3) / 1e9,
                        "hbm_frac": amps * 32 / (ms * 1e-3) / 1e9 / hbm_peak, "x_groups": info["n_x_groups"]}
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

    ctx.set_async(False)
    guard("config1_pauli_op_apply_10q_64strings_b16_c128", cfg1)
    guard("reference_published_shapes", published)
    guard("summed_pauli_op_square_12q_weight2_1000ops_c64", square)
    ctx.set_async(True)
    guard("pauli_op_apply_20q_b64_c128", op20)
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
    string = make_string(n)
    coeff = np.array([0.75 - 0.5j], dtype=DTYPE)
    codes, _ = fp._encode([string])
    # batch shard of this rank: columns [rank*B, (rank+1)*B) of the global (dim, B*world) batch -- the generator is
    # counter based, so give every rank a distinct stream offset
    psi = ctx.uniform((dim, B), DTYPE, seed=SEED + rank)
    out = ctx.empty((dim, B), DTYPE)
    ev = ctx.empty((B,), DTYPE)
    ctx.set_async(True)

    def apply_call():
        rc = fp.lib.fp_string_apply(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data), _vp(coeff.ctypes.data), _vp(out.ptr),
                                    _vp(psi.ptr), _sz(dim), _sz(B), 0)
        if rc:
            raise RuntimeError(fp.lib.fp_last_error().decode())

    def expval_call():
        rc = fp.lib.fp_string_expval(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data), _vp(coeff.ctypes.data), _vp(ev.ptr),
                                     _vp(psi.ptr), _sz(dim), _sz(B), 0)
        if rc:
            raise RuntimeError(fp.lib.fp_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        if torch is not None:
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        apply_call()
        expval_call()
    barrier()

    K = args.steps
    evs = []
    for _ in range(3 * K + 1):
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
    for k in range(K):
        apply_call()
        fp.lib.fp_event_record(ctx._h, evs[3 * k + 1])
        expval_call()
        fp.lib.fp_event_record(ctx._h, evs[3 * k + 2])
    fp.lib.fp_event_record(ctx._h, evs[3 * K])
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else {}

    def el(a, b) -> float:
        ms = C.c_float()
        fp.lib.fp_event_elapsed_ms(evs[a], evs[b], C.byref(ms))
        return float(ms.value)

    total_ms = el(0, 3 * K)
    apply_ms = [el(3 * k if k == 0 else 3 * k - 1, 3 * k + 1) for k in range(K)]
    expval_ms = [el(3 * k + 1, 3 * k + 2) for k in range(K)]
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * 2.0 * dim * B / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (the streaming apply): algorithmic bytes = dim*B*16 B read + 16 B written
    apply_avg = sum(apply_ms) / K
    expval_avg = sum(expval_ms) / K
    alg_bytes = dim * B * 32.0
    achieved = alg_bytes / (apply_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("string_apply_c128_20q_b256_bytes")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "op_kernel<double,EPV=1,V=4,MODE=0,INLINE1> (PauliString.apply_batch)",
                "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "traffic": traffic, "peak_source": pk["source"], "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": apply_avg,
                "note": "peak is the driver-measured torch copy bandwidth (b.copy_(a)); frac > 1 means this kernel moves "
                        "its compulsory bytes faster than that copy (nominal HBM3e: 8000 GB/s -> frac_nominal below)",
                "frac_nominal_8TBps": alg_bytes / (apply_avg * 1e-3) / 1e9 / 8000.0,
                "second_kernel": {"kernel": "expval_pairs_kernel<double,1,4,1> (PauliString.expectation_value)",
                                  "algorithmic_bytes_per_launch": dim * B * 16.0, "avg_launch_ms": expval_avg,
                                  "achieved": dim * B * 16.0 / (expval_avg * 1e-3) / 1e9,
                                  "frac": dim * B * 16.0 / (expval_avg * 1e-3) / 1e9 / pk["hbm_gbs"]}}

    # ---- end to end through the public C ABI with pinned HOST buffers (H2D + kernel + D2H inside the call)
    ctx.set_async(False)
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        h_in = ctx.pinned_empty((dim, B), DTYPE)
        h_out = ctx.pinned_empty((dim, B), DTYPE)
        h_ev = ctx.pinned_empty((B,), DTYPE)
        fp.lib.fp_memcpy(ctx._h, _vp(h_in.ctypes.data), _vp(psi.ptr), _sz(h_in.nbytes))

        def e2e_step():
            rc = fp.lib.fp_string_apply(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data), _vp(coeff.ctypes.data),
                                        _vp(h_out.ctypes.data), _vp(h_in.ctypes.data), _sz(dim), _sz(B), 0)
            rc |= fp.lib.fp_string_expval(ctx._h, fp.FP_C128, n, _vp(codes.ctypes.data), _vp(coeff.ctypes.data),
                                          _vp(h_ev.ctypes.data), _vp(h_in.ctypes.data), _sz(dim), _sz(B), 0)
            if rc:
                raise RuntimeError(fp.lib.fp_last_error().decode())

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
        e2e = {"value": world * 2.0 * dim * B * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * h_in.nbytes,
               "d2h_bytes_per_step": h_out.nbytes + h_ev.nbytes, "steps": Ke, "ms_per_step": 1e3 * dt / Ke,
               "host_numa_node_rank0": numa_node,
               "path": "fp_string_apply + fp_string_expval with pinned host pointers: apply streams the batch through "
                       "the GPU in 32 MiB row blocks (upload j+1 | kernel j | download j-1 on two copy engines), "
                       "expectation_value uploads with the copy engine and reduces on the device"}

        def timed_variant() -> float:
            e2e_step()
            barrier()
            t1 = time.perf_counter()
            for _ in range(2):
                e2e_step()
            barrier()
            return 1e3 * (time.perf_counter() - t1) / 2

        # the same calls without the chunk pipeline, for reference: (a) kernels reading / writing the pinned buffers in
        # place over PCIe, (b) one-shot staging (cudaMemcpyAsync H2D -> kernel -> D2H)
        ctx.set_pipeline(False)
        e2e["zero_copy_ms_per_step"] = timed_variant()
        ctx.set_zero_copy(False)
        e2e["staged_copy_ms_per_step"] = timed_variant()
        ctx.set_zero_copy(True)
        ctx.set_pipeline(True)
        # keep the device result honest: the host copy of the output must equal the device-resident one
        chk = out.get_rows(12345, 12346)
        if not np.array_equal(chk, h_out[12345:12346]):
            e2e["warning"] = "host-staged result differs from device-resident result"
        ctx.pinned_free(h_in)
        ctx.pinned_free(h_out)
        ctx.pinned_free(h_ev)
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "error": f"{type(ex).__name__}: {ex}", "h2d_bytes_per_step": None,
               "d2h_bytes_per_step": None}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as orc

            be = orc.reference() or orc.port()
            be.use_all_threads()
            cols = 16
            cpu_reference_step(be, string, n, cols)  # warm-up
            best, spent, reps = None, 0.0, 0
            while spent < 10.0 and reps < 2000:
                dt, w = cpu_reference_step(be, string, n, cols)
                spent += dt
                reps += 1
                best = dt if best is None else min(best, dt)
            cpu_baseline = {"value": 2.0 * dim * cols / best, "unit": UNIT, "cores": be.max_threads(), "kind": be.kind,
                            "sample": f"{cols} of {B} batch columns, best of {spent:.1f} s of repeats; "
                                      f"{'unmodified reference headers (std::execution::par)' if be.kind == 'reference' else 'plain-C port + OpenMP'}"}
        except Exception as ex:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        del out
        extras = run_extras(fp, ctx, pk["hbm_gbs"])

    if dist is not None:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: PauliString.apply_batch + PauliString.expectation_value, "
                                   "20 qubits, batch 256 per GPU, complex128, batch-axis sharded (no collective)",
                       "n_qubits": n, "batch_per_gpu": B, "global_batch": B * world, "string": string,
                       "state_bytes_per_gpu": dim * B * 16,
                       "l2": "inputs (4 GiB per GPU) are 32x larger than L2; no flush needed",
                       "input_generator": f"counter-based splitmix64 U[0,1)+iU[0,1), seed {SEED}+rank"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "extras": extras,
        }
        EMIT(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
