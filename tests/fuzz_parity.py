#!/usr/bin/env python
"""Randomised dispatch fuzzer (GPU): random register sizes, batch widths, operator structures, dtypes and array
residency through every hot-path entry point of the Python front-end, each result checked against the CPU oracle.

    python tests/fuzz_parity.py --seconds 120 --seed 1        # prints one line per failure + a summary

The operator families are chosen to land on every kernel-selection branch (single mask, few masks x many z, k-local
dense = register / tensor-core cosets, low weight = shared-memory cosets, i.i.d. = generic gather, chains, diagonal
only, duplicates).  tests/test_gpu_parity.py::test_fuzz_fixed_seeds replays a bounded slice of it.
"""
from __future__ import annotations

import argparse
import itertools as it
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LETTERS = np.array(list("IXYZ"))
LAST_NOTES: list[str] = []  # arbiter decisions of the last run() (counted by the pytest audit summary)


def make_strings(rng: np.random.Generator, n: int) -> tuple[str, list[str]]:
    fam = rng.choice(["iid", "lowweight", "klocal", "fewmasks", "onemask", "diag", "chain", "dups", "single"])
    if fam == "iid":
        S = int(rng.integers(1, 80))
        s = ["".join(LETTERS[rng.integers(0, 4, size=n)]) for _ in range(S)]
    elif fam == "lowweight":
        S = int(rng.integers(1, 200))
        s = []
        for _ in range(S):
            w = int(rng.integers(1, min(4, n) + 1))
            pos = rng.choice(n, size=w, replace=False)
            st = ["I"] * n
            for p in pos:
                st[p] = "XYZ"[int(rng.integers(0, 3))]
            s.append("".join(st))
    elif fam == "klocal":
        k = int(rng.integers(1, min(5, n) + 1))
        pos = sorted(rng.choice(n, size=k, replace=False))
        s = []
        for combo in it.product("IXYZ", repeat=k):
            st = ["I"] * n
            for p, ch in zip(pos, combo):
                st[p] = ch
            s.append("".join(st))
        if rng.random() < 0.5:  # a random subset of the dense term
            keep = rng.random(len(s)) < 0.6
            s = [x for x, kf in zip(s, keep) if kf] or s[:1]
    elif fam in ("fewmasks", "onemask"):
        G = 1 if fam == "onemask" else int(rng.integers(2, 10))
        s = []
        for _ in range(G):
            x = rng.integers(0, 2, size=n)
            for _ in range(int(rng.integers(1, 9))):
                z = rng.integers(0, 2, size=n)
                s.append("".join("IXZY"[a + 2 * b] for a, b in zip(x, z)))
    elif fam == "diag":
        S = int(rng.integers(1, 40))
        s = ["".join("IZ"[b] for b in rng.integers(0, 2, size=n)) for _ in range(S)]
    elif fam == "chain":
        s = []
        for q in range(n - 1):
            for ch in "XYZ":
                st = ["I"] * n
                st[q] = st[q + 1] = ch
                s.append("".join(st))
        s = s or ["Z" * n]
    elif fam == "dups":
        base = "".join(LETTERS[rng.integers(0, 4, size=n)])
        s = [base] * int(rng.integers(2, 20)) + ["I" * n] * int(rng.integers(0, 3))
    else:
        s = ["".join(LETTERS[rng.integers(0, 4, size=n)])]
    return str(fam), s


def exact_string_expvals(strings: list[str], psi: np.ndarray, with_abs: bool = False):
    """E(s, t) = <psi_t| P_s |psi_t> in 80-bit extended precision (numpy clongdouble), the arbiter when the GPU and
    the oracle disagree on a long, cancelling reduction: the reference sums sequentially (PS:534), so over 2^21 terms
    its OWN rounding error can exceed 1e-12 of the result."""
    n = len(strings[0])
    i = np.arange(1 << n, dtype=np.int64)
    ph = psi.astype(np.clongdouble)
    out = np.zeros((len(strings), psi.shape[1]), dtype=np.clongdouble)
    absum = np.zeros((len(strings), psi.shape[1]), dtype=np.longdouble)
    for k, st in enumerate(strings):
        x = z = 0
        for q, ch in enumerate(st):
            bit = 1 << (n - 1 - q)
            x |= bit if ch in "XY" else 0
            z |= bit if ch in "YZ" else 0
        par = np.zeros(1 << n, dtype=np.int64)
        zz = i & z
        while zz.any():
            par ^= zz & 1
            zz >>= 1
        m = (np.array([1, -1j, -1, 1j])[st.count("Y") & 3] * (1 - 2 * par)).astype(np.clongdouble)
        terms = np.conj(ph) * (m[:, None] * ph[i ^ x])
        out[k] = terms.sum(0)
        if with_abs:
            absum[k] = np.abs(terms).sum(0)
    return (out, absum) if with_abs else out


def gen_cases(seed: int, n_max: int = 14, log2_elems: int = 21, s_cap: int = 10**9):
    """The deterministic case stream of one seed (no GPU needed: failures can be replayed anywhere)."""
    rng = np.random.default_rng(seed)
    case = 0
    while True:
        case += 1
        n = int(rng.integers(1, n_max + 1))
        dtype = np.complex128 if rng.random() < 0.5 else np.complex64
        bmax = max(1, min(300, (1 << log2_elems) >> n))
        B = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 12, 16, 31, 33, 64, 100, 128, 257]))
        B = min(B, bmax)
        fam, strings = make_strings(rng, n)
        strings = strings[:s_cap]  # large registers: bound the single-threaded oracle's work
        S = len(strings)
        K = int(rng.integers(1, 6))
        on_dev = rng.random() < 0.4
        rdt = np.float64 if dtype == np.complex128 else np.float32
        psi = (rng.random((1 << n, B)) + 1j * rng.random((1 << n, B))).astype(dtype)
        h = (rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)).astype(dtype)
        hk = (rng.uniform(-1, 1, (S, K)) + 1j * rng.uniform(-1, 1, (S, K))).astype(dtype)
        data = rng.random((K, B)).astype(rdt)
        tag = f"seed={seed} case={case} fam={fam} n={n} S={S} B={B} K={K} dtype={np.dtype(dtype).name} dev={on_dev}"
        yield dict(n=n, dtype=dtype, B=B, fam=fam, strings=strings, S=S, K=K, on_dev=on_dev, psi=psi, h=h, hk=hk,
                   data=data, tag=tag, case=case, accumulate=(case % 2 == 0))


def load_native():
    """The pybind11 front-end (fast-pauli_b200/_fast_pauli*.so), or None when it is not built."""
    import glob
    import importlib.util

    if "_fast_pauli" in sys.modules:
        return sys.modules["_fast_pauli"]
    hits = glob.glob(os.path.join(ROOT, "fast-pauli_b200", "_fast_pauli*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("_fast_pauli", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["_fast_pauli"] = mod
    return mod


def run(seconds: float, seed: int, max_cases: int | None = None, verbose: bool = False, n_max: int = 14,
        log2_elems: int = 21, s_cap: int = 10**9, native: bool = False) -> tuple[int, list[str]]:
    from __graft_entry__ import load_package
    from oracle import oracle as orc

    fp = load_package()
    ORC = orc.port()
    # the oracle-shaped one-shot entry points of the GPU library: `out=` is accumulated into, like the C++ methods
    G = orc.Backend(os.path.join(ROOT, "fast-pauli_b200", "lib", "libfastpauli_b200.so"), "fp_", "gpu")
    ctx = fp.default_context()
    nf = load_native() if native else None
    t_end = time.time() + seconds
    failures: list[str] = []
    notes: list[str] = []  # disagreements settled in the GPU's favour by the extended-precision arbiter
    cases = 0

    def rel(a, b):
        a = a.get() if hasattr(a, "get") else np.asarray(a)
        scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
        return float(np.max(np.abs(a - b))) / scale if b.size else 0.0

    for case in gen_cases(seed, n_max, log2_elems, s_cap):
        if time.time() >= t_end or (max_cases is not None and cases >= max_cases):
            break
        cases += 1
        n, dtype, B, fam, strings, S, K, on_dev = (case[k] for k in ("n", "dtype", "B", "fam", "strings", "S", "K", "on_dev"))
        psi, h, hk, data, tag = case["psi"], case["h"], case["hk"], case["data"], case["tag"]
        tol = 1e-12 if dtype == np.complex128 else 2e-5
        arg = ctx.to_device(psi) if on_dev else psi
        darg = ctx.to_device(data) if on_dev else data
        # complex64 expectation values: compare with the complex128 oracle (the reference's own float32 running sums
        # lose more than the tolerance over long reductions, see tests/test_gpu_parity.assert_parity)
        psi_hi, h_hi, hk_hi = psi.astype(np.complex128), h.astype(np.complex128), hk.astype(np.complex128)
        try:
            checks = {
                "string.apply": (fp.PauliString(strings[0]).apply(arg, 0.5 - 2j),
                                 ORC.string_apply(strings[0], psi, 0.5 - 2j)),
                "string.expval": (fp.PauliString(strings[0]).expectation_value(arg, 0.5 - 2j),
                                  ORC.string_expval(strings[0], psi_hi, 0.5 - 2j)),
                "op.apply": (fp.PauliOp(h, strings).apply(arg), ORC.op_apply(strings, h, psi)),
                "op.expval": (fp.PauliOp(h, strings).expectation_value(arg), ORC.op_expval(strings, h_hi, psi_hi)),
                "sop.apply": (fp.SummedPauliOp(strings, hk).apply(arg), ORC.sop_apply(strings, hk, psi)),
                "sop.apply_weighted": (fp.SummedPauliOp(strings, hk).apply_weighted(arg, darg),
                                       ORC.sop_apply_weighted(strings, hk, psi, data)),
                "sop.expval": (fp.SummedPauliOp(strings, hk).expectation_value(arg),
                               ORC.sop_expval(strings, hk_hi, psi_hi)),
            }
            if nf is not None and dtype == np.complex128:
                # the same seven calls through the pybind11 module over the C++ classes (host complex128 arrays)
                nop, nsop, nps = nf.PauliOp(h, strings), nf.SummedPauliOp(strings, hk), nf.PauliString(strings[0])
                checks.update({
                    "native string.apply": (nps.apply(psi, 0.5 - 2j), checks["string.apply"][1]),
                    "native string.expval": (nps.expectation_value(psi, 0.5 - 2j), checks["string.expval"][1]),
                    "native op.apply": (nop.apply(psi), checks["op.apply"][1]),
                    "native op.expval": (nop.expectation_value(psi), checks["op.expval"][1]),
                    "native sop.apply": (nsop.apply(psi), checks["sop.apply"][1]),
                    "native sop.apply_weighted": (nsop.apply_weighted(psi, data), checks["sop.apply_weighted"][1]),
                    "native sop.expval": (nsop.expectation_value(psi), checks["sop.expval"][1]),
                })
            if case["accumulate"] and dtype == np.complex128:
                # C++ semantics (+= into the caller's buffer, PS:419,432,523,534; PO:453,465; SPO:331,346,464,498,
                # 588,609) through the raw C ABI; complex128 only so the bases do not mask float32 rounding
                r2 = np.random.default_rng(case["case"])
                base = (r2.random(psi.shape) + 1j * r2.random(psi.shape)).astype(dtype)
                e0 = (r2.random(B) + 1j * r2.random(B)).astype(dtype)
                ek0 = (r2.random((K, B)) + 1j * r2.random((K, B))).astype(dtype)
                checks.update({
                    "acc string.apply": (G.string_apply(strings[0], psi, 0.5 - 2j, out=base.copy()),
                                         ORC.string_apply(strings[0], psi, 0.5 - 2j, out=base.copy())),
                    "acc op.apply": (G.op_apply(strings, h, psi, out=base.copy()),
                                     ORC.op_apply(strings, h, psi, out=base.copy())),
                    "acc op.expval": (G.op_expval(strings, h, psi, out=e0.copy()),
                                      ORC.op_expval(strings, h, psi, out=e0.copy())),
                    "acc sop.apply": (G.sop_apply(strings, hk, psi, out=base.copy()),
                                      ORC.sop_apply(strings, hk, psi, out=base.copy())),
                    "acc sop.apply_weighted": (G.sop_apply_weighted(strings, hk, psi, data, out=base.copy()),
                                               ORC.sop_apply_weighted(strings, hk, psi, data, out=base.copy())),
                    "acc sop.expval": (G.sop_expval(strings, hk, psi, out=ek0.copy()),
                                       ORC.sop_expval(strings, hk, psi, out=ek0.copy())),
                })
            for name, (got, want) in checks.items():
                e = rel(got, want)
                if e < tol:
                    continue
                if name.endswith("expval") and not name.startswith("acc"):
                    name_key = name.replace("native ", "")
                    # arbitrate in extended precision: the GPU must be within tolerance of the exact value and at
                    # least as close to it as the reference-order sum is
                    used = strings[:1] if name_key.startswith("string") else strings
                    E, A = exact_string_expvals(used, psi_hi, with_abs=True)
                    exact = {"string.expval": lambda: E[0] * np.clongdouble(0.5 - 2j),
                             "op.expval": lambda: h_hi.astype(np.clongdouble) @ E,
                             "sop.expval": lambda: hk_hi.astype(np.clongdouble).T @ E}[name_key]()
                    e_gpu, e_ref = rel(got, exact.reshape(want.shape)), rel(want, exact.reshape(want.shape))
                    # The arbiter can only pass a GPU result that is within tolerance of the EXACT value, so it cannot
                    # hide a GPU bug; it is still held to account: it may fire only where the reference's sequential sum
                    # is expected to miss the bar -- a long reduction (>= 2^21 terms complex128, 2^12 complex64) or an
                    # ill-conditioned one (kappa = sum|terms| / |sum terms| >= 100: cancellation amplifies the
                    # reference's rounding error by kappa).  Anything else is reported as a failure.
                    terms = psi.shape[0] * len(used)
                    min_terms = (1 << 21) if psi.dtype == np.complex128 else (1 << 12)
                    w = {"string.expval": lambda: np.array([abs(0.5 - 2j)]),
                         "op.expval": lambda: np.abs(h_hi), "sop.expval": lambda: None}[name_key]()
                    absres = (A[0] * w[0] if name_key == "string.expval" else
                              (w.astype(np.longdouble) @ A if name_key == "op.expval" else
                               np.abs(hk_hi).astype(np.longdouble).T @ A))
                    kappa = float(np.max(absres.reshape(-1) / np.maximum(np.abs(exact).reshape(-1), 1e-300)))
                    if e_gpu < tol and e_gpu <= e_ref and (terms >= min_terms or kappa >= 100.0):
                        notes.append(f"{name}: reference-order rounding {e_ref:.2e} > tol, gpu {e_gpu:.2e}, "
                                     f"terms 2^{np.log2(terms):.1f}, kappa {kappa:.1e} | {tag}")
                        continue
                    e = e_gpu
                failures.append(f"{name}: rel err {e:.3e} | {tag}")
        except Exception as exc:  # noqa: BLE001 - report and keep fuzzing
            failures.append(f"EXC {type(exc).__name__}: {exc} | {tag}")
        if verbose:
            print(tag, "FAIL" if failures and tag in failures[-1] else "ok", flush=True)
    for nt in notes:
        print("NOTE", nt)
    LAST_NOTES[:] = notes
    return cases, failures


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--n-max", type=int, default=14, help="largest register (qubits)")
    ap.add_argument("--log2-elems", type=int, default=21, help="cap on dim * n_states")
    ap.add_argument("--s-cap", type=int, default=10**9, help="cap on the number of strings per operator")
    ap.add_argument("--native", action="store_true", help="also drive the pybind11 front-end (complex128 host cases)")
    a = ap.parse_args()
    n_cases, fails = run(a.seconds, a.seed, verbose=a.verbose, n_max=a.n_max, log2_elems=a.log2_elems, s_cap=a.s_cap,
                         native=a.native)
    for f in fails:
        print("FAIL", f)
    print(f"fuzz: {n_cases} cases, {len(fails)} failures (seed {a.seed})")
    sys.exit(1 if fails else 0)
