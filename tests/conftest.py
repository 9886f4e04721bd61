"""pytest configuration: markers, import paths, shared input generators."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Every time a parity check does NOT pass the plain |gpu - reference| < tol comparison and is settled some other way,
# it is recorded here and printed at the end of the run (the relaxations must be countable).
#   like_ordered : passed against the pairwise-summed host reference in the SAME precision (no relaxation of the bar)
#   arbiter      : passed only against a higher-precision arbiter (the documented exception)
AUDIT: dict[str, list[str]] = {"like_ordered": [], "arbiter": []}
# the arbiter may only ever be needed for long reductions: below these term counts it is a failure, not an exception
ARBITER_MIN_TERMS = {"complex64": 1 << 12, "complex128": 1 << 21}


def pytest_terminal_summary(terminalreporter):
    tr = terminalreporter
    tr.write_line(f"parity audit: {len(AUDIT['like_ordered'])} check(s) settled by the like-ordered (pairwise, same "
                  f"precision) reference, {len(AUDIT['arbiter'])} by the higher-precision arbiter")
    for kind in ("like_ordered", "arbiter"):
        for line in AUDIT[kind][:40]:
            tr.write_line(f"  [{kind}] {line}")


def rand_states(rng: np.random.Generator, dim: int, n_states: int | None, dtype=np.complex128) -> np.ndarray:
    """U[0,1) + i U[0,1) like the reference fixtures (tests/conftest.py:86-90, __factory.hpp:110-122)."""
    shape = (dim,) if n_states is None else (dim, n_states)
    return (rng.random(shape) + 1j * rng.random(shape)).astype(dtype)


def rand_strings(rng: np.random.Generator, n_qubits: int, n_strings: int, max_weight: int | None = None) -> list[str]:
    """i.i.d. uniform IXYZ strings (tests/benchmarks/test_qiskit_adv.py:122-125) or weight <= w strings."""
    out = []
    letters = np.array(list("IXYZ"))
    for _ in range(n_strings):
        if max_weight is None:
            out.append("".join(letters[rng.integers(0, 4, size=n_qubits)]))
        else:
            w = int(rng.integers(1, max_weight + 1))
            pos = rng.choice(n_qubits, size=min(w, n_qubits), replace=False)
            s = ["I"] * n_qubits
            for p in pos:
                s[p] = "XYZ"[int(rng.integers(0, 3))]
            out.append("".join(s))
    return out


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max(|b|_inf, tiny): the parity metric of SURVEY.md 8(d)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 1e-5}


@pytest.fixture
def rng():
    return np.random.default_rng(18)
