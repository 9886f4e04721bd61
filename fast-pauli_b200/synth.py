"""Host-side twin of the device input generator (fp_fill_uniform / fill_uniform_kernel in csrc/kernels.cuh).

Real scalar number e of a buffer (two per complex element: re then im) is
    u01(splitmix64(seed * 0xD1342543DE82EF95 + e)),  u01(h) = (h >> 11) * 2^-53,
so any element of a 2^34-amplitude state can be regenerated on the host without storing it.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def uniform_reals(idx_real: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = np.uint64((seed * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF) + idx_real.astype(np.uint64)
    return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_complex_at(idx_complex: np.ndarray, dtype=np.complex128, seed: int = 18) -> np.ndarray:
    """Values of the complex elements with flat indices ``idx_complex``."""
    idx = np.asarray(idx_complex, dtype=np.uint64)
    re = uniform_reals(idx * np.uint64(2), seed)
    im = uniform_reals(idx * np.uint64(2) + np.uint64(1), seed)
    if np.dtype(dtype) == np.complex64:
        return (re.astype(np.float32) + 1j * im.astype(np.float32)).astype(np.complex64)
    return re + 1j * im


def uniform_host(shape, dtype=np.complex128, seed: int = 18, first: int = 0) -> np.ndarray:
    n = int(np.prod(shape))
    idx = np.arange(first, first + n, dtype=np.uint64)
    return uniform_complex_at(idx, dtype, seed).reshape(shape)


def random_strings(rng: np.random.Generator, n_qubits: int, n_strings: int, max_weight: int | None = None) -> list[str]:
    """Synthetic operators of the benchmark: "random" = each qubit i.i.d. uniform over IXYZ
    (like the reference's tests/benchmarks/test_qiskit_adv.py:122-125); "weight <= w" = weight uniform in 1..w,
    positions uniform without replacement, letters uniform over XYZ."""
    letters = np.array(list("IXYZ"))
    out = []
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
