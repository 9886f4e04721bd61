"""``import fast_pauli`` for code written against the reference package (fast_pauli/__init__.py:18-25).

Put ``<repo>/fast-pauli_b200/compat`` on ``sys.path`` / ``PYTHONPATH`` and existing scripts keep their imports:

    import fast_pauli as fp
    op = fp.PauliOp(coeffs, strings); new_states = op.apply(states)      # runs on the B200

The names are the GPU-backed classes of ``fast_pauli_b200`` (ctypes front-end; same signatures)
or, with ``FASTPAULI_FRONTEND=native`` in the environment, of the pybind11 module ``_fast_pauli`` over the C++ classes
(complex128 host arrays only, exactly the reference's binding surface).  The qiskit
converters of the reference are outside the hot path and are not rebuilt: calling them raises
``NotImplementedError`` instead of failing at import time.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # .../fast-pauli_b200

if "fast_pauli_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "fast_pauli_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG]
    )
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["fast_pauli_b200"] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:
        sys.modules.pop("fast_pauli_b200", None)
        raise

if os.environ.get("FASTPAULI_FRONTEND", "ctypes") == "native":
    import glob

    _hits = glob.glob(os.path.join(_PKG, "_fast_pauli*.so"))
    if not _hits:
        raise ImportError("FASTPAULI_FRONTEND=native but fast-pauli_b200/_fast_pauli*.so is not built "
                          "(make -C fast-pauli_b200)")
    _nspec = importlib.util.spec_from_file_location("_fast_pauli", _hits[0])
    _fast_pauli = importlib.util.module_from_spec(_nspec)
    _nspec.loader.exec_module(_fast_pauli)
    sys.modules[__name__ + "._fast_pauli"] = sys.modules["_fast_pauli"] = _fast_pauli  # pickle looks the classes up
    from ._fast_pauli import Pauli, PauliOp, PauliString, SummedPauliOp, helpers  # noqa: E402,F401
else:
    from fast_pauli_b200 import Pauli, PauliOp, PauliString, SummedPauliOp, helpers  # noqa: E402,F401

sys.modules.setdefault(__name__ + ".helpers", helpers)


def from_qiskit(*_args, **_kwargs):
    raise NotImplementedError("fast_pauli.from_qiskit is not part of the B200 hot-path build")


def to_qiskit(*_args, **_kwargs):
    raise NotImplementedError("fast_pauli.to_qiskit is not part of the B200 hot-path build")


__all__ = ["Pauli", "PauliOp", "PauliString", "SummedPauliOp", "helpers", "from_qiskit", "to_qiskit"]
