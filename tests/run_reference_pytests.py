#!/usr/bin/env python
"""Run the reference's OWN pytest files (tests/fast_pauli/test_{pauli,helpers,pauli_string,pauli_op,summed_pauli_op}.py)
against this package, as ``import fast_pauli``.

Works only where the reference checkout exists (``--reference``, default /root/reference -- it is never copied into
this repository: the files are staged in a temporary directory for the run).  The run directory gets

* ``fast_pauli/__init__.py`` -- a three-line alias package re-exporting ``Pauli, PauliString, PauliOp, SummedPauliOp,
  helpers`` from fast_pauli_b200; its ``__path__`` also points at the reference package so ``fast_pauli.pypauli`` (the
  reference's numpy implementation the tests compare against) resolves;
* ``tests/`` -- the reference's test tree, with ONE mechanical edit: ``parametrize("name,", [cls])`` is rewritten to
  ``parametrize("name", [cls])`` because pytest >= 8 no longer wraps a bare value for a single trailing-comma argname.

With ``--mock`` (default when no CUDA device is present) the C ABI is answered by tests/mock_abi.MockABI (CPU oracle):
that validates the Python surface and host logic only.  On a GPU box the real library runs.

Last run here (no GPU, --mock): 202 passed, 6 skipped (the skips are the reference's own `skip` marks);
--native without a GPU: 107 passed, 95 failed -- every failure is the no-device RuntimeError of a hot-path call.
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALIAS = '''import glob, importlib.util, os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests")]
import __graft_entry__ as _g
_m = _g.load_package()
if os.environ.get("FASTPAULI_REFTEST_NATIVE") == "1":  # the pybind11 module over the C++ classes
    _spec = importlib.util.spec_from_file_location(
        "_fast_pauli", glob.glob(os.path.join({root!r}, "fast-pauli_b200", "_fast_pauli*.so"))[0])
    _fast_pauli = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_fast_pauli)
    sys.modules["fast_pauli._fast_pauli"] = sys.modules["_fast_pauli"] = _fast_pauli  # pickle looks the classes up
    from fast_pauli._fast_pauli import Pauli, PauliOp, PauliString, SummedPauliOp, helpers  # noqa: F401,E402
else:
    if os.environ.get("FASTPAULI_REFTEST_MOCK") == "1":
        import mock_abi
        _m._default_ctx = None
        _m.lib = mock_abi.MockABI()
    from fast_pauli_b200 import Pauli, PauliOp, PauliString, SummedPauliOp, helpers  # noqa: F401,E402
sys.modules["fast_pauli.helpers"] = helpers
__path__.append({refpkg!r})  # fast_pauli.pypauli: the reference's numpy implementation used as the tests' oracle
'''

FILES = ["test_pauli.py", "test_helpers.py", "test_pauli_string.py", "test_pauli_op.py", "test_summed_pauli_op.py"]


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--mock", action="store_true", help="answer the C ABI with the CPU oracle (host-logic check)")
    ap.add_argument("--native", action="store_true",
                    help="test the pybind11 module _fast_pauli instead of the ctypes front-end (no mock possible: "
                         "without a GPU only the host-side cases pass, every hot-path case raises the no-device error)")
    ap.add_argument("pytest_args", nargs="*")
    args = ap.parse_args()
    ref_tests = os.path.join(args.reference, "tests")
    if not os.path.isdir(ref_tests):
        print(f"reference checkout not found at {args.reference}: nothing to run")
        return 0
    mock = args.mock and not args.native
    if not mock and not args.native:
        import ctypes

        lib = ctypes.CDLL(os.path.join(ROOT, "fast-pauli_b200", "lib", "libfastpauli_b200.so"))
        n = ctypes.c_int(0)
        if lib.fp_device_count(ctypes.byref(n)) != 0 or n.value == 0:
            print("no CUDA device: falling back to --mock (host-logic check only)")
            mock = True
    with tempfile.TemporaryDirectory(prefix="fp_reftests_") as run:
        shutil.copytree(ref_tests, os.path.join(run, "tests"))
        for f in FILES:
            p = os.path.join(run, "tests", "fast_pauli", f)
            src = open(p).read()
            src = re.sub(r'parametrize\(\s*"([a-z_]+),"', r'parametrize("\1"', src)
            src = re.sub(r'^(\s*)"([a-z_]+),",$', r'\1"\2",', src, flags=re.M)
            open(p, "w").write(src)
        alias = os.path.join(run, "alias", "fast_pauli")
        os.makedirs(alias)
        with open(os.path.join(alias, "__init__.py"), "w") as f:
            f.write(ALIAS.format(root=ROOT, refpkg=os.path.join(args.reference, "fast_pauli")))
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([run, os.path.join(run, "alias")]),
                   FASTPAULI_REFTEST_MOCK="1" if mock else "0", FASTPAULI_REFTEST_NATIVE="1" if args.native else "0")
        cmd = [sys.executable, "-m", "pytest", "-q", "--no-header", "-p", "no:cacheprovider"]
        cmd += [os.path.join("tests", "fast_pauli", f) for f in FILES] + args.pytest_args
        return subprocess.run(cmd, cwd=run, env=env).returncode


if __name__ == "__main__":
    sys.exit(main())
