"""The C++ class API (fast_pauli::PauliString / PauliOp / SummedPauliOp over the C ABI), driven through
tests/cpp/test_api.cpp -- a doctest-style driver that mirrors the reference's C++ test files.

CPU leg: the host-only part (value types, operator algebra, generator order, exceptions).
GPU leg: all nine hot-path methods for complex128 and complex64 against dense Kronecker-product oracles.
"""
from __future__ import annotations

import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "test_api")


def build() -> None:
    subprocess.run(["make", "-C", os.path.join(ROOT, "fast-pauli_b200")], check=True, capture_output=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)


def test_cpp_api_host_only():
    build()
    r = subprocess.run([EXE, "--host-only"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_cpp_api_gpu():
    if not os.path.exists(EXE):
        build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout and "host-only" not in r.stdout
