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


def test_host_planners_evaluate_to_the_definition():
    """pack.hpp + coset_plan.hpp (host code, no GPU): the packed / planned operator, evaluated on the host with the
    kernels' indexing, equals the definition for 165 random operators x every tile rank x reserved low bits."""
    build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_host_plan")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "failed: 0" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_cpp_api_gpu():
    if not os.path.exists(EXE):
        build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout and "host-only" not in r.stdout


# ---- the reference's own example programs (fast_pauli/cpp/examples/01..05), compiled UNMODIFIED from where they lie
# against this repository's fast_pauli.hpp and library by tests/cpp/Makefile (binaries only, tests/cpp/_ref_examples/)
EX_DIR = os.path.join(ROOT, "tests", "cpp", "_ref_examples")
EXAMPLES = ["01_pauli_op", "02_pauli_op_multistate", "03_summed_pauli_op", "04_get_sparse_repr", "05_summed_pauli_op_sq"]


def test_reference_examples_compile_against_this_api():
    if not os.path.isdir("/root/reference/fast_pauli/cpp/examples"):
        pytest.skip("reference checkout not present (the binaries are prebuilt for the GPU box)")
    build()  # -Wall -Wextra -Werror
    for ex in EXAMPLES:
        assert os.path.exists(os.path.join(EX_DIR, ex)), ex
    # example 04 only needs host code (get_sparse_repr); it must run anywhere
    r = subprocess.run([os.path.join(EX_DIR, "04_get_sparse_repr")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("example", EXAMPLES)
def test_reference_examples_run_on_gpu(example):
    exe = os.path.join(EX_DIR, example)
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_ref_examples was not built (needs the reference checkout at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


# ---- the reference's own C++ test files (fast_pauli/cpp/tests/test_*.cpp), compiled UNMODIFIED against this
# repository's headers with tests/cpp/shim/doctest/doctest.h standing in for doctest (binaries: tests/cpp/_ref_tests/)
RT_DIR = os.path.join(ROOT, "tests", "cpp", "_ref_tests")
REF_TESTS_HOST = ["test_factory", "test_pauli", "test_pauli_helpers"]  # no hot-path call: must pass anywhere
REF_TESTS_GPU = ["test_pauli_string", "test_pauli_op", "test_summed_pauli_op"]


def test_reference_cpp_tests_compile_and_host_cases_pass():
    if not os.path.isdir("/root/reference/fast_pauli/cpp/tests"):
        pytest.skip("reference checkout not present (the binaries are prebuilt for the GPU box)")
    build()
    for t in REF_TESTS_HOST:
        r = subprocess.run([os.path.join(RT_DIR, t)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "failed: 0" in r.stdout, r.stdout[-2000:]
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "fast-pauli_b200", "lib", "libfastpauli_b200.so"))
    n = ctypes.c_int(0)
    if lib.fp_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        return  # the GPU leg below covers the rest
    # without a device every failure must be the loud "no CUDA device" error of a hot-path call, never a wrong value
    for t in REF_TESTS_GPU:
        r = subprocess.run([os.path.join(RT_DIR, t)], capture_output=True, text=True, timeout=600)
        assert "CHECK(" not in r.stdout, r.stdout[-2000:]
        threw = [ln for ln in r.stdout.splitlines() if "threw:" in ln]
        assert threw and all("fastpauli_b200:" in ln for ln in threw), threw[:5]


@pytest.mark.gpu
@pytest.mark.parametrize("name", REF_TESTS_HOST + REF_TESTS_GPU)
def test_reference_cpp_tests_run_on_gpu(name):
    exe = os.path.join(RT_DIR, name)
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_ref_tests was not built (needs the reference checkout at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    tail = "\n".join(ln for ln in r.stdout.splitlines() if "CHECK(" in ln or "threw" in ln or "doctest-shim" in ln)
    assert r.returncode == 0 and "failed: 0" in r.stdout, tail[-3000:]
