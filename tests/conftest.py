import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkgpath import load_oracle, load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    mod = load_oracle()
    if not mod.available("port"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"),
                               "lib/libogjk_oracle_f32.so", "lib/libogjk_oracle_f64.so"])
    return mod


@pytest.fixture(scope="session")
def host_harness():
    """The product's per-thread GJK core compiled for the host (tests/host_harness.cpp)."""
    out = os.path.join(ROOT, "tests", "_build", "libhost_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    csrc = os.path.join(ROOT, "opengjk-gpu_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("gjk_core.cuh", "gjk_math.cuh", "gjk_tables.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               "-I", csrc, src, "-o", out])
    return ctypes.CDLL(out)


def _same(got, want):
    """float arrays equal as BIT PATTERNS (so -0.0 differs from +0.0); two NaNs match whatever their payload (the CPU and
    the GPU generate different default NaNs)"""
    got, want = np.ascontiguousarray(got), np.ascontiguousarray(want)
    u = np.uint32 if got.dtype.itemsize == 4 else np.uint64
    return bool(np.all((got.view(u) == want.view(u)) | (np.isnan(got) & np.isnan(want))))


def live_simplex_equal(got, want):
    """gkSimplex equality on what is defined: nvrtx, the live slots, the witnesses -- floats as bit patterns."""
    if not np.array_equal(got["nvrtx"], want["nvrtx"]):
        return False
    if not _same(got["witnesses"], want["witnesses"]):
        return False
    for j in range(4):
        live = want["nvrtx"] > j
        if not _same(got["vrtx"][live, j], want["vrtx"][live, j]):
            return False
        if not np.array_equal(got["vrtx_idx"][live, j], want["vrtx_idx"][live, j]):
            return False
    return True


def same_bits(got, want):
    """bit-pattern equality of two float arrays: -0.0 differs from +0.0, a NaN only matches the same NaN"""
    got, want = np.ascontiguousarray(got), np.ascontiguousarray(want)
    if got.shape != want.shape or got.dtype != want.dtype:
        return False
    u = np.uint32 if got.dtype.itemsize == 4 else np.uint64
    return bool(np.array_equal(got.view(u), want.view(u)))
