"""CPU check of the guarantee behind the fp16 pre-scan slot kernel (csrc/gjk_slots16.cuh): the vertices whose fp16
approximate support value is within the kernel's slack of the approximate maximum always include the reference's
support vertex and everything that ties it (GJK/cpu/openGJK.c:615-639).  oracle/fp16_prescan_model.py restates the
kernel's filter in numpy; the GPU parity tests (tests/test_gpu_slots16.py) check the kernel's results themselves."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("fp16_prescan_model", os.path.join(ROOT, "oracle", "fp16_prescan_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _dirs(rng, n, scale=1.0):
    return (rng.normal(size=(n, 3)) * scale).astype(F32)


@pytest.mark.parametrize("nv,spread", [(64, 10.0), (32, 1.0), (48, 10.0), (40, 0.5)])
def test_benchmark_generator(pkg, model, nv, spread):
    n = 40000
    rng = np.random.default_rng(nv)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=777, dtype=F32)
    d = (b.mean(axis=1) - a.mean(axis=1)).astype(F32)  # GJK's first directions: centre differences
    for body, dirs in ((a, d), (b, (-d).astype(F32)), (a, _dirs(rng, n))):
        ok, cnt = model.check(body, dirs)
        assert ok.all()
        assert cnt.mean() < 1.5  # the verification loop is sized for ~1.2 candidates per scan


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 1e3, 1e6])
def test_scales(pkg, model, scale):
    n = 20000
    rng = np.random.default_rng(3)
    a, _ = pkg.workloads.random_pairs(n, 64, 10.0, seed=3, dtype=F32)
    ok, cnt = model.check((a * F32(scale)).astype(F32), _dirs(rng, n))
    assert ok.all() and cnt.mean() < 1.5


def test_degenerate_inputs(pkg, model):
    n = 20000
    rng = np.random.default_rng(5)
    a, _ = pkg.workloads.random_pairs(n, 64, 10.0, seed=5, dtype=F32)
    tiny_far = ((a - a.mean(axis=1, keepdims=True)) * F32(1e-4) + F32(7.0)).astype(F32)
    dup = a.copy()
    dup[:, 32:] = dup[:, :32]
    lattice = rng.integers(-1, 2, size=(n, 64, 3)).astype(F32)
    point = np.repeat(a[:, :1], 64, axis=1)
    planar = a.copy()
    planar[:, :, 2] = planar[:, :1, 2]
    cases = [(tiny_far, _dirs(rng, n)), (dup, _dirs(rng, n)), (lattice, rng.integers(-2, 3, size=(n, 3)).astype(F32)),
             (lattice, np.tile(np.array([[0, 0, 1]], F32), (n, 1))), (point, _dirs(rng, n)), (planar, _dirs(rng, n)),
             (a, _dirs(rng, n, 1e-30)), (a, _dirs(rng, n, 1e30)), (a, _dirs(rng, n, 1e-12)),
             (planar, np.tile(np.array([[0, 0, 1]], F32), (n, 1)))]
    for body, dirs in cases:
        ok, _ = model.check(body, dirs)
        assert ok.all()


def test_adversarial_near_ties(model):
    """vertices whose exact support values differ by a few fp32 ulps: all of them must be candidates' superset"""
    rng = np.random.default_rng(11)
    n = 20000
    base = rng.normal(size=(n, 1, 3)).astype(F32)
    a = np.repeat(base, 64, axis=1)
    a = (a + (rng.integers(-3, 4, size=a.shape) * np.spacing(np.abs(a)))).astype(F32)  # a cloud of ulp-neighbours
    a[:, 0] += F32(0.25)  # a body needs some extent, or everything is a candidate trivially
    ok, _ = model.check(a, _dirs(rng, n))
    assert ok.all()
