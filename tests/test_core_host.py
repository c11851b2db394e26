"""The product's per-thread GJK core (gjk_core.cuh + the table-driven sub-algorithm, gjk_tables.h) compiled for
the host and compared bit-for-bit with the oracle -- catches logic errors without a GPU."""
import ctypes

import numpy as np
import pytest

from conftest import live_simplex_equal
from golden_util import assert_matches_golden, golden_cases


def run_harness(lib, oracle_mod, a, b, dtype, unified=False):
    n = a.shape[0]
    simp = np.zeros(n, oracle_mod.simplex_dtype(dtype))
    dist = np.zeros(n, dtype)
    iters = np.zeros(n, np.int32)
    f32 = np.dtype(dtype) == np.float32
    if unified:
        fn = lib.harness_gjku_f32 if f32 else lib.harness_gjku_f64
    else:
        fn = lib.harness_gjk_f32 if f32 else lib.harness_gjk_f64
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    fn(ctypes.c_long(n), ctypes.c_void_p(a.ctypes.data), None, ctypes.c_int(a.shape[1]), ctypes.c_void_p(b.ctypes.data),
       None, ctypes.c_int(b.shape[1]), ctypes.c_void_p(simp.ctypes.data), ctypes.c_void_p(dist.ctypes.data),
       ctypes.c_void_p(iters.ctypes.data))
    return simp, dist, iters


@pytest.mark.parametrize("unified", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_core_reproduces_golden_gjk(host_harness, oracle_mod, dtype, unified):
    for _name, g in golden_cases(dtype):
        s, d, _ = run_harness(host_harness, oracle_mod, g["a"], g["b"], dtype, unified)
        assert_matches_golden(g, s, d, "gjk")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nverts,spread", [(64, 10.0), (32, 1.0), (8, 10.0), (8, 1.0), (4, 2.0), (5, 0.5), (300, 3.0), (1, 2.0)])
@pytest.mark.parametrize("unified", [False, True])
def test_core_equals_oracle(host_harness, oracle_mod, pkg, dtype, nverts, spread, unified):
    n = 30000 if nverts <= 64 else 3000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=2718, dtype=dtype)
    s, d, it = run_harness(host_harness, oracle_mod, a, b, dtype, unified)
    os_, od, oit = oracle_mod.Oracle("port", dtype).gjk(a, b, want_iters=True)
    assert np.array_equal(d, od)
    assert np.array_equal(it, oit)
    assert live_simplex_equal(s, os_)


def test_symmetric_inputs_tie_breaks(host_harness, oracle_mod, pkg):
    """cubes / grids: many exactly equal support values -> lowest-index rule must hold"""
    W = pkg.workloads
    for dtype in (np.float32, np.float64):
        base = W.cube_grid(6, 1.0, (0, 0, 0), dtype)
        shifts = [(0.5, 0, 0), (2, 0, 0), (2, 2, 0), (3, 3, 3), (0, 0, 0), (2.5, 0.25, -0.5), (0, 2, 0)]
        a = np.stack([base] * len(shifts))
        b = np.stack([W.cube_grid(6, 1.0, s, dtype) for s in shifts])
        for unified in (False, True):
            s, d, _ = run_harness(host_harness, oracle_mod, a, b, dtype, unified)
            os_, od = oracle_mod.Oracle("port", dtype).gjk(a, b)
            assert np.array_equal(d, od) and live_simplex_equal(s, os_)
