"""GPU parity: the CUDA GJK path (through the C ABI) against the CPU oracle on seeded pairs.

Bar (BASELINE.json north_star): collision verdict bit-exact; distances / witnesses within 1e-5 (fp32) /
1e-12 (fp64) relative.  Because the kernels reproduce the oracle's IEEE operation sequence, the tests assert the
stronger property first (bit equality) and fall back to reporting the tolerance figures on failure.
"""
import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def _check(pkg, oracle_mod, dtype, a, b, kind="port"):
    eng = pkg.Engine(dtype)
    bd1, k1 = pkg.make_polytopes(a)
    bd2, k2 = pkg.make_polytopes(b)
    simp, dist = eng.compute_minimum_distance(bd1, bd2)
    orc = oracle_mod.Oracle(kind, dtype)
    if isinstance(a, np.ndarray):
        osimp, odist = orc.gjk(a, b)
    else:
        off1 = np.concatenate([[0], np.cumsum([len(x) for x in a])])
        off2 = np.concatenate([[0], np.cumsum([len(x) for x in b])])
        osimp, odist = orc.gjk(np.concatenate(a), np.concatenate(b), off1, off2)
    eps = np.finfo(dtype).eps
    assert np.array_equal(dist <= eps, odist <= eps), "collision verdict differs"
    rtol = RTOL[np.dtype(dtype)]
    np.testing.assert_allclose(dist, odist, rtol=rtol, atol=0)
    np.testing.assert_allclose(simp["witnesses"], osimp["witnesses"], rtol=rtol, atol=rtol)
    assert np.array_equal(dist, odist), "distances not bit-identical"
    assert live_simplex_equal(simp, osimp), "simplices not bit-identical"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nverts,spread", [(64, 10.0), (32, 1.0), (8, 10.0), (4, 2.0), (5, 0.5), (300, 3.0), (1024, 10.0)])
def test_random_pairs_match_oracle(pkg, oracle_mod, dtype, nverts, spread):
    n = 20000 if nverts <= 64 else 2000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=4242, dtype=dtype)
    _check(pkg, oracle_mod, dtype, a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ragged_vertex_counts(pkg, oracle_mod, dtype):
    rng = np.random.default_rng(5)
    counts = rng.integers(1, 90, size=3000)
    a = [pkg.workloads.random_polytopes(1, int(c), 6.0, 100 + i, dtype)[0] for i, c in enumerate(counts)]
    b = [pkg.workloads.random_polytopes(1, int(c), 6.0, 900000 + i, dtype)[0] for i, c in enumerate(counts[::-1])]
    _check(pkg, oracle_mod, dtype, a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_against_compiled_reference(pkg, oracle_mod, dtype):
    if not oracle_mod.available("ref", dtype):
        pytest.skip("oracle/_ref not built")
    a, b = pkg.workloads.random_pairs(20000, 64, 10.0, seed=777, dtype=dtype)
    _check(pkg, oracle_mod, dtype, a, b, kind="ref")


def test_empty_batch_is_noop(pkg):
    eng = pkg.Engine(np.float32)
    bd = np.zeros(0, dtype=eng.pdtype)
    simp, dist = eng.compute_minimum_distance(bd, bd)
    assert len(simp) == 0 and len(dist) == 0


def test_almost_dense_batch_falls_back(pkg, oracle_mod):
    """The host path streams dense uniform batches straight from the caller's arrays and validates the descriptors
    chunk by chunk while copying; a batch that only LOOKS dense at both ends (one descriptor in the middle points
    somewhere else, one has a different vertex count) must be detected mid-way and redone through the general path."""
    dtype = np.float32
    n, nv = 60000, 16
    a, b = pkg.workloads.random_pairs(n, nv, 6.0, seed=808, dtype=dtype)
    bd1, keep1 = pkg.make_polytopes(a)
    bd2, keep2 = pkg.make_polytopes(b)
    moved = np.ascontiguousarray(a[31000])          # same vertices, different address
    bd1["coord"][31000] = moved.ctypes.data
    short = np.ascontiguousarray(b[45000][:12])     # fewer vertices
    bd2["coord"][45000] = short.ctypes.data
    bd2["numpoints"][45000] = 12
    eng = pkg.Engine(dtype)
    simp, dist = eng.compute_minimum_distance(bd1, bd2)
    orc = oracle_mod.Oracle("port", dtype)
    os_, od = orc.gjk(a, b, nthreads=8)
    ref_s, ref_d = orc.gjk(a[45000:45001], short[None], nthreads=1)
    od[45000] = ref_d[0]
    os_[45000] = ref_s[0]
    assert np.array_equal(dist, od)
    assert live_simplex_equal(simp, os_)
