"""GPU parity: EPA (penetration depth, witnesses, contact normal) against the CPU oracle."""
import os

import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def _oracle_gjk_epa(oracle_mod, dtype, a, b, kind="port"):
    orc = oracle_mod.Oracle(kind, dtype)
    s, d = orc.gjk(a, b)
    return orc.epa(a, b, s, d)


def _compare(dtype, got, want):
    simp, dist, nrm = got
    osimp, odist, onrm = want
    eps = np.finfo(dtype).eps
    assert np.array_equal(dist <= eps, odist <= eps), "collision verdict differs"
    rtol = RTOL[np.dtype(dtype)]
    np.testing.assert_allclose(dist, odist, rtol=rtol, atol=0)
    np.testing.assert_allclose(nrm, onrm, rtol=rtol, atol=rtol)
    np.testing.assert_allclose(simp["witnesses"], osimp["witnesses"], rtol=rtol, atol=rtol)
    assert np.array_equal(dist, odist), "distances not bit-identical"
    assert np.array_equal(nrm, onrm), "normals not bit-identical"
    assert live_simplex_equal(simp, osimp), "simplices not bit-identical"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nverts,spread", [(32, 1.0), (8, 1.0), (64, 2.0), (200, 1.5), (32, 10.0), (5, 0.5)])
def test_gjk_then_epa_matches_oracle(pkg, oracle_mod, dtype, nverts, spread):
    n = 20000 if nverts <= 64 else 3000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=99, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    got = eng.compute_gjk_epa(bd1, bd2)
    _compare(dtype, got, _oracle_gjk_epa(oracle_mod, dtype, a, b))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_epa_only_from_oracle_gjk(pkg, oracle_mod, dtype):
    """computeCollisionInformation / GJK::GPU::computeEPA: caller supplies GJK's simplices + distances."""
    a, b = pkg.workloads.random_pairs(10000, 32, 1.0, seed=5, dtype=dtype)
    orc = oracle_mod.Oracle("port", dtype)
    s, d = orc.gjk(a, b)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    got = eng.compute_collision_information(bd1, bd2, s.copy(), d.copy())
    _compare(dtype, got, orc.epa(a, b, s, d))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_against_compiled_reference(pkg, oracle_mod, dtype):
    if not oracle_mod.available("ref", dtype):
        pytest.skip("oracle/_ref not built")
    a, b = pkg.workloads.random_pairs(10000, 32, 1.0, seed=31, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    got = eng.compute_gjk_epa(bd1, bd2)
    _compare(dtype, got, _oracle_gjk_epa(oracle_mod, dtype, a, b, kind="ref"))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_symmetric_shapes_cubes(pkg, oracle_mod, dtype):
    """Reference EPATesting cases 1-3 (examples/gpu/example.cu:385-580): cubes shifted 1 / 2 / 5 along x, plus the
    README rotated cube.  These exercise exact ties in the support and closest-face searches."""
    W = pkg.workloads
    base = W.unit_cube(dtype=dtype)
    a = np.stack([base, base, base, base])
    b = np.stack([W.unit_cube((1, 0, 0), dtype), W.unit_cube((2, 0, 0), dtype), W.unit_cube((5, 0, 0), dtype),
                  W.rotated_cube_readme(dtype)])
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    got = eng.compute_gjk_epa(bd1, bd2)
    want = _oracle_gjk_epa(oracle_mod, dtype, a, b)
    _compare(dtype, got, want)
    # known answers (SURVEY.md section 4 golden table)
    assert got[1][0] == -1.0 and tuple(got[2][0]) == (1.0, 0.0, 0.0)
    assert got[1][2] == 3.0
    assert abs(got[1][3] + 1.5) < 1e-6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grid_cubes_and_spheres(pkg, oracle_mod, dtype):
    """Reference EPATesting cases 6-8: 9600-vertex grid cubes, 1000-point spheres 1 apart."""
    W = pkg.workloads
    c1 = W.cube_grid(40, 1.0, (0, 0, 0), dtype)
    c2 = W.cube_grid(40, 1.0, (1.2, 0.3, 0.1), dtype)
    s1 = W.sphere_surface(1000, 2.0, (0, 0, 0), 7, dtype)
    s2 = W.sphere_surface(1000, 2.0, (1, 0, 0), 8, dtype)
    eng = pkg.Engine(dtype)
    orc = oracle_mod.Oracle("port", dtype)
    for a, b in ((c1, c2), (s1, s2)):
        bd1, _k1 = pkg.make_polytopes(a[None])
        bd2, _k2 = pkg.make_polytopes(b[None])
        got = eng.compute_gjk_epa(bd1, bd2)
        s, d = orc.gjk(a[None], b[None])
        _compare(dtype, got, orc.epa(a[None], b[None], s, d))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", ["warp", "group", "small4", "small4+small", "small8"])
@pytest.mark.parametrize("nverts,spread", [(32, 1.0), (64, 2.0), (200, 1.5), (8, 0.5), (16, 0.3)])
def test_epa_kernel_families_match_oracle(pkg, oracle_mod, dtype, kernel, nverts, spread):
    """every EPA kernel family forced through OGJK_EPA_KERNEL (needs >= 8192 pairs to leave the tiny-batch kernel):
    warp per pair; sub-warp group with the full-size work area; sub-warp groups of 4 / 8 lanes with the small work
    areas + overflow pass -- deep overlaps (spread 0.3 .. 1) send the long-tailed pairs through the overflow queue.
    `small4` in fp32 is the default policy's pair of 96-register instantiations (1.4 KB area up to 16 vertices, the lean
    area above: bodies of more than 32 vertices all take the overflow pass there); `small4+small` (OGJK_EPA_AREA=small) and
    fp64 run the 1.7 KB area, where 200-vertex bodies take the uncached support path of the group kernel"""
    import os
    n = 30000 if nverts <= 64 else 9000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=123, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    saved = {k: os.environ.get(k) for k in ("OGJK_EPA_KERNEL", "OGJK_EPA_AREA")}
    os.environ["OGJK_EPA_KERNEL"] = kernel.split("+")[0]
    if "+" in kernel:
        os.environ["OGJK_EPA_AREA"] = kernel.split("+")[1]
    try:
        got = eng.compute_gjk_epa(bd1, bd2)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    orc = oracle_mod.Oracle("port", dtype)
    s, d = orc.gjk(a, b, nthreads=8)
    _compare(dtype, got, orc.epa(a, b, s, d, nthreads=8))


def test_small_work_area_overflow_is_exercised(pkg, oracle_mod):
    """the workload above really has pairs beyond the small work areas' 22 .. 24 iterations (so the overflow pass ran)"""
    a, b = pkg.workloads.random_pairs(30000, 32, 1.0, seed=123, dtype=np.float32)
    orc = oracle_mod.Oracle("port", np.float32)
    s, d = orc.gjk(a, b, nthreads=8)
    _s, _d, _n, it = orc.epa(a, b, s, d, nthreads=8, want_iters=True)
    assert (it > 26).sum() >= 3


@pytest.mark.parametrize("kernel", ["auto", "warp", "small4", "small8", "group"])
def test_sign_of_zero_on_lattice_cubes(pkg, oracle_mod, kernel):
    """Axis-aligned cubes on a quarter-integer lattice (and the README's rotated cube): face normals are full of exact
    zeros, a sixth of them negative (the reference's README prints `-0.000000` for one).  Outputs must match the reference as BIT PATTERNS --
    np.array_equal alone would accept +0.0 for -0.0.  (Caught a shared-reciprocal division whose correction step turned
    -0 / len into +0.)"""
    from conftest import same_bits
    import torch
    dtype = np.float32
    rng = np.random.Generator(np.random.Philox(key=[41, 7]))
    n = 40000
    cube = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
    a = cube[np.argsort(rng.random((n, 8)), axis=1)]  # the corners in a random order per pair
    perm = np.argsort(rng.random((n, 8)), axis=1)
    rot = pkg.workloads.rotated_cube_readme(np.float64) - np.array([1.0, 0.0, 0.0])
    half = n // 2
    b = np.empty((n, 8, 3))
    b[:half] = cube[perm[:half]] + rng.integers(-6, 7, size=(half, 1, 3)) * 0.25
    b[half:] = rot[perm[half:]] + rng.integers(-6, 7, size=(n - half, 1, 3)) * 0.25
    a, b = np.ascontiguousarray(a.astype(dtype)), np.ascontiguousarray(b.astype(dtype))
    orc = oracle_mod.Oracle("ref" if oracle_mod.available("ref", dtype) else "port", dtype)
    s, d = orc.gjk(a, b, nthreads=8)
    es, ed, en = orc.epa(a, b, s, d, nthreads=8)
    assert np.signbit(en[en == 0]).any(), "the case is meant to contain negative zeros"
    eng = pkg.Engine(dtype)
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
    saved = os.environ.get("OGJK_EPA_KERNEL")
    try:
        if kernel == "auto":
            os.environ.pop("OGJK_EPA_KERNEL", None)
        else:
            os.environ["OGJK_EPA_KERNEL"] = kernel
        eng.gjk_epa_uniform_device(n, 8, d_a, 8, d_b, d_simp, d_dist, d_nrm)
        torch.cuda.synchronize()
    finally:
        if saved is None:
            os.environ.pop("OGJK_EPA_KERNEL", None)
        else:
            os.environ["OGJK_EPA_KERNEL"] = saved
    got = d_simp.cpu().numpy().view(eng.sdtype)
    assert same_bits(d_dist.cpu().numpy(), ed)
    assert same_bits(d_nrm.cpu().numpy(), en)
    assert same_bits(got["witnesses"], es["witnesses"])
