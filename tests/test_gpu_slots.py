"""GPU parity of the persistent slot kernels (self-service and warp-specialised) and of the fused GJK+EPA device
entry point, each forced through the development override so that batches small enough for the oracle still take
the kernel under test.  Bit-exact against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_kernel():
    saved = {k: os.environ.get(k) for k in ("OGJK_GJK_KERNEL", "OGJK_WS_MIN_SLOT", "OGJK_WS_LP")}

    def setter(name, ws_min_slot=None):
        os.environ["OGJK_GJK_KERNEL"] = name
        if ws_min_slot is not None:
            os.environ["OGJK_WS_MIN_SLOT"] = str(ws_min_slot)

    yield setter
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _device_batch(pkg, a, b, dtype=np.float32):
    import torch
    n = a.shape[0]
    eng = pkg.Engine(dtype)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=tdt, device="cuda")
    d_nrm = torch.zeros(n, 3, dtype=tdt, device="cuda")
    return eng, d_a, d_b, d_simp, d_dist, d_nrm


@pytest.mark.parametrize("kernel", ["slots", "slotsws"])
@pytest.mark.parametrize("nv1,nv2,spread", [(64, 64, 10.0), (32, 32, 1.0), (32, 32, 10.0), (8, 8, 10.0), (4, 4, 2.0),
                                            (12, 20, 3.0), (64, 16, 6.0), (68, 68, 8.0), (16, 16, 0.5), (96, 96, 4.0), (128, 64, 5.0),
                                            (140, 140, 10.0)])
def test_slot_kernels_match_oracle(pkg, oracle_mod, force_kernel, kernel, nv1, nv2, spread):
    import torch
    n = 40000
    a = pkg.workloads.random_polytopes(n, nv1, spread, 11, np.float32, stream=1)
    b = pkg.workloads.random_polytopes(n, nv2, spread, 11, np.float32, stream=2)
    eng, d_a, d_b, d_simp, d_dist, _ = _device_batch(pkg, a, b)
    force_kernel(kernel)
    eng.gjk_uniform_device(n, nv1, d_a, nv2, d_b, d_simp, d_dist)
    torch.cuda.synchronize()
    os_, od = oracle_mod.Oracle("port", np.float32).gjk(a, b, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), od)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_)
    eng.set_stream(0)


@pytest.mark.parametrize("nv1,nv2,spread", [(64, 64, 10.0), (64, 48, 3.0), (60, 68, 1.0)])
def test_ws_two_lanes_per_pair(pkg, oracle_mod, force_kernel, nv1, nv2, spread):
    """the experimental two-lanes-per-pair mode of the warp-specialised kernel (OGJK_WS_LP=2)"""
    import torch
    n = 40000
    a = pkg.workloads.random_polytopes(n, nv1, spread, 21, np.float32, stream=1)
    b = pkg.workloads.random_polytopes(n, nv2, spread, 21, np.float32, stream=2)
    eng, d_a, d_b, d_simp, d_dist, _ = _device_batch(pkg, a, b)
    force_kernel("slotsws")
    os.environ["OGJK_WS_LP"] = "2"
    eng.gjk_uniform_device(n, nv1, d_a, nv2, d_b, d_simp, d_dist)
    torch.cuda.synchronize()
    os_, od = oracle_mod.Oracle("port", np.float32).gjk(a, b, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), od)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_)
    eng.set_stream(0)


@pytest.mark.parametrize("kernel", ["slots", "slotsws", "auto"])
@pytest.mark.parametrize("nv1,nv2,spread", [(64, 64, 10.0), (32, 32, 1.0), (16, 16, 3.0), (8, 8, 10.0), (24, 40, 4.0),
                                            (48, 48, 0.5), (4, 4, 2.0)])
def test_slot_kernels_fp64(pkg, oracle_mod, force_kernel, kernel, nv1, nv2, spread):
    """the slot kernels in double precision (96-byte vertex blocks, scalar DMUL/DADD scan)"""
    import torch
    n = 40000
    a = pkg.workloads.random_polytopes(n, nv1, spread, 31, np.float64, stream=1)
    b = pkg.workloads.random_polytopes(n, nv2, spread, 31, np.float64, stream=2)
    eng, d_a, d_b, d_simp, d_dist, d_nrm = _device_batch(pkg, a, b, np.float64)
    force_kernel(kernel)
    orc = oracle_mod.Oracle("port", np.float64)
    os_, od = orc.gjk(a, b, nthreads=8)
    eng.gjk_uniform_device(n, nv1, d_a, nv2, d_b, d_simp, d_dist)
    torch.cuda.synchronize()
    assert np.array_equal(d_dist.cpu().numpy(), od)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_)
    if nv1 == nv2:  # fused GJK + EPA entry (finisher-warp gate when the ws kernel is taken)
        eng.gjk_epa_uniform_device(n, nv1, d_a, nv2, d_b, d_simp, d_dist, d_nrm)
        torch.cuda.synchronize()
        es, ed, en = orc.epa(a, b, os_, od, nthreads=8)
        assert np.array_equal(d_dist.cpu().numpy(), ed)
        assert np.array_equal(d_nrm.cpu().numpy(), en)
        assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), es)
    eng.set_stream(0)


def test_slot_kernels_symmetric_shapes(pkg, oracle_mod, force_kernel):
    """grids of cube points: many exactly tied support values (lowest-index rule), touching and overlapping cases"""
    import torch
    W = pkg.workloads
    base = W.cube_grid(2, 1.0, (0, 0, 0), np.float32)  # the 8 corners, three copies each: 24 points
    shifts = [(0.5, 0, 0), (2, 0, 0), (2, 2, 0), (3, 3, 3), (0, 0, 0), (2.5, 0.25, -0.5), (0, 2, 0), (1, 1, 1)]
    reps = 40000 // len(shifts)
    a = np.ascontiguousarray(np.stack([base] * (len(shifts) * reps)))
    b = np.ascontiguousarray(np.stack([W.cube_grid(2, 1.0, s, np.float32) for s in shifts] * reps))
    assert a.shape[1] % 4 == 0
    os_, od = oracle_mod.Oracle("port", np.float32).gjk(a, b, nthreads=8)
    for kernel in ("slots", "slotsws"):
        eng, d_a, d_b, d_simp, d_dist, _ = _device_batch(pkg, a, b)
        force_kernel(kernel)
        eng.gjk_uniform_device(a.shape[0], a.shape[1], d_a, b.shape[1], d_b, d_simp, d_dist)
        torch.cuda.synchronize()
        assert np.array_equal(d_dist.cpu().numpy(), od), kernel
        assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_), kernel
        eng.set_stream(0)


@pytest.mark.parametrize("nverts,spread,ws_min_slot", [(64, 10.0, None), (32, 1.0, None), (32, 10.0, None),
                                                        (16, 2.0, 0), (16, 2.0, None)])
def test_fused_gjk_epa_uniform_device(pkg, oracle_mod, force_kernel, nverts, spread, ws_min_slot):
    """gjk_epa_uniform_device: the finisher warp's fused EPA gate + the queue kernel against oracle GJK then EPA"""
    import torch
    n = 40000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=99, dtype=np.float32)
    eng, d_a, d_b, d_simp, d_dist, d_nrm = _device_batch(pkg, a, b)
    if ws_min_slot is not None:
        force_kernel("auto", ws_min_slot)  # "auto" is not a kernel name: normal selection, custom threshold
    eng.gjk_epa_uniform_device(n, nverts, d_a, nverts, d_b, d_simp, d_dist, d_nrm)
    torch.cuda.synchronize()
    orc = oracle_mod.Oracle("port", np.float32)
    s, d = orc.gjk(a, b, nthreads=8)
    s, d, nr = orc.epa(a, b, s, d, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), d)
    assert np.array_equal(d_nrm.cpu().numpy(), nr)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), s)
    eng.set_stream(0)


@pytest.mark.parametrize("kernel", ["auto", "slots", "slotsws"])
@pytest.mark.parametrize("nverts", [32, 64])
def test_indexed_uniform_pool_takes_slot_kernels(pkg, oracle_mod, force_kernel, kernel, nverts):
    """BASELINE config 5 in small: uniform pool + gkCollisionPair list (broad-phase candidates) through the host-level
    and device-level indexed entry points; >= 32768 pairs so that the slot kernels are fed from the pair list."""
    npoly = 1500
    pool, pairs = pkg.workloads.broadphase_pool(npoly, nverts, 45000, seed=17)
    assert pairs.shape[0] >= 32768
    off = np.arange(npoly + 1) * nverts
    flat = pool.reshape(-1, 3)
    orc = oracle_mod.Oracle("port", np.float32)
    gs, gd, _ = orc.gjk_epa_indexed(flat, pairs, off, do_epa=False, nthreads=8)
    es, ed, en = orc.gjk_epa_indexed(flat, pairs, off, nthreads=8)
    eng = pkg.Engine(np.float32)
    desc, _keep = pkg.make_polytopes(pool)
    force_kernel(kernel)
    s, d = eng.compute_minimum_distance_indexed(desc, pairs)
    assert np.array_equal(d, gd) and live_simplex_equal(s, gs)
    s3, d3, n3 = eng.compute_gjk_epa_indexed(desc, pairs)
    assert np.array_equal(d3, ed) and np.array_equal(n3, en) and live_simplex_equal(s3, es)
    dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(desc, len(pairs))
    try:
        eng.upload_pairs_device(pairs, dpairs)
        eng.compute_minimum_distance_indexed_device(len(pairs), dp, dpairs, dsimp, ddist)
        s4, d4 = eng.copy_results_from_device(len(pairs), dsimp, ddist)
        assert np.array_equal(d4, gd) and live_simplex_equal(s4, gs)
        eng.compute_epa_indexed_device(len(pairs), dp, dpairs, dsimp, ddist, dnrm)
        s5, d5 = eng.copy_results_from_device(len(pairs), dsimp, ddist)
        assert np.array_equal(d5, ed) and live_simplex_equal(s5, es)
    finally:
        eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)


@pytest.mark.parametrize("nverts", [32, 64])
def test_indexed_pool_soa4_packed(pkg, oracle_mod, force_kernel, nverts):
    """the SoA-4 re-packed pool (packed multiplies and packed adds in the scan) against the oracle, and against the
    same batch through the unpacked pool"""
    npoly = 1500
    pool, pairs = pkg.workloads.broadphase_pool(npoly, nverts, 45000, seed=23)
    off = np.arange(npoly + 1) * nverts
    gs, gd, _ = oracle_mod.Oracle("port", np.float32).gjk_epa_indexed(pool.reshape(-1, 3), pairs, off, do_epa=False, nthreads=8)
    eng = pkg.Engine(np.float32)
    desc, _keep = pkg.make_polytopes(pool)
    force_kernel("slots")
    saved = os.environ.get("OGJK_POOL_PACK")
    try:
        for pack in ("1", "0"):
            os.environ["OGJK_POOL_PACK"] = pack
            s, d = eng.compute_minimum_distance_indexed(desc, pairs)
            assert ("SoA-4" in eng.last_kernel()) == (pack == "1")
            assert np.array_equal(d, gd) and live_simplex_equal(s, gs)
    finally:
        if saved is None:
            os.environ.pop("OGJK_POOL_PACK", None)
        else:
            os.environ["OGJK_POOL_PACK"] = saved
