"""GPU parity of the fp16 pre-scan slot kernel (csrc/gjk_slots16.cuh): the support scan runs over a centred, scaled
fp16 copy of the vertices and the candidates are re-evaluated exactly, so every output must still be the reference's,
bit for bit (reference GJK/gpu/openGJK.cu:1199-1425 / GJK/cpu/openGJK.c:615-1056) -- dense batches of every supported
vertex count, the fused GJK+EPA entry, indexed batches over a pool, exact ties, and bodies at awkward scales and
distances from the origin (where the filter degrades to 'every vertex is a candidate' but must stay correct)."""
import os

import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_kernel():
    keys = ("OGJK_GJK_KERNEL", "OGJK_S16_CFG")
    saved = {k: os.environ.get(k) for k in keys}

    def setter(name, cfg=None):
        os.environ["OGJK_GJK_KERNEL"] = name
        if cfg is not None:
            os.environ["OGJK_S16_CFG"] = str(cfg)

    yield setter
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _device_batch(pkg, a, b):
    import torch
    n = a.shape[0]
    eng = pkg.Engine(np.float32)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
    return eng, d_a, d_b, d_simp, d_dist, d_nrm


def _gjk_matches(pkg, oracle_mod, a, b, nthreads=8):
    import torch
    n, nv = a.shape[0], a.shape[1]
    eng, d_a, d_b, d_simp, d_dist, _ = _device_batch(pkg, a, b)
    launches = eng.launch_count()
    eng.gjk_uniform_device(n, nv, d_a, nv, d_b, d_simp, d_dist)
    torch.cuda.synchronize()
    assert eng.launch_count() == launches + 1 and "fp16 pre-scan" in eng.last_kernel()
    os_, od = oracle_mod.Oracle("port", np.float32).gjk(a, b, nthreads=nthreads)
    gd = d_dist.cpu().numpy()
    assert np.array_equal(gd, od, equal_nan=True), f"{np.count_nonzero(gd != od)} distances differ"
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_)
    eng.set_stream(0)


@pytest.mark.parametrize("nv,spread", [(64, 10.0), (64, 1.0), (56, 6.0), (48, 3.0), (40, 10.0), (32, 1.0), (32, 10.0)])
def test_slots16_matches_oracle(pkg, oracle_mod, force_kernel, nv, spread):
    n = 40000
    a = pkg.workloads.random_polytopes(n, nv, spread, 11, np.float32, stream=1)
    b = pkg.workloads.random_polytopes(n, nv, spread, 11, np.float32, stream=2)
    force_kernel("slots16")
    _gjk_matches(pkg, oracle_mod, a, b)


@pytest.mark.parametrize("cfg", [2, 3, 5])
def test_slots16_converter_configurations(pkg, oracle_mod, force_kernel, cfg):
    """the other fetch-ring depths of the 64+64-vertex kernel"""
    n = 60000
    a, b = pkg.workloads.random_pairs(n, 64, 10.0, seed=5, dtype=np.float32)
    force_kernel("slots16", cfg)
    _gjk_matches(pkg, oracle_mod, a, b)


def test_slots16_small_and_ragged_batch_sizes(pkg, oracle_mod, force_kernel):
    """fewer pairs than slots, one pair, a count that is not a multiple of anything"""
    force_kernel("slots16")
    for n in (1, 7, 255, 257, 4097):
        a, b = pkg.workloads.random_pairs(n, 64, 4.0, seed=100 + n, dtype=np.float32)
        _gjk_matches(pkg, oracle_mod, a, b, nthreads=1)


def test_slots16_exact_ties_and_lattices(pkg, oracle_mod, force_kernel):
    """cube corners repeated four times (every support value tied at least four ways: lowest index must win) against
    shifted copies: touching, overlapping, separated"""
    W = pkg.workloads
    corners = W.unit_cube((0, 0, 0), np.float32)
    base = np.concatenate([corners] * 4)  # 32 vertices
    shifts = [(0.5, 0, 0), (2, 0, 0), (2, 2, 0), (3, 3, 3), (0, 0, 0), (2.5, 0.25, -0.5), (0, 2, 0), (1, 1, 1)]
    reps = 40000 // len(shifts)
    a = np.ascontiguousarray(np.stack([base] * (len(shifts) * reps)))
    b = np.ascontiguousarray(np.stack([np.concatenate([W.unit_cube(s, np.float32)[::-1]] * 4) for s in shifts] * reps))
    force_kernel("slots16")
    _gjk_matches(pkg, oracle_mod, a, b)


@pytest.mark.parametrize("case", ["far_small", "scale_1e-5", "scale_1e5", "mixed_scales", "first_vertex_outlier"])
def test_slots16_awkward_scales(pkg, oracle_mod, force_kernel, case):
    """inputs that stress the centring / scaling of the fp16 copy and the slack that covers the reference's own fp32
    rounding: small bodies far from the origin, tiny and huge coordinates, one body tiny and the other huge, a first
    vertex (the centre of the copy) far away from the rest of the body"""
    n = 40000
    a, b = pkg.workloads.random_pairs(n, 64, 10.0, seed=321, dtype=np.float32)
    if case == "far_small":
        a = ((a - a.mean(axis=1, keepdims=True)) * np.float32(1e-3) + np.float32(40.0)).astype(np.float32)
        b = ((b - b.mean(axis=1, keepdims=True)) * np.float32(1e-3) + np.float32(40.001)).astype(np.float32)
    elif case == "scale_1e-5":
        a, b = (a * np.float32(1e-5)).astype(np.float32), (b * np.float32(1e-5)).astype(np.float32)
    elif case == "scale_1e5":
        a, b = (a * np.float32(1e5)).astype(np.float32), (b * np.float32(1e5)).astype(np.float32)
    elif case == "mixed_scales":
        a = (a * np.float32(1e-4)).astype(np.float32)
        b = (b * np.float32(1e3)).astype(np.float32)
    else:
        a = a.copy()
        a[:, 0] += np.float32(1000.0)
    force_kernel("slots16")
    _gjk_matches(pkg, oracle_mod, np.ascontiguousarray(a), np.ascontiguousarray(b))


@pytest.mark.parametrize("nv,spread", [(64, 10.0), (64, 1.0), (48, 1.0), (32, 1.0)])
def test_slots16_fused_gjk_epa(pkg, oracle_mod, force_kernel, nv, spread):
    """gjk_epa_uniform_device: the finisher warp's fused EPA gate behind the fp16 pre-scan kernel"""
    import torch
    n = 40000
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=99, dtype=np.float32)
    eng, d_a, d_b, d_simp, d_dist, d_nrm = _device_batch(pkg, a, b)
    force_kernel("slots16")
    eng.gjk_epa_uniform_device(n, nv, d_a, nv, d_b, d_simp, d_dist, d_nrm)
    torch.cuda.synchronize()
    orc = oracle_mod.Oracle("port", np.float32)
    s, d = orc.gjk(a, b, nthreads=8)
    s, d, nr = orc.epa(a, b, s, d, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), d)
    assert np.array_equal(d_nrm.cpu().numpy(), nr)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), s)
    eng.set_stream(0)


@pytest.mark.parametrize("nverts", [32, 64])
def test_slots16_indexed_pool(pkg, oracle_mod, force_kernel, nverts):
    """uniform pool + gkCollisionPair list (BASELINE config 5 in small) through the indexed entry points"""
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
    force_kernel("slots16")
    s, d = eng.compute_minimum_distance_indexed(desc, pairs)
    assert np.array_equal(d, gd) and live_simplex_equal(s, gs)
    s3, d3, n3 = eng.compute_gjk_epa_indexed(desc, pairs)
    assert np.array_equal(d3, ed) and np.array_equal(n3, en) and live_simplex_equal(s3, es)
