"""Pins the C restatement (oracle/ogjk_oracle.c) to the reference: golden vectors produced by the reference's
own CPU code, README known answers, and -- when oracle/_ref is present -- a live bit-for-bit comparison."""
import numpy as np
import pytest

from conftest import live_simplex_equal
from golden_util import assert_matches_golden, golden_cases

DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_port_reproduces_golden_vectors(oracle_mod, dtype):
    orc = oracle_mod.Oracle("port", dtype)
    seen = 0
    for name, g in golden_cases(dtype):
        s, d = orc.gjk(g["a"], g["b"])
        assert_matches_golden(g, s, d, "gjk")
        es, ed, en = orc.epa(g["a"], g["b"], s, d)
        assert_matches_golden(g, es, ed, "epa", en)
        seen += 1
    assert seen >= 7


@pytest.mark.parametrize("dtype", DTYPES)
def test_readme_known_answers(oracle_mod, dtype):
    """reference README.md:111-115 (GJK on userP/userQ) and :131-137 (EPA, rotated cube)."""
    cases = dict(golden_cases(dtype))
    g = cases["userPQ"]
    assert f"{float(g['gjk_dist'][0]):.6f}" == "3.653650"
    w = g["gjk_wit"][0]
    assert [f"{x:.6f}" for x in w[0]] == ["1.025173", "1.490318", "0.255463"]
    assert [f"{x:.6f}" for x in w[1]] == ["-1.025173", "-1.490318", "-0.255463"]
    c = cases["cubes"]
    assert f"{-float(c['epa_dist'][3]):.6f}" == "1.500000"
    assert [f"{x:.6f}" for x in c["epa_wit"][3][0]] == ["1.000000", "0.500000", "0.707107"]
    assert [f"{x:.6f}" for x in c["epa_wit"][3][1]] == ["-0.500000", "0.500000", "0.707107"]
    assert [f"{abs(x):.6f}" for x in c["epa_nrm"][3]] == ["1.000000", "0.000000", "0.000000"]
    # shifted cubes
    assert c["gjk_dist"][0] == 0 and c["epa_dist"][0] == -1 and tuple(c["epa_nrm"][0]) == (1, 0, 0)
    assert c["gjk_dist"][2] == 3 and c["epa_dist"][2] == 3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("nverts,spread", [(64, 10.0), (32, 1.0), (8, 10.0), (8, 1.0), (300, 3.0), (4, 2.0), (1, 3.0)])
def test_port_equals_compiled_reference(oracle_mod, pkg, dtype, nverts, spread):
    if not oracle_mod.available("ref", dtype):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    n = 20000 if nverts <= 64 else 2000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=31337, dtype=dtype)
    port, ref = oracle_mod.Oracle("port", dtype), oracle_mod.Oracle("ref", dtype)
    s1, d1 = port.gjk(a, b)
    s2, d2 = ref.gjk(a, b)
    assert np.array_equal(d1, d2)
    assert live_simplex_equal(s1, s2)
    # the unused slots keep the algorithm's history in both implementations: compare those too
    assert np.array_equal(s1["vrtx"], s2["vrtx"]) and np.array_equal(s1["vrtx_idx"], s2["vrtx_idx"])
    e1 = port.epa(a, b, s1, d1)
    e2 = ref.epa(a, b, s2, d2)
    assert np.array_equal(e1[1], e2[1]) and np.array_equal(e1[2], e2[2])
    assert live_simplex_equal(e1[0], e2[0])


@pytest.mark.parametrize("dtype", DTYPES)
def test_ragged_and_indexed_drivers(oracle_mod, pkg, dtype):
    rng = np.random.default_rng(3)
    counts = rng.integers(1, 40, size=500)
    pool = [pkg.workloads.random_polytopes(1, int(c), 4.0, 50 + i, dtype)[0] for i, c in enumerate(counts)]
    off = np.concatenate([[0], np.cumsum(counts)])
    flat = np.concatenate(pool)
    pairs = rng.integers(0, 500, size=(2000, 2)).astype(np.int32)
    orc = oracle_mod.Oracle("port", dtype)
    s, d, nrm = orc.gjk_epa_indexed(flat, pairs, off)
    # same pairs through the non-indexed driver
    a = [pool[i] for i in pairs[:, 0]]
    b = [pool[i] for i in pairs[:, 1]]
    offa = np.concatenate([[0], np.cumsum([len(x) for x in a])])
    offb = np.concatenate([[0], np.cumsum([len(x) for x in b])])
    s2, d2 = orc.gjk(np.concatenate(a), np.concatenate(b), offa, offb)
    s2, d2, n2 = orc.epa(np.concatenate(a), np.concatenate(b), s2, d2, offa, offb)
    assert np.array_equal(d, d2) and np.array_equal(nrm, n2) and live_simplex_equal(s, s2)
    if oracle_mod.available("ref", dtype):
        s3, d3, n3 = oracle_mod.Oracle("ref", dtype).gjk_epa_indexed(flat, pairs, off)
        assert np.array_equal(d, d3) and np.array_equal(nrm, n3) and live_simplex_equal(s, s3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("nv", [8, 32])
def test_port_equals_reference_on_degenerate_geometry(oracle_mod, dtype, nv):
    """duplicated vertices, point / segment / planar bodies, lattice cubes, identical bodies, tiny and large scales
    (generator shared with tests/test_gpu_degenerate.py).  EPA is compared at scale 1e3: at 1e6 the reference's EPA
    support search falls below its -1e10 sentinel and continues with an uninitialised point (GJK/cpu/EPA.c:311-344),
    so its output there is not a function of its input."""
    if not oracle_mod.available("ref", dtype):
        pytest.skip("oracle/_ref not built")
    from test_gpu_degenerate import degenerate_pairs
    port, ref = oracle_mod.Oracle("port", dtype), oracle_mod.Oracle("ref", dtype)
    a, b, _cat = degenerate_pairs(8000, nv, seed=3 + nv, dtype=dtype)
    s1, d1 = port.gjk(a, b, nthreads=4)
    s2, d2 = ref.gjk(a, b, nthreads=4)
    assert np.array_equal(d1, d2, equal_nan=True) and live_simplex_equal(s1, s2)
    a, b, _cat = degenerate_pairs(8000, nv, seed=3 + nv, dtype=dtype, large=1e3)
    s1, d1 = port.gjk(a, b, nthreads=4)
    s2, d2 = ref.gjk(a, b, nthreads=4)
    e1, e2 = port.epa(a, b, s1, d1, nthreads=4), ref.epa(a, b, s2, d2, nthreads=4)
    assert np.array_equal(e1[1], e2[1], equal_nan=True) and np.array_equal(e1[2], e2[2], equal_nan=True)
    assert np.array_equal(e1[0]["witnesses"], e2[0]["witnesses"], equal_nan=True)
