import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(dtype):
    tag = np.dtype(dtype).name
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, f"*_{tag}.npz"))):
        yield os.path.basename(path)[: -len(f"_{tag}.npz")], np.load(path)


def assert_matches_golden(g, simp, dist, stage, nrm=None):
    """Bit-exact comparison of (simplices, distances[, normals]) with a golden file's `stage` ('gjk'|'epa')."""
    assert np.array_equal(dist, g[f"{stage}_dist"]), f"{stage} distance"
    assert np.array_equal(simp["nvrtx"], g[f"{stage}_nvrtx"]), f"{stage} nvrtx"
    assert np.array_equal(simp["witnesses"], g[f"{stage}_wit"], equal_nan=True), f"{stage} witnesses"
    for j in range(4):
        live = g[f"{stage}_nvrtx"] > j
        assert np.array_equal(simp["vrtx"][live, j], g[f"{stage}_vrtx"][live, j]), f"{stage} vrtx[{j}]"
        assert np.array_equal(simp["vrtx_idx"][live, j], g[f"{stage}_idx"][live, j]), f"{stage} vrtx_idx[{j}]"
    if nrm is not None:
        assert np.array_equal(nrm, g["epa_nrm"]), "contact normal"
