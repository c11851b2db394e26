"""Regenerates tests/golden/*.npz from the reference's own CPU code (oracle/_ref, built by oracle/build_ref.sh
from /root/reference).  Run in the build container only: /root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py

Each file holds the inputs (vertex arrays) and the reference's outputs (GJK distance + simplex, then EPA
distance / witnesses / normal) for one case and one precision, so the oracle restatement and the CUDA path can
both be pinned to the reference without the reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from _pkgpath import load_oracle, load_package  # noqa: E402

# examples/userP.dat / userQ.dat of the reference (config 1 fixture, 9 vertices each)
USER_P = [[0.0, 5.5, 0.0], [2.3, 1.0, -2.0], [8.1, 4.0, 2.4], [4.3, 5.0, 2.2], [2.5, 1.0, 2.3],
          [7.1, 1.0, 2.4], [1.0, 1.5, 0.3], [3.3, 0.5, 0.3], [6.0, 1.4, 0.2]]
USER_Q = [[-0.0, -5.5, 0.0], [-2.3, -1.0, 2.0], [-8.1, -4.0, -2.4], [-4.3, -5.0, -2.2], [-2.5, -1.0, -2.3],
          [-7.1, -1.0, -2.4], [-1.0, -1.5, -0.3], [-3.3, -0.5, -0.3], [-6.0, -1.4, -0.2]]


def cases(W, dtype):
    cube = W.unit_cube(dtype=dtype)
    yield "userPQ", np.asarray([USER_P], dtype), np.asarray([USER_Q], dtype)
    yield "cubes", np.stack([cube] * 4), np.stack([W.unit_cube((1, 0, 0), dtype), W.unit_cube((2, 0, 0), dtype),
                                                  W.unit_cube((5, 0, 0), dtype), W.rotated_cube_readme(dtype)])
    for name, nv, spread in (("rand64_s10", 64, 10.0), ("rand32_s1", 32, 1.0), ("rand8_s10", 8, 10.0),
                             ("rand8_s1", 8, 1.0), ("rand5_s05", 5, 0.5)):
        a, b = W.random_pairs(256, nv, spread, seed=20261017, dtype=dtype)
        yield name, a, b


def main():
    pkg = load_package()
    om = load_oracle()
    for dtype in (np.float32, np.float64):
        ref = om.Oracle("ref", dtype)
        for name, a, b in cases(pkg.workloads, dtype):
            s, d = ref.gjk(a, b)
            es, ed, en = ref.epa(a, b, s, d)
            path = os.path.join(HERE, f"{name}_{np.dtype(dtype).name}.npz")
            np.savez_compressed(path, a=a, b=b,
                                gjk_dist=d, gjk_nvrtx=s["nvrtx"], gjk_vrtx=s["vrtx"], gjk_idx=s["vrtx_idx"],
                                gjk_wit=s["witnesses"],
                                epa_dist=ed, epa_nvrtx=es["nvrtx"], epa_vrtx=es["vrtx"], epa_idx=es["vrtx_idx"],
                                epa_wit=es["witnesses"], epa_nrm=en)
            print(path, a.shape, "collide", int((d <= np.finfo(dtype).eps).sum()))


if __name__ == "__main__":
    main()
