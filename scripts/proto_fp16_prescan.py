"""Design study for the round-2 slot format: can a support scan over fp16 copies of the vertices,
followed by an exact fp32 re-evaluation of the near-maximal candidates, reproduce the reference's support point
(maximum of the individually rounded fp32 dot product, lowest index on ties) -- and how many candidates does it take?

CPU only (numpy).  Directions are the ones GJK really uses: the search vectors of the first iterations of the
reference algorithm on config-2 pairs are approximated by (centre difference + noise), plus pure random directions.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()

def exact_dots(v, d):  # reference order: (x*dx + y*dy) + z*dz, every op rounded to fp32
    p = v * d[:, None, :]
    return (p[..., 0] + p[..., 1]) + p[..., 2]

def study(n, nv, spread, kind, rng):
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=777, dtype=np.float32)
    if kind == "centre":
        d = (b.mean(axis=1) - a.mean(axis=1)).astype(np.float32)
        d += rng.normal(scale=0.05, size=d.shape).astype(np.float32) * np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    else:
        d = rng.normal(size=(n, 3)).astype(np.float32)
    ex = exact_dots(a, d)
    best = ex.max(axis=1)
    want = (ex == best[:, None]).argmax(axis=1)  # lowest index among maxima
    # fp16 copies relative to the body's first vertex (keeps the magnitudes small: offsets are up to +-5, radii <= 1.5)
    # -- the maximiser of dot(v - v0, d) is the maximiser of dot(v, d) only in exact arithmetic, hence the bound below
    rel = (a - a[:, :1, :]).astype(np.float16)
    d16 = d.astype(np.float16)
    p16 = rel * d16[:, None, :]
    ap = ((p16[..., 0] + p16[..., 1]) + p16[..., 2]).astype(np.float32)
    # conservative bound on |approx - exact(rel . d)| + on the fp32 rounding of the exact evaluation itself
    mag = (np.abs(rel.astype(np.float32)) * np.abs(d)[:, None, :]).sum(axis=2).max(axis=1)
    magabs = (np.abs(a) * np.abs(d)[:, None, :]).sum(axis=2).max(axis=1)
    E = mag * np.float32(2.0 ** -8) + magabs * np.float32(2.0 ** -21)
    cand = ap >= (ap.max(axis=1) - 2 * E)[:, None]
    ok = cand[np.arange(n), want]
    # every vertex whose exact value ties the maximum must be a candidate too (tie-break needs all of them)
    ties_ok = np.all(cand | (ex != best[:, None]), axis=1)
    cnt = cand.sum(axis=1)
    print(f"{kind:7s} V={nv:3d} spread={spread:4.1f}: true support among candidates {ok.mean()*100:.4f} %  all ties covered "
          f"{ties_ok.mean()*100:.4f} %  candidates mean {cnt.mean():.2f} median {np.median(cnt):.0f} p99 {np.percentile(cnt, 99):.0f} "
          f"max {cnt.max()}  (1 candidate: {100*(cnt==1).mean():.1f} %)")

if __name__ == "__main__":
    rng = np.random.default_rng(5)
    for kind in ("centre", "random"):
        for nv, spread in ((64, 10.0), (32, 1.0), (128, 10.0)):
            study(200000, nv, spread, kind, rng)
