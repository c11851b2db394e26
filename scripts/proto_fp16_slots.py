"""CPU emulation (numpy) of the arithmetic of the fp16 pre-scan slot kernel (csrc/gjk_slots16.cuh), to check the
candidate guarantee and to count candidates before the kernel is trusted:

  converter:  e = fl32(c - c0)          c0 = first vertex of the body
              s = 16000 / max|e|        (the kernel uses the approximate reciprocal; any s with max|s e| <= 16001 will do)
              ch = fl16(fl32(s * e))
              W_j = KC + KD * s * |c0_j|
  scan:       t = 2^-(exponent(max|d_j|) + 1),  q_j = |t d_j|,  dh = fl16(t d)
              a_i = fma16(ch_z, dh_z, fma16(ch_y, dh_y, fl16(ch_x * dh_x)))
              M = max a_i, slack = (sum_j q_j W_j) * (1 + 2^-12) + 0.5, thr = round_down16(M - slack)
              candidates = {i : a_i >= thr}   (the kernel tests 8-vertex block maxima first)
  verify:     the reference's scan restricted to the candidates, in index order, strict '>'

The guarantee: the lowest-index maximiser of the individually rounded fp32 dot products, and every vertex
that ties it, is a candidate.  This script checks it on the benchmark generator and on degenerate sets and prints the
candidate statistics that size the verification loop.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()

KC = np.float32(86.1)
KD = np.float32(3.7e-7)
F32 = np.float32


def exact_dots(v, d):  # reference order: (x*dx + y*dy) + z*dz, every op rounded to fp32
    p = v * d[:, None, :]
    return (p[..., 0] + p[..., 1]) + p[..., 2]


def fma16(a, b, c):  # single rounding to fp16 (float64 holds a*b+c of fp16 operands exactly enough)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float16)


def round_down16(x):
    h = x.astype(np.float16)
    up = h.astype(np.float32) > x
    return np.where(up, np.nextafter(h, np.float16(-np.inf)), h)


def convert(a):
    c0 = a[:, :1, :]
    e = (a - c0).astype(F32)
    mm = np.abs(e).max(axis=(1, 2))
    with np.errstate(divide="ignore", over="ignore"):
        s = np.where(mm > F32(2.0 ** -60), F32(16000.0) / mm, F32(0.0)).astype(F32)
    ch = (s[:, None, None] * e).astype(F32).astype(np.float16)
    W = (KC + KD * s[:, None] * np.abs(c0[:, 0, :])).astype(F32)
    W = np.where((mm > F32(2.0 ** -60))[:, None], W, F32(np.inf))
    return ch, W


def scan(ch, W, d):
    m = np.abs(d).max(axis=1)
    ex = np.frexp(m)[1]  # m = f * 2^ex, f in [0.5, 1)
    t = np.ldexp(F32(1.0), -ex).astype(F32)
    td = (t[:, None] * d).astype(F32)
    q = np.abs(td)
    dh = td.astype(np.float16)
    ax = (ch[..., 0].astype(F32) * dh[:, None, 0].astype(F32)).astype(np.float16)
    ay = fma16(ch[..., 1], dh[:, None, 1], ax)
    az = fma16(ch[..., 2], dh[:, None, 2], ay)
    M = az.max(axis=1).astype(F32)
    slack = ((q * W).sum(axis=1) * F32(1.0 + 2.0 ** -12) + F32(0.5)).astype(F32)
    with np.errstate(invalid="ignore"):
        thr = round_down16((M - slack).astype(F32))
    thr = np.where(np.isnan(thr.astype(F32)), np.float16(-np.inf), thr)
    small = m < F32(2.0 ** -60)
    cand = (az >= thr[:, None]) | small[:, None]
    return cand


def study(name, a, d):
    n, nv, _ = a.shape
    ex = exact_dots(a, d)
    best = ex.max(axis=1)
    want = (ex == best[:, None]).argmax(axis=1)
    ch, W = convert(a)
    cand = scan(ch, W, d)
    ok = cand[np.arange(n), want]
    ties_ok = np.all(cand | (ex != best[:, None]), axis=1)
    cnt = cand.sum(axis=1)
    blocks = cand.reshape(n, nv // 8, 8).any(axis=2).sum(axis=1)
    print(f"{name:28s} V={nv:3d}: support among candidates {ok.mean()*100:.4f} %  ties covered {ties_ok.mean()*100:.4f} %  "
          f"vertices mean {cnt.mean():.3f} p99 {np.percentile(cnt, 99):.0f} max {cnt.max()}  (one: {100*(cnt==1).mean():.1f} %)  "
          f"blocks mean {blocks.mean():.3f} max {blocks.max()}")
    assert ok.all() and ties_ok.all(), "candidate guarantee violated"
    return cnt


def gjk_like_dirs(a, b, rng):
    d = (b.mean(axis=1) - a.mean(axis=1)).astype(F32)
    d += rng.normal(scale=0.05, size=d.shape).astype(F32) * np.linalg.norm(d, axis=1, keepdims=True).astype(F32)
    return d


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    n = 200000
    for nv, spread in ((64, 10.0), (32, 1.0), (48, 10.0)):
        a, b = pkg.workloads.random_pairs(n, nv, spread, seed=777, dtype=F32)
        c1 = study(f"cfg spread {spread} centre", a, gjk_like_dirs(a, b, rng))
        c2 = study(f"cfg spread {spread} random", b, rng.normal(size=(n, 3)).astype(F32))
        both = np.maximum(c1, c2)
        for lanes in (25,):
            k = (n // lanes) * lanes
            rounds = both[:k].reshape(-1, lanes).max(axis=1)
            print(f"    verification rounds per warp trip ({lanes} running lanes, one candidate per body per round): mean {rounds.mean():.2f}")
    # degenerate sets: scales, far-away tiny bodies, duplicated vertices, lattice cubes (ties), tiny directions
    a, b = pkg.workloads.random_pairs(50000, 64, 10.0, seed=3, dtype=F32)
    for sc in (1e-6, 1e-3, 1e3, 1e6):
        study(f"scale {sc:g}", (a * F32(sc)).astype(F32), rng.normal(size=(50000, 3)).astype(F32))
    tiny_far = ((a - a.mean(axis=1, keepdims=True)) * F32(1e-4) + F32(7.0)).astype(F32)
    study("tiny bodies far from origin", tiny_far, rng.normal(size=(50000, 3)).astype(F32))
    dup = a.copy(); dup[:, 32:] = dup[:, :32]
    study("duplicated vertices", dup, rng.normal(size=(50000, 3)).astype(F32))
    lat = rng.integers(-1, 2, size=(50000, 64, 3)).astype(F32)
    study("lattice {-1,0,1}^3", lat, rng.integers(-2, 3, size=(50000, 3)).astype(F32) + F32(0.0))
    study("lattice, axis directions", lat, np.tile(np.array([[0, 0, 1]], F32), (50000, 1)))
    study("tiny directions", a, (rng.normal(size=(50000, 3)) * 1e-30).astype(F32))
    study("huge directions", a, (rng.normal(size=(50000, 3)) * 1e30).astype(F32))
    pt = np.repeat(a[:, :1], 64, axis=1)
    study("point bodies", pt, rng.normal(size=(50000, 3)).astype(F32))
    print("ok")
