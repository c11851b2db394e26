"""TEST INFRASTRUCTURE -- numpy model of the arithmetic of the fp16 pre-scan slot kernel
(opengjk-gpu_b200/csrc/gjk_slots16.cuh).  Not an oracle of results (the kernel's results are checked against
oracle/_ref and ogjk_oracle.c bit for bit); it restates the kernel's candidate filter so that the guarantee the kernel
relies on -- the reference's support vertex, and every vertex tying it, is always among the candidates -- can be
checked on a CPU, on the benchmark generator and on degenerate inputs (tests/test_fp16_prescan_bound.py).

  converter:  e = fl32(c - c0), c0 = first vertex;  s = 16000 / max|e|;  ch = fl16(fl32(s e));  W_j = 86.1 + 3.7e-7 s |c0_j|
  scan:       t = 2^-(exponent(max|d_j|)+1), q = |t d|, dh = fl16(t d);  a_i = fma16(ch_z, dh_z, fma16(ch_y, dh_y, fl16(ch_x dh_x)))
              thr = round_down16(max a_i - (sum_j q_j W_j (1 + 2^-12) + 0.5));  candidates = {i: a_i >= thr}
Reference semantics the candidates are checked against: GJK/cpu/openGJK.c:615-639 (maximum of the individually
rounded fp32 products summed left to right, lowest index on ties).
"""
import numpy as np

KC = np.float32(86.1)
KD = np.float32(3.7e-7)
F32 = np.float32
TINY = np.float32(2.0 ** -60)


def exact_dots(v, d):
    """reference order: (x*dx + y*dy) + z*dz, every operation rounded to fp32"""
    p = v * d[:, None, :]
    return (p[..., 0] + p[..., 1]) + p[..., 2]


def _fma16(a, b, c):  # one rounding to fp16 (float64 holds a*b + c of fp16 operands to far more than 11 bits)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float16)


def _round_down16(x):
    h = x.astype(np.float16)
    up = h.astype(np.float32) > x
    return np.where(up, np.nextafter(h, np.float16(-np.inf)), h)


def convert(a):
    """[n, nv, 3] fp32 -> (fp16 copy, W[n, 3])"""
    c0 = a[:, :1, :]
    e = (a - c0).astype(F32)
    mm = np.abs(e).max(axis=(1, 2))
    ok = mm > TINY
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        s = np.where(ok, F32(16000.0) / mm, F32(0.0)).astype(F32)
        ch = (s[:, None, None] * e).astype(F32).astype(np.float16)
        W = (KC + KD * s[:, None] * np.abs(c0[:, 0, :])).astype(F32)
    W = np.where(ok[:, None], W, F32(1e30))
    return ch, W


def candidates(ch, W, d):
    """boolean [n, nv]: the vertices the kernel re-evaluates exactly for direction d[n, 3]"""
    m = np.abs(d).max(axis=1)
    eb = (m.view(np.uint32) >> 23) & 0xff
    wide = (eb < 67) | (eb > 250)
    t = ((253 - np.where(wide, 127, eb)).astype(np.uint32) << 23).view(F32)
    td = (t[:, None] * d).astype(F32)
    q = np.abs(td)
    dh = td.astype(np.float16)
    ax = (ch[..., 0].astype(F32) * dh[:, None, 0].astype(F32)).astype(np.float16)
    az = _fma16(ch[..., 2], dh[:, None, 2], _fma16(ch[..., 1], dh[:, None, 1], ax))
    M = az.max(axis=1).astype(F32)
    with np.errstate(invalid="ignore", over="ignore"):
        slack = ((q * W).sum(axis=1) * F32(1.0 + 2.0 ** -12) + F32(0.5)).astype(F32)
        thr = _round_down16((M - slack).astype(F32))
    return (az >= thr[:, None]) | wide[:, None]


def check(a, d):
    """(guarantee holds for every row, candidate count per row)"""
    ex = exact_dots(a, d)
    best = ex.max(axis=1)
    cand = candidates(*convert(a), d)
    ties_ok = np.all(cand | (ex != best[:, None]), axis=1)  # every maximiser (hence the lowest-index one) is a candidate
    return ties_ok, cand.sum(axis=1)
