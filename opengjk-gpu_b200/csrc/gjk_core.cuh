// gjk_core.cuh -- per-thread (register-resident) pieces of the GJK distance algorithm:
// simplex state, the table-driven signed-volumes sub-algorithm, the exit tests and the witness
// (barycentric) stage.  No memory traffic except through the VertexFetch functor handed to the
// witness stage.  Host+device so tests/host_harness.cpp can run it against the oracle on a CPU.
//
// Behavioural contract: SURVEY.md Appendix A.3-A.5, i.e. reference GJK/gpu/openGJK.cu:288-1425
// (identical logic in GJK/cpu/openGJK.c:257-1056).
#pragma once
#include <stdint.h>

#include "gjk_math.cuh"

namespace ogjk {

// one simplex vertex: a point of the Minkowski difference and the two source vertex indices
template <typename T>
struct SV {
  V3<T> p;
  int i1, i2;
};

template <typename T>
struct Simplex {
  SV<T> s0, s1, s2, s3;  // slots 0..3; statically named so they live in registers
  int n;
};

template <typename T>
OGJK_HD SV<T> sel(bool c, const SV<T>& a, const SV<T>& b) {  // c ? a : b, field-wise selects
  SV<T> r;
  r.p.x = c ? a.p.x : b.p.x;
  r.p.y = c ? a.p.y : b.p.y;
  r.p.z = c ? a.p.z : b.p.z;
  r.i1 = c ? a.i1 : b.i1;
  r.i2 = c ? a.i2 : b.i2;
  return r;
}
template <typename T>
OGJK_HD V3<T> selv(bool c, const V3<T>& a, const V3<T>& b) {
  return mk<T>(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z);
}
template <typename T>
OGJK_HD SV<T> pick4(uint32_t code, const SV<T>& a0, const SV<T>& a1, const SV<T>& a2, const SV<T>& a3) {
  return sel((code & 2u) != 0, sel((code & 1u) != 0, a3, a2), sel((code & 1u) != 0, a1, a0));
}

// closest point to the origin on the line a + t*e, e = x - a      (projectOnLine, openGJK.cu:207-220:
// written there with pq = a - x = -e; the two negations cancel exactly)
template <typename T>
OGJK_HD V3<T> closest_on_line(const V3<T>& a, const V3<T>& e) {
  const T t = div_rn(dot(a, e), dot(e, e));
  return mk<T>(sub_rn(a.x, mul_rn(e.x, t)), sub_rn(a.y, mul_rn(e.y, t)), sub_rn(a.z, mul_rn(e.z, t)));
}
// closest point to the origin on the plane through a with normal m (projectOnPlane, openGJK.cu:222-245;
// the result is bitwise independent of the sign of m)
template <typename T>
OGJK_HD V3<T> closest_on_plane(const V3<T>& a, const V3<T>& m) {
  const T t = div_rn(dot(m, a), dot(m, m));
  return vscale(m, t);
}

// ---- sub-algorithm, 2 points (S1D, openGJK.cu:288-300) ----------------------------------------
template <typename T>
OGJK_HD void sub_1d(Simplex<T>& S, V3<T>& v) {
  const V3<T> a = S.s1.p;
  const V3<T> pp = mk<T>(mul_rn(a.x, a.x), mul_rn(a.y, a.y), mul_rn(a.z, a.z));
  if (edge_test(a, pp, S.s0.p)) {
    v = closest_on_line(a, vsub(S.s0.p, a));
  } else {
    v = a;
    S.s0 = S.s1;
    S.n = 1;
  }
}

// ---- sub-algorithm, 3 points (S2D, openGJK.cu:302-447), table driven -----------------------------
// returns true iff the newest point survives (always, for 3 points)
template <typename T>
OGJK_HD void sub_2d(Simplex<T>& S, V3<T>& v, const uint32_t* __restrict__ t2) {
  const V3<T> a = S.s2.p;
  const V3<T> pp = mk<T>(mul_rn(a.x, a.x), mul_rn(a.y, a.y), mul_rn(a.z, a.z));
  const V3<T> eb = vsub(S.s1.p, a), ec = vsub(S.s0.p, a);
  const V3<T> m = cross(eb, ec);
  uint32_t idx = 0;
  idx |= edge_test(a, pp, S.s1.p) ? 1u : 0u;
  idx |= edge_test(a, pp, S.s0.p) ? 2u : 0u;
  idx |= (dot(a, cross(eb, m)) < T(0)) ? 4u : 0u;  // hff2(a,b,c)
  idx |= (dot(a, cross(ec, m)) > T(0)) ? 8u : 0u;  // hff2(a,c,b): cross(ec, cross(ec,eb)) = -cross(ec, m)
  const uint32_t leaf = t2[idx];
  const uint32_t kind = (leaf >> 12) & 7u;
  if (kind == 2u) {
    v = closest_on_plane(a, m);
  } else if (kind == 1u) {
    v = closest_on_line(a, selv(((leaf >> 16) & 3u) == 1u, eb, ec));
  } else {
    v = a;
  }
  const SV<T> o0 = pick4((leaf >> 4) & 3u, S.s0, S.s1, S.s2, S.s2);
  const SV<T> o1 = pick4((leaf >> 6) & 3u, S.s0, S.s1, S.s2, S.s2);
  S.s0 = o0;
  S.s1 = o1;
  S.n = (int)(leaf & 7u);
}

// ---- sub-algorithm, 4 points (S3D, openGJK.cu:449-827), table driven -----------------------------
// returns true iff the newest point survives
template <typename T>
OGJK_HD bool sub_3d(Simplex<T>& S, V3<T>& v, const uint32_t* __restrict__ t3) {
  const V3<T> a = S.s3.p;
  const V3<T> pp = mk<T>(mul_rn(a.x, a.x), mul_rn(a.y, a.y), mul_rn(a.z, a.z));
  const V3<T> e0 = vsub(S.s0.p, a), e1 = vsub(S.s1.p, a), e2 = vsub(S.s2.p, a);
  uint32_t idx = 0;
  idx |= edge_test(a, pp, S.s0.p) ? 1u : 0u;
  idx |= edge_test(a, pp, S.s1.p) ? 2u : 0u;
  idx |= edge_test(a, pp, S.s2.p) ? 4u : 0u;
  // sss = det(s1s3, s1s4, s1s2) <= 0 with s3 = slot1, s4 = slot0, s2 = slot2 (openGJK.cu:345-346)
  const bool sss = det3(e1, e0, e2) <= T(0);
  // hff3(a, q, r) = dot(a, q x r) <= 0 for the faces omitting slot 2, 1, 0
  const bool f2 = dot(a, cross(S.s1.p, S.s0.p)) <= T(0);
  const bool f1 = dot(a, cross(S.s0.p, S.s2.p)) <= T(0);
  const bool f0 = dot(a, cross(S.s2.p, S.s1.p)) <= T(0);
  idx |= (f0 != sss) ? 8u : 0u;
  idx |= (f1 != sss) ? 16u : 0u;
  idx |= (f2 != sss) ? 32u : 0u;
  // hff2(a,x,y) = dot(a, ex x (ex x ey)) < 0; (ex x ey) = -(ey x ex) exactly, so three first-level
  // cross products serve all six ordered pairs.
  const V3<T> m01 = cross(e0, e1), m02 = cross(e0, e2), m12 = cross(e1, e2);
  idx |= (dot(a, cross(e0, m01)) < T(0)) ? (1u << 6) : 0u;   // (0,1)
  idx |= (dot(a, cross(e1, m01)) > T(0)) ? (1u << 7) : 0u;   // (1,0)
  idx |= (dot(a, cross(e0, m02)) < T(0)) ? (1u << 8) : 0u;   // (0,2)
  idx |= (dot(a, cross(e2, m02)) > T(0)) ? (1u << 9) : 0u;   // (2,0)
  idx |= (dot(a, cross(e1, m12)) < T(0)) ? (1u << 10) : 0u;  // (1,2)
  idx |= (dot(a, cross(e2, m12)) > T(0)) ? (1u << 11) : 0u;  // (2,1)
  const uint32_t leaf = t3[idx];
  const uint32_t kind = (leaf >> 12) & 7u;
  const uint32_t x = (leaf >> 16) & 3u;
  if (kind == 2u) {
    v = closest_on_plane(a, selv(x == 0u, m01, selv(x == 1u, m02, m12)));
  } else if (kind == 1u) {
    v = closest_on_line(a, selv(x == 0u, e0, selv(x == 1u, e1, e2)));
  } else if (kind == 0u) {
    v = a;
  } else if (kind == 3u) {
    v = mk<T>(T(0), T(0), T(0));
  }
  const SV<T> o0 = pick4((leaf >> 4) & 3u, S.s0, S.s1, S.s2, S.s3);
  const SV<T> o1 = pick4((leaf >> 6) & 3u, S.s0, S.s1, S.s2, S.s3);
  const SV<T> o2 = pick4((leaf >> 8) & 3u, S.s0, S.s1, S.s2, S.s3);
  S.s0 = o0;
  S.s1 = o1;
  S.s2 = o2;
  S.n = (int)(leaf & 7u);
  return ((leaf >> 20) & 1u) != 0u;
}

// ---- per-pair GJK state and one iteration's scalar part ---------------------------------------
template <typename T>
struct GjkState {
  Simplex<T> S;
  V3<T> v;
  V3<T> sup1, sup2;  // current support points of body 1 / body 2 (gkPolytope::s)
  int idx1, idx2;    // ... and their vertex indices          (gkPolytope::s_idx)
  T norm2_wmax;
  int k;
};

// openGJK.cu:1291-1322: v = first vertex of body 1 minus first vertex of body 2
template <typename T>
OGJK_HD void gjk_init(GjkState<T>& g, const V3<T>& p0, const V3<T>& q0) {
  g.v = vsub(p0, q0);
  g.S.n = 1;
  g.S.s0.p = g.v;
  g.S.s0.i1 = 0;
  g.S.s0.i2 = 0;
  g.S.s1 = g.S.s0;
  g.S.s2 = g.S.s0;
  g.S.s3 = g.S.s0;
  g.sup1 = p0;
  g.sup2 = q0;
  g.idx1 = 0;
  g.idx2 = 0;
  g.norm2_wmax = T(0);
  g.k = 0;
}

// Everything of one GJK iteration after the two support searches (openGJK.cu:1355-1410).
// Returns true when the loop must stop.
template <typename T>
OGJK_HD bool gjk_advance(GjkState<T>& g, const uint32_t* __restrict__ t2, const uint32_t* __restrict__ t3) {
  const T eps_rel = Tol<T>::eps_rel();
  const T eps_tot = Tol<T>::eps_tot();
  const V3<T> w = vsub(g.sup1, g.sup2);
  const T vv = norm2(g.v);
  const T gap = sub_rn(vv, dot(g.v, w));
  if (gap <= mul_rn(eps_rel, vv) || gap < eps_tot) return true;
  if (vv < mul_rn(eps_rel, eps_rel)) return true;

  SV<T> nw;
  nw.p = w;
  nw.i1 = g.idx1;
  nw.i2 = g.idx2;
  bool keeps = true;
  if (g.S.n == 1) {
    g.S.s1 = nw;
    g.S.n = 2;
    sub_1d(g.S, g.v);
    // the initial point p0-q0 was never "added": account for it here if it survived
    if (g.S.n == 2) {
      const T n0 = norm2(g.S.s0.p);
      if (n0 > g.norm2_wmax) g.norm2_wmax = n0;
    }
  } else if (g.S.n == 2) {
    g.S.s2 = nw;
    g.S.n = 3;
    sub_2d(g.S, g.v, t2);
  } else {
    g.S.s3 = nw;
    g.S.n = 4;
    keeps = sub_3d(g.S, g.v, t3);
  }
  // running max of |vertex|^2 over surviving vertices (openGJK.cu:1396-1402): older survivors were
  // accounted for when they were added, so only the newest point can raise it.
  if (keeps) {
    const T nn = norm2(w);
    if (nn > g.norm2_wmax) g.norm2_wmax = nn;
  }
  if (norm2(g.v) <= mul_rn(mul_rn(eps_tot, eps_tot), g.norm2_wmax)) return true;
  return g.S.n == 4 || g.k == 25;
}

// ---- lane-uniform iteration (persistent slot kernel) --------------------------------------------
// Same arithmetic as gjk_advance, arranged so that threads holding 2-, 3- and 4-point simplices execute ONE
// instruction stream: the newest point `a` is always treated as slot 3, the older points stay in slots 0..m-1,
// all 12 predicate bits of the 4-point index are evaluated (bits that involve absent slots are garbage computed
// from stale but finite slot contents and are masked or ignored by the table, gjk_tables.h "unified table"), the
// closest-point division is shared between the line and the plane case (dot(a,e) and dot(m,a) multiply the same
// operands), and the surviving points are picked out of {slot0, slot1, slot2, a} by the leaf's permutation.
// In a thread-per-pair warp this replaces three serialised divergent paths by a single one.
// `tab` = the 16-bit unified table (any address space).
// The two exit tests that precede the sub-algorithm (openGJK.cu:1363-1378): true = the pair has terminated with the
// simplex it already holds.  Split out so that the slot kernel can hand a finished pair's slot back for refilling
// before it runs the sub-algorithm of the others.
template <typename T>
OGJK_HD bool gjk_converged_u(const GjkState<T>& g) {
  const T eps_rel = Tol<T>::eps_rel();
  const T eps_tot = Tol<T>::eps_tot();
  const V3<T> a = vsub(g.sup1, g.sup2);
  const T vv = norm2(g.v);
  const T gap = sub_rn(vv, dot(g.v, a));
  if (gap <= mul_rn(eps_rel, vv) || gap < eps_tot) return true;
  return vv < mul_rn(eps_rel, eps_rel);
}
template <typename T>
OGJK_HD bool gjk_substep_u(GjkState<T>& g, const uint16_t* __restrict__ tab);
template <typename T>
OGJK_HD bool gjk_advance_u(GjkState<T>& g, const uint16_t* __restrict__ tab) {
  if (gjk_converged_u(g)) return true;
  return gjk_substep_u(g, tab);
}
// everything after the two pre-tests: add the new point, sub-algorithm, running max, third exit test
template <typename T>
OGJK_HD bool gjk_substep_u(GjkState<T>& g, const uint16_t* __restrict__ tab) {
  const T eps_tot = Tol<T>::eps_tot();
  const V3<T> a = vsub(g.sup1, g.sup2);

  const int m = g.S.n;  // 1..3 older points
  const V3<T> pp = mk<T>(mul_rn(a.x, a.x), mul_rn(a.y, a.y), mul_rn(a.z, a.z));
  const V3<T> p0 = g.S.s0.p, p1 = g.S.s1.p, p2 = g.S.s2.p;
  const V3<T> e0 = vsub(p0, a), e1 = vsub(p1, a), e2 = vsub(p2, a);
  uint32_t idx = 0;
  idx |= edge_test(a, pp, p0) ? 1u : 0u;
  idx |= edge_test(a, pp, p1) ? 2u : 0u;
  idx |= edge_test(a, pp, p2) ? 4u : 0u;
  const bool sss = det3(e1, e0, e2) <= T(0);
  const bool f2 = dot(a, cross(p1, p0)) <= T(0);
  const bool f1 = dot(a, cross(p0, p2)) <= T(0);
  const bool f0 = dot(a, cross(p2, p1)) <= T(0);
  idx |= (f0 != sss) ? 8u : 0u;
  idx |= (f1 != sss) ? 16u : 0u;
  idx |= (f2 != sss) ? 32u : 0u;
  const V3<T> m01 = cross(e0, e1), m02 = cross(e0, e2), m12 = cross(e1, e2);
  idx |= (dot(a, cross(e0, m01)) < T(0)) ? (1u << 6) : 0u;
  idx |= (dot(a, cross(e1, m01)) > T(0)) ? (1u << 7) : 0u;
  idx |= (dot(a, cross(e0, m02)) < T(0)) ? (1u << 8) : 0u;
  idx |= (dot(a, cross(e2, m02)) > T(0)) ? (1u << 9) : 0u;
  idx |= (dot(a, cross(e1, m12)) < T(0)) ? (1u << 10) : 0u;
  idx |= (dot(a, cross(e2, m12)) > T(0)) ? (1u << 11) : 0u;
  const uint32_t mask = m == 3 ? 0xfffu : (m == 2 ? 0xffu : 0x1u);
  const uint32_t base = m == 3 ? 0u : (m == 2 ? 4096u : 4352u);
  const uint32_t leaf = tab[base + (idx & mask)];
  const uint32_t kind = (leaf >> 10) & 7u;
  const uint32_t x = (leaf >> 13) & 3u;

  // closest point on the line (a, slot x) or on the plane of face x; one shared division
  const V3<T> el = selv(x == 0u, e0, selv(x == 1u, e1, e2));
  const V3<T> ml = selv(x == 0u, m01, selv(x == 1u, m02, m12));
  const bool plane = kind == 2u;
  const V3<T> q = selv(plane, ml, el);
  const T t = div_rn(dot(q, a), dot(q, q));
  const V3<T> qt = vscale(q, t);
  V3<T> nv = selv(plane, qt, vsub(a, qt));
  nv = selv(kind == 0u, a, nv);
  nv = selv(kind == 3u, mk<T>(T(0), T(0), T(0)), nv);
  g.v = selv(kind == 4u, g.v, nv);

  SV<T> nw;
  nw.p = a;
  nw.i1 = g.idx1;
  nw.i2 = g.idx2;
  const T n0 = norm2(p0);
  const SV<T> o0 = pick4((leaf >> 2) & 3u, g.S.s0, g.S.s1, g.S.s2, nw);
  const SV<T> o1 = pick4((leaf >> 4) & 3u, g.S.s0, g.S.s1, g.S.s2, nw);
  const SV<T> o2 = pick4((leaf >> 6) & 3u, g.S.s0, g.S.s1, g.S.s2, nw);
  g.S.s0 = o0;
  g.S.s1 = o1;
  g.S.s2 = o2;
  g.S.s3 = nw;  // slot 3 is only live when all four points survive (identity permutation)
  g.S.n = (int)(leaf & 3u) + 1;
  // running max of |vertex|^2 (see gjk_advance): the initial point is accounted for when it survives the first
  // iteration; afterwards only the newest point can raise the maximum.
  if (m == 1 && g.S.n == 2 && n0 > g.norm2_wmax) g.norm2_wmax = n0;
  if ((leaf >> 15) & 1u) {
    const T nn = norm2(a);
    if (nn > g.norm2_wmax) g.norm2_wmax = nn;
  }
  if (norm2(g.v) <= mul_rn(mul_rn(eps_tot, eps_tot), g.norm2_wmax)) return true;
  return g.S.n == 4 || g.k == 25;
}

// ---- witnesses (openGJK.cu:884-1189) ----------------------------------------------------------
// Fetch must provide: V3<T> operator()(int body /*0|1*/, int vertex_index) const
template <typename T, typename Fetch>
OGJK_HD void blend_witness(const Fetch& fetch, const Simplex<T>& S, int m, T a0, T a1, T a2, T a3, V3<T>& w1,
                           V3<T>& w2) {
  const V3<T> p0 = fetch(0, S.s0.i1), q0 = fetch(1, S.s0.i2);
  const V3<T> p1 = fetch(0, S.s1.i1), q1 = fetch(1, S.s1.i2);
  w1 = mk<T>(add_rn(mul_rn(p0.x, a0), mul_rn(p1.x, a1)), add_rn(mul_rn(p0.y, a0), mul_rn(p1.y, a1)),
             add_rn(mul_rn(p0.z, a0), mul_rn(p1.z, a1)));
  w2 = mk<T>(add_rn(mul_rn(q0.x, a0), mul_rn(q1.x, a1)), add_rn(mul_rn(q0.y, a0), mul_rn(q1.y, a1)),
             add_rn(mul_rn(q0.z, a0), mul_rn(q1.z, a1)));
  if (m >= 3) {
    const V3<T> p2 = fetch(0, S.s2.i1), q2 = fetch(1, S.s2.i2);
    w1 = mk<T>(add_rn(w1.x, mul_rn(p2.x, a2)), add_rn(w1.y, mul_rn(p2.y, a2)), add_rn(w1.z, mul_rn(p2.z, a2)));
    w2 = mk<T>(add_rn(w2.x, mul_rn(q2.x, a2)), add_rn(w2.y, mul_rn(q2.y, a2)), add_rn(w2.z, mul_rn(q2.z, a2)));
  }
  if (m >= 4) {
    const V3<T> p3 = fetch(0, S.s3.i1), q3 = fetch(1, S.s3.i2);
    w1 = mk<T>(add_rn(w1.x, mul_rn(p3.x, a3)), add_rn(w1.y, mul_rn(p3.y, a3)), add_rn(w1.z, mul_rn(p3.z, a3)));
    w2 = mk<T>(add_rn(w2.x, mul_rn(q3.x, a3)), add_rn(w2.y, mul_rn(q3.y, a3)), add_rn(w2.z, mul_rn(q3.z, a3)));
  }
}

// a0 = (T)(1.0 - a1 [- a2 [- a3]]) with the subtractions carried out in double (openGJK.cu:917, 987, 1113)
OGJK_HD float one_minus(float a1) { return (float)(1.0 - (double)a1); }
OGJK_HD float one_minus(float a1, float a2) { return (float)(1.0 - (double)a1 - (double)a2); }
OGJK_HD float one_minus(float a1, float a2, float a3) {
  return (float)(1.0 - (double)a1 - (double)a2 - (double)a3);
}
OGJK_HD double one_minus(double a1) { return sub_rn(1.0, a1); }
OGJK_HD double one_minus(double a1, double a2) { return sub_rn(sub_rn(1.0, a1), a2); }
OGJK_HD double one_minus(double a1, double a2, double a3) { return sub_rn(sub_rn(sub_rn(1.0, a1), a2), a3); }

template <typename T, typename Fetch>
OGJK_HD void wit_1d(const Fetch& fetch, Simplex<T>& S, V3<T>& w1, V3<T>& w2) {
  const V3<T> p = S.s0.p;
  const V3<T> pq = vsub(S.s1.p, p), po = vneg(p);
  const T det = dot(pq, pq);
  // det == 0 calls the 0-D routine first and then falls through (openGJK.cu:911-916); the fall-through
  // overwrites its result, so only the blend below is observable.
  const T a1 = div_rn(dot(pq, po), det);
  const T a0 = one_minus(a1);
  blend_witness(fetch, S, 2, a0, a1, T(0), T(0), w1, w2);
}

// barycentric weights of the origin in the triangle S.s0,S.s1,S.s2 (openGJK.cu:955-987)
template <typename T>
OGJK_HD void tri_weights(const Simplex<T>& S, T& a0, T& a1, T& a2) {
  const V3<T> p = S.s0.p;
  const V3<T> pq = vsub(S.s1.p, p), pr = vsub(S.s2.p, p), po = vneg(p);
  const T T00 = dot(pq, pq), T01 = dot(pq, pr), T11 = dot(pr, pr);
  const T det = sub_rn(mul_rn(T00, T11), mul_rn(T01, T01));
  const T b0 = dot(pq, po), b1 = dot(pr, po);
  const T I00 = div_rn(T11, det), I01 = div_rn(-T01, det), I11 = div_rn(T00, det);
  a1 = add_rn(mul_rn(I00, b0), mul_rn(I01, b1));
  a2 = add_rn(mul_rn(I01, b0), mul_rn(I11, b1));
  a0 = one_minus(a1, a2);
}
// near-edge demotion of a triangle (openGJK.cu:991-1010): shrinks/reorders the stored simplex.  The 1-D
// witnesses the reference computes at this point are always overwritten by the caller's blend
// (fall-through, :1012-1027), so only the simplex change is observable.  (A zero determinant calls the
// 1-D routine first, :975-978, which has no side effect on the simplex.)
template <typename T>
OGJK_HD void tri_demote(Simplex<T>& S, T a0, T a1, T a2) {
  const T eps = Tol<T>::eps();
  if (a0 < eps) {
    S.n = 2;
    S.s0 = S.s2;
  } else if (a1 < eps) {
    S.n = 2;
    S.s1 = S.s2;
  } else if (a2 < eps) {
    S.n = 2;
  }
}

template <typename T, typename Fetch>
OGJK_HD void wit_2d(const Fetch& fetch, Simplex<T>& S, V3<T>& w1, V3<T>& w2) {
  T a0, a1, a2;
  tri_weights(S, a0, a1, a2);
  tri_demote(S, a0, a1, a2);
  blend_witness(fetch, S, 3, a0, a1, a2, T(0), w1, w2);
}

template <typename T, typename Fetch>
OGJK_HD void wit_3d(const Fetch& fetch, Simplex<T>& S, V3<T>& w1, V3<T>& w2) {
  const V3<T> p = S.s0.p;
  const V3<T> pq = vsub(S.s1.p, p), pr = vsub(S.s2.p, p), ps = vsub(S.s3.p, p), po = vneg(p);
  const T T00 = dot(pq, pq), T01 = dot(pq, pr), T02 = dot(pq, ps);
  const T T11 = dot(pr, pr), T12 = dot(pr, ps), T22 = dot(ps, ps);
  const T det00 = sub_rn(mul_rn(T11, T22), mul_rn(T12, T12));
  const T det01 = sub_rn(mul_rn(T01, T22), mul_rn(T02, T12));
  const T det02 = sub_rn(mul_rn(T01, T12), mul_rn(T02, T11));
  const T det = add_rn(sub_rn(mul_rn(T00, det00), mul_rn(T01, det01)), mul_rn(T02, det02));
  const T b0 = dot(pq, po), b1 = dot(pr, po), b2 = dot(ps, po);
  const T det11 = sub_rn(mul_rn(T00, T22), mul_rn(T02, T02));
  const T det12 = sub_rn(mul_rn(T00, T12), mul_rn(T01, T02));
  const T det22 = sub_rn(mul_rn(T00, T11), mul_rn(T01, T01));
  const T I00 = div_rn(det00, det), I01 = div_rn(-det01, det), I02 = div_rn(det02, det);
  const T I11 = div_rn(det11, det), I12 = div_rn(-det12, det), I22 = div_rn(det22, det);
  const T a1 = add_rn(add_rn(mul_rn(I00, b0), mul_rn(I01, b1)), mul_rn(I02, b2));
  const T a2 = add_rn(add_rn(mul_rn(I01, b0), mul_rn(I11, b1)), mul_rn(I12, b2));
  const T a3 = add_rn(add_rn(mul_rn(I02, b0), mul_rn(I12, b1)), mul_rn(I22, b2));
  const T a0 = one_minus(a1, a2, a3);
  // a zero determinant runs the 2-D routine on the first three slots before falling through
  // (openGJK.cu:1089-1092); its witnesses are overwritten below but its demotion of the simplex stays.
  if (det == T(0)) {
    T c0, c1, c2;
    tri_weights(S, c0, c1, c2);
    tri_demote(S, c0, c1, c2);
  }
  // near-face demotion (openGJK.cu:1117-1147) followed by the 2-D routine, which may demote again; all
  // witnesses computed on the way are overwritten by the 4-weight blend below (:1149-1166).
  const T eps = Tol<T>::eps();
  bool demoted = true;
  if (a0 < eps) S.s0 = S.s3;
  else if (a1 < eps) S.s1 = S.s3;
  else if (a2 < eps) S.s2 = S.s3;
  else if (a3 < eps) { /* slots stay */ }
  else demoted = false;
  if (demoted) {
    S.n = 3;
    T c0, c1, c2;
    tri_weights(S, c0, c1, c2);
    tri_demote(S, c0, c1, c2);
  }
  blend_witness(fetch, S, 4, a0, a1, a2, a3, w1, w2);
}

template <typename T, typename Fetch>
OGJK_HD void gjk_witnesses(const Fetch& fetch, Simplex<T>& S, V3<T>& w1, V3<T>& w2) {
  if (S.n == 4) {
    wit_3d(fetch, S, w1, w2);
  } else if (S.n == 3) {
    wit_2d(fetch, S, w1, w2);
  } else if (S.n == 2) {
    wit_1d(fetch, S, w1, w2);
  } else {
    w1 = fetch(0, S.s0.i1);
    w2 = fetch(1, S.s0.i2);
  }
}

}  // namespace ogjk
