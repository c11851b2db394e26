/*
 * oracle/ogjk_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the reference's GJK distance + EPA penetration path, written
 * from the behavioural contract in SURVEY.md Appendix A.  It is the checker the CUDA kernels
 * are compared against; it is never linked into, imported by, or called from the product
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against
 *   (1) the reference's own CPU sources compiled unmodified into oracle/_ref/ (build_ref.sh),
 *   (2) the golden vectors under tests/golden/ (generated from (1) by tests/golden/make_golden.py),
 *   (3) the README known answers (reference README.md:111-115, 131-137).
 *
 * Every arithmetic expression follows the reference's evaluation order (IEEE, no FMA
 * contraction: build with -ffp-contract=off); each function cites the reference lines it
 * restates.  Compile twice: default = fp32 (reference USE_32BITS, GJK/common.h:44-52),
 * -DOGJK_ORACLE_F64 = fp64 (GJK/common.h:53-60).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <string.h>

#ifdef OGJK_ORACLE_F64
typedef double real;
#define R_EPS DBL_EPSILON
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMAX fmax
#define R_FMIN fmin
#else
typedef float real;
#define R_EPS FLT_EPSILON
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FMAX fmaxf
#define R_FMIN fminf
#endif

/* Byte-compatible with gkSimplex (reference GJK/common.h:84-89). */
typedef struct {
  int nvrtx;
  real vrtx[4][3];
  int vrtx_idx[4][2];
  real witnesses[2][3];
} osimplex;

typedef struct {
  const real* xyz; /* flattened x0 y0 z0 x1 ... (reference GJK/gpu/openGJK.h:88-89) */
  int n;
} obody;

/* one simplex vertex with its provenance (vertex index on body 1, body 2) */
typedef struct {
  real p[3];
  int id[2];
} svert;

/* ------------------------------------------------------------------------------------------
 * A.1 primitives (reference GJK/cpu/openGJK.c:44-45, 167-255)
 * ---------------------------------------------------------------------------------------- */
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline real nrm2(const real* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
static inline void cross3(const real* a, const real* b, real* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void sub3(const real* a, const real* b, real* c) {
  c[0] = a[0] - b[0];
  c[1] = a[1] - b[1];
  c[2] = a[2] - b[2];
}
/* openGJK.c:167-173 */
static inline real det3(const real* p, const real* q, const real* r) {
  return p[0] * ((q[1] * r[2]) - (r[1] * q[2])) - p[1] * (q[0] * r[2] - r[0] * q[2]) +
         p[2] * (q[0] * r[1] - r[0] * q[1]);
}
/* openGJK.c:183-196: closest point to the origin on the line through p and q */
static inline void origin_on_line(const real* p, const real* q, real* v) {
  real pq[3];
  sub3(p, q, pq);
  const real t = dot3(p, pq) / dot3(pq, pq);
  for (int i = 0; i < 3; ++i) v[i] = p[i] - pq[i] * t;
}
/* openGJK.c:198-217: closest point to the origin on the plane through p, q, r */
static inline void origin_on_plane(const real* p, const real* q, const real* r, real* v) {
  real n[3], pq[3], pr[3];
  sub3(p, q, pq);
  sub3(p, r, pr);
  cross3(pq, pr, n);
  const real t = dot3(n, p) / dot3(n, n);
  for (int i = 0; i < 3; ++i) v[i] = n[i] * t;
}
/* openGJK.c:219-230 (hff1): does the origin project onto the open ray p->q past p? */
static inline int edge_test(const real* p, const real* q) {
  real acc = 0;
  for (int i = 0; i < 3; ++i) acc += (p[i] * p[i] - p[i] * q[i]);
  return acc > 0;
}
/* openGJK.c:232-248 (hff2): true => r can be discarded w.r.t. edge pq */
static inline int face_edge_test(const real* p, const real* q, const real* r) {
  real pq[3], pr[3], m[3], n[3];
  sub3(q, p, pq);
  sub3(r, p, pr);
  cross3(pq, pr, m);
  cross3(pq, m, n);
  return dot3(p, n) < 0;
}
/* openGJK.c:250-255 (hff3) */
static inline int plane_test(const real* p, const real* q, const real* r) {
  real n[3];
  cross3(q, r, n);
  return dot3(p, n) <= 0;
}

/* ------------------------------------------------------------------------------------------
 * slot bookkeeping
 * ---------------------------------------------------------------------------------------- */
static inline void get_slot(const osimplex* s, int k, svert* o) {
  for (int t = 0; t < 3; ++t) o->p[t] = s->vrtx[k][t];
  o->id[0] = s->vrtx_idx[k][0];
  o->id[1] = s->vrtx_idx[k][1];
}
static inline void put_slot(osimplex* s, int k, const svert* o) {
  for (int t = 0; t < 3; ++t) s->vrtx[k][t] = o->p[t];
  s->vrtx_idx[k][0] = o->id[0];
  s->vrtx_idx[k][1] = o->id[1];
}
static inline void set3(real* v, const real* a) {
  v[0] = a[0];
  v[1] = a[1];
  v[2] = a[2];
}

/* ------------------------------------------------------------------------------------------
 * A.4 signed-volumes sub-algorithm
 * ---------------------------------------------------------------------------------------- */
/* 2 points: openGJK.c:257-268 (+ S1Dregion1 :118-127) */
static void sub_1d(osimplex* s, real* v) {
  svert newest, old;
  get_slot(s, 1, &newest);
  get_slot(s, 0, &old);
  if (edge_test(newest.p, old.p)) {
    origin_on_line(newest.p, old.p, v);
  } else {
    set3(v, newest.p);
    s->nvrtx = 1;
    put_slot(s, 0, &newest);
  }
}

/* 3 points: openGJK.c:270-313 (+ regions :129-154) */
static void sub_2d(osimplex* s, real* v) {
  svert a, b, c; /* a newest (slot 2), b slot 1, c slot 0 */
  get_slot(s, 2, &a);
  get_slot(s, 1, &b);
  get_slot(s, 0, &c);
  const int ab = edge_test(a.p, b.p);
  const int ac = edge_test(a.p, c.p);
  enum { FACE, EDGE_AB, EDGE_AC, VERT } leaf;
  if (ab) {
    if (!face_edge_test(a.p, b.p, c.p)) {
      if (ac)
        leaf = !face_edge_test(a.p, c.p, b.p) ? FACE : EDGE_AC;
      else
        leaf = FACE;
    } else {
      leaf = EDGE_AB;
    }
  } else if (ac) {
    leaf = !face_edge_test(a.p, c.p, b.p) ? FACE : EDGE_AC;
  } else {
    leaf = VERT;
  }
  switch (leaf) {
    case FACE:
      origin_on_plane(a.p, b.p, c.p, v);
      break;
    case EDGE_AB: /* S2Dregion12: slot0 := newest, slot1 keeps b */
      origin_on_line(a.p, b.p, v);
      s->nvrtx = 2;
      put_slot(s, 0, &a);
      break;
    case EDGE_AC: /* S2Dregion13: slot1 := newest, slot0 keeps c */
      origin_on_line(a.p, c.p, v);
      s->nvrtx = 2;
      put_slot(s, 1, &a);
      break;
    case VERT: /* S2Dregion1 */
      set3(v, a.p);
      s->nvrtx = 1;
      put_slot(s, 0, &a);
      break;
  }
}

/* leaves of the 4-point case (select_1xy / select_1x macros, openGJK.c:53-99) */
static inline void keep_face(osimplex* s, const svert* top, const svert* mid, const svert* low, real* v) {
  s->nvrtx = 3;
  put_slot(s, 2, top);
  put_slot(s, 1, mid);
  put_slot(s, 0, low);
  origin_on_plane(top->p, mid->p, low->p, v); /* bitwise symmetric in (mid, low) */
}
static inline void keep_edge(osimplex* s, const svert* top, const svert* low, real* v) {
  s->nvrtx = 2;
  put_slot(s, 1, top);
  put_slot(s, 0, low);
  origin_on_line(top->p, low->p, v);
}

/* 4 points: openGJK.c:315-613 */
static void sub_3d(osimplex* s, real* v) {
  svert w[4]; /* w[3] newest ("s1"), w[2] = s2, w[1] = s3, w[0] = s4 */
  for (int k = 0; k < 4; ++k) get_slot(s, k, &w[k]);
  const svert* s1 = &w[3];
  real e[3][3]; /* e[k] = w[k] - s1 */
  for (int k = 0; k < 3; ++k) sub3(w[k].p, s1->p, e[k]);

  int along[3]; /* along[k] = hff1(s1, slot k) */
  for (int k = 0; k < 3; ++k) along[k] = edge_test(s1->p, w[k].p);
  const int n_along = along[2] + along[1] + along[0];
  if (n_along == 0) { /* S3Dregion1 */
    set3(v, s1->p);
    s->nvrtx = 1;
    put_slot(s, 0, s1);
    return;
  }

  const int sss = det3(e[1], e[0], e[2]) <= 0;
  /* facing[k]: plane test of the face that omits slot k, 1 = origin on the inner side */
  int facing[3];
  facing[2] = plane_test(s1->p, w[1].p, w[0].p) - sss;
  facing[1] = plane_test(s1->p, w[0].p, w[2].p) - sss;
  facing[0] = plane_test(s1->p, w[2].p, w[1].p) - sss;
  for (int k = 0; k < 3; ++k) facing[k] *= facing[k];
  const int n_facing = facing[2] + facing[1] + facing[0];

  int k, i, j;
#define ROLES(kk) (k = (kk), i = ((kk) + 2) % 3, j = ((kk) + 1) % 3)
  if (n_facing == 3) { /* S3Dregion1234: origin enclosed */
    v[0] = v[1] = v[2] = 0;
    s->nvrtx = 4;
    return;
  }
  if (n_facing == 2) { /* drop the vertex opposite the only outward face, then 3-point case */
    s->nvrtx = 3;
    if (!facing[2]) {
      put_slot(s, 2, s1);
    } else if (!facing[1]) {
      put_slot(s, 1, &w[2]);
      put_slot(s, 2, s1);
    } else if (!facing[0]) {
      put_slot(s, 0, &w[1]);
      put_slot(s, 1, &w[2]);
      put_slot(s, 2, s1);
    }
    sub_2d(s, v);
    return;
  }
  if (n_facing == 1) {
    s->nvrtx = 3;
    if (facing[2]) ROLES(2);
    else if (facing[1]) ROLES(1);
    else ROLES(0);
    const svert *si = &w[i], *sj = &w[j], *sk = &w[k];
    if (n_along == 1) {
      if (along[k]) {
        if (!face_edge_test(s1->p, sk->p, si->p)) keep_face(s, s1, si, sk, v);
        else if (!face_edge_test(s1->p, sk->p, sj->p)) keep_face(s, s1, sj, sk, v);
        else keep_edge(s, s1, sk, v);
      } else if (along[i]) {
        if (!face_edge_test(s1->p, si->p, sk->p)) keep_face(s, s1, si, sk, v);
        else keep_edge(s, s1, si, v);
      } else {
        if (!face_edge_test(s1->p, sj->p, sk->p)) keep_face(s, s1, sj, sk, v);
        else keep_edge(s, s1, sj, v);
      }
    } else if (n_along == 2) {
      if (along[i]) {
        if (!face_edge_test(s1->p, sk->p, si->p)) {
          if (!face_edge_test(s1->p, si->p, sk->p)) keep_face(s, s1, si, sk, v);
          else keep_edge(s, s1, sk, v);
        } else {
          if (!face_edge_test(s1->p, sk->p, sj->p)) keep_face(s, s1, sj, sk, v);
          else keep_edge(s, s1, sk, v);
        }
      } else if (along[j]) {
        if (!face_edge_test(s1->p, sk->p, sj->p)) {
          if (!face_edge_test(s1->p, sj->p, sk->p)) keep_face(s, s1, sj, sk, v);
          else keep_edge(s, s1, sj, v);
        } else {
          if (!face_edge_test(s1->p, sk->p, si->p)) keep_face(s, s1, si, sk, v);
          else keep_edge(s, s1, sk, v);
        }
      }
      /* else: unreachable per the reference ("ERROR"); simplex keeps nvrtx = 3, v untouched */
    } else { /* n_along == 3 */
      const int ik = face_edge_test(s1->p, si->p, sk->p);
      const int jk = face_edge_test(s1->p, sj->p, sk->p);
      const int ki = face_edge_test(s1->p, sk->p, si->p);
      const int kj = face_edge_test(s1->p, sk->p, sj->p);
      if (ki && kj) keep_edge(s, s1, sk, v);
      else if (ki) {
        if (jk) keep_edge(s, s1, sj, v);
        else keep_face(s, s1, sj, sk, v);
      } else {
        if (ik) keep_edge(s, s1, si, v);
        else keep_face(s, s1, si, sk, v);
      }
    }
    return;
  }
  /* n_facing == 0: the origin is outside all three faces through s1 */
  if (n_along == 1) {
    if (along[1]) ROLES(2);
    else if (along[0]) ROLES(1);
    else ROLES(0);
    const svert *si = &w[i], *sj = &w[j], *sk = &w[k];
    if (!face_edge_test(s1->p, si->p, sj->p)) keep_face(s, s1, si, sj, v);
    else if (!face_edge_test(s1->p, si->p, sk->p)) keep_face(s, s1, si, sk, v);
    else keep_edge(s, s1, si, v);
  } else if (n_along == 2) {
    s->nvrtx = 3;
    if (!along[1]) ROLES(2);
    else if (!along[0]) ROLES(1);
    else ROLES(0);
    const svert *si = &w[i], *sj = &w[j], *sk = &w[k];
    if (!face_edge_test(s1->p, sj->p, sk->p)) {
      if (!face_edge_test(s1->p, sk->p, sj->p)) keep_face(s, s1, sj, sk, v);
      else if (!face_edge_test(s1->p, sk->p, si->p)) keep_face(s, s1, si, sk, v);
      else keep_edge(s, s1, sk, v);
    } else if (!face_edge_test(s1->p, sj->p, si->p)) {
      keep_face(s, s1, si, sj, v);
    } else {
      keep_edge(s, s1, sj, v);
    }
  }
  /* n_along == 3 with n_facing == 0: the reference does nothing (4 vertices kept, v kept) */
#undef ROLES
}

/* ------------------------------------------------------------------------------------------
 * A.2 support (openGJK.c:615-639): lowest index attaining a value strictly above the value
 * of the current support point; otherwise the current support point is kept.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  real s[3];
  int idx;
} osupport;

static void support_update(const obody* b, osupport* cur, const real* d) {
  real best = dot3(cur->s, d);
  int better = -1;
  for (int i = 0; i < b->n; ++i) {
    const real val = dot3(&b->xyz[3 * i], d);
    if (val > best) {
      best = val;
      better = i;
    }
  }
  if (better >= 0) {
    set3(cur->s, &b->xyz[3 * better]);
    cur->idx = better;
  }
}

/* ------------------------------------------------------------------------------------------
 * A.5 witnesses (openGJK.c:658-950), including the fall-through after a demotion
 * ---------------------------------------------------------------------------------------- */
static void wit_0d(const obody* b1, const obody* b2, osimplex* s) {
  for (int t = 0; t < 3; ++t) {
    s->witnesses[0][t] = b1->xyz[3 * s->vrtx_idx[0][0] + t];
    s->witnesses[1][t] = b2->xyz[3 * s->vrtx_idx[0][1] + t];
  }
}
static void blend(const obody* b1, const obody* b2, osimplex* s, int m, const real* a) {
  for (int t = 0; t < 3; ++t) {
    real w1 = b1->xyz[3 * s->vrtx_idx[0][0] + t] * a[0];
    real w2 = b2->xyz[3 * s->vrtx_idx[0][1] + t] * a[0];
    for (int q = 1; q < m; ++q) {
      w1 = w1 + b1->xyz[3 * s->vrtx_idx[q][0] + t] * a[q];
      w2 = w2 + b2->xyz[3 * s->vrtx_idx[q][1] + t] * a[q];
    }
    s->witnesses[0][t] = w1;
    s->witnesses[1][t] = w2;
  }
}
static void wit_1d(const obody* b1, const obody* b2, osimplex* s) {
  real pq[3], po[3];
  const real* p = s->vrtx[0];
  sub3(s->vrtx[1], p, pq);
  for (int t = 0; t < 3; ++t) po[t] = -p[t];
  const real det = dot3(pq, pq);
  if (det == 0.0) wit_0d(b1, b2, s); /* and falls through, openGJK.c:683-688 */
  real a[2];
  a[1] = dot3(pq, po) / det;
  a[0] = (real)(1.0 - (double)a[1]);
  blend(b1, b2, s, 2, a);
}
static void wit_2d(const obody* b1, const obody* b2, osimplex* s) {
  real pq[3], pr[3], po[3];
  const real* p = s->vrtx[0];
  sub3(s->vrtx[1], p, pq);
  sub3(s->vrtx[2], p, pr);
  for (int t = 0; t < 3; ++t) po[t] = -p[t];
  const real T00 = dot3(pq, pq), T01 = dot3(pq, pr), T11 = dot3(pr, pr);
  const real det = T00 * T11 - T01 * T01;
  if (det == 0.0) wit_1d(b1, b2, s);
  const real b0 = dot3(pq, po), b1v = dot3(pr, po);
  const real I00 = T11 / det, I01 = -T01 / det, I11 = T00 / det;
  real a[3];
  a[1] = I00 * b0 + I01 * b1v;
  a[2] = I01 * b0 + I11 * b1v;
  a[0] = (real)(1.0 - (double)a[1] - (double)a[2]);
  if (a[0] < R_EPS) { /* openGJK.c:761-780 */
    s->nvrtx = 2;
    svert t;
    get_slot(s, 2, &t);
    put_slot(s, 0, &t);
    wit_1d(b1, b2, s);
  } else if (a[1] < R_EPS) {
    s->nvrtx = 2;
    svert t;
    get_slot(s, 2, &t);
    put_slot(s, 1, &t);
    wit_1d(b1, b2, s);
  } else if (a[2] < R_EPS) {
    s->nvrtx = 2;
    wit_1d(b1, b2, s);
  }
  blend(b1, b2, s, 3, a); /* always executed, openGJK.c:782-794 */
}
static void wit_3d(const obody* b1, const obody* b2, osimplex* s) {
  real pq[3], pr[3], ps[3], po[3];
  const real* p = s->vrtx[0];
  sub3(s->vrtx[1], p, pq);
  sub3(s->vrtx[2], p, pr);
  sub3(s->vrtx[3], p, ps);
  for (int t = 0; t < 3; ++t) po[t] = -p[t];
  const real T00 = dot3(pq, pq), T01 = dot3(pq, pr), T02 = dot3(pq, ps);
  const real T11 = dot3(pr, pr), T12 = dot3(pr, ps), T22 = dot3(ps, ps);
  const real det00 = T11 * T22 - T12 * T12;
  const real det01 = T01 * T22 - T02 * T12;
  const real det02 = T01 * T12 - T02 * T11;
  const real det = T00 * det00 - T01 * det01 + T02 * det02;
  if (det == 0.0) wit_2d(b1, b2, s);
  const real b0 = dot3(pq, po), b1v = dot3(pr, po), b2v = dot3(ps, po);
  const real det11 = T00 * T22 - T02 * T02;
  const real det12 = T00 * T12 - T01 * T02;
  const real det22 = T00 * T11 - T01 * T01;
  const real I00 = det00 / det, I01 = -det01 / det, I02 = det02 / det;
  const real I11 = det11 / det, I12 = -det12 / det, I22 = det22 / det;
  real a[4];
  a[1] = I00 * b0 + I01 * b1v + I02 * b2v;
  a[2] = I01 * b0 + I11 * b1v + I12 * b2v;
  a[3] = I02 * b0 + I12 * b1v + I22 * b2v;
  a[0] = (real)(1.0 - (double)a[1] - (double)a[2] - (double)a[3]);
  int demote = -1; /* openGJK.c:883-910 */
  if (a[0] < R_EPS) demote = 0;
  else if (a[1] < R_EPS) demote = 1;
  else if (a[2] < R_EPS) demote = 2;
  else if (a[3] < R_EPS) demote = 3;
  if (demote >= 0) {
    s->nvrtx = 3;
    if (demote < 3) {
      svert t;
      get_slot(s, 3, &t);
      put_slot(s, demote, &t);
    }
    wit_2d(b1, b2, s);
  }
  blend(b1, b2, s, 4, a); /* always executed, openGJK.c:912-928 */
}
static void witnesses(const obody* b1, const obody* b2, osimplex* s) {
  switch (s->nvrtx) {
    case 4: wit_3d(b1, b2, s); break;
    case 3: wit_2d(b1, b2, s); break;
    case 2: wit_1d(b1, b2, s); break;
    case 1: wit_0d(b1, b2, s); break;
    default: break;
  }
}

/* ------------------------------------------------------------------------------------------
 * A.3 GJK loop (openGJK.c:952-1056).  Returns the distance; *iters = iterations executed.
 * ---------------------------------------------------------------------------------------- */
real ogjk_oracle_gjk(const real* xyz1, int n1, const real* xyz2, int n2, osimplex* s, int* iters) {
  const obody b1 = {xyz1, n1}, b2 = {xyz2, n2};
  const int max_iter = 25;
  const real eps_rel = (real)R_EPS * 1e4f;
  const real eps_tot = (real)R_EPS * 1e2f;
  const real eps_rel2 = eps_rel * eps_rel;
  real v[3], w[3], vneg[3];
  real norm2_wmax = 0;
  int k = 0;

  sub3(xyz1, xyz2, v);
  s->nvrtx = 1;
  set3(s->vrtx[0], v);
  s->vrtx_idx[0][0] = 0;
  s->vrtx_idx[0][1] = 0;
  osupport sp1, sp2;
  set3(sp1.s, xyz1);
  sp1.idx = 0;
  set3(sp2.s, xyz2);
  sp2.idx = 0;

  do {
    ++k;
    for (int t = 0; t < 3; ++t) vneg[t] = -v[t];
    support_update(&b1, &sp1, vneg);
    support_update(&b2, &sp2, v);
    sub3(sp1.s, sp2.s, w);

    const real gap = nrm2(v) - dot3(v, w);
    if (gap <= eps_rel * nrm2(v) || gap < eps_tot) break;
    if (nrm2(v) < eps_rel2) break;

    const int m = s->nvrtx;
    set3(s->vrtx[m], w);
    s->vrtx_idx[m][0] = sp1.idx;
    s->vrtx_idx[m][1] = sp2.idx;
    s->nvrtx = m + 1;

    switch (s->nvrtx) { /* openGJK.c:641-656 */
      case 4: sub_3d(s, v); break;
      case 3: sub_2d(s, v); break;
      case 2: sub_1d(s, v); break;
      default: break;
    }

    for (int q = 0; q < s->nvrtx; ++q) {
      const real t = nrm2(s->vrtx[q]);
      if (t > norm2_wmax) norm2_wmax = t;
    }
    if (nrm2(v) <= eps_tot * eps_tot * norm2_wmax) break;
  } while (s->nvrtx != 4 && k != max_iter);

  witnesses(&b1, &b2, s);
  if (iters) *iters = k;
  return R_SQRT(nrm2(v));
}

/* ==========================================================================================
 * A.6 EPA (reference GJK/cpu/EPA.c)
 * ======================================================================================== */
#define EPA_MAX_FACES 128 /* EPA.c:43 */
#define EPA_MAX_VERTS (EPA_MAX_FACES + 4)
#define EPA_MAX_ITERS 64 /* EPA.c:592 */

typedef struct {
  int v[3];      /* polytope vertex ids */
  int src[3][2]; /* provenance of each corner: [corner][body] */
  real n[3];
  real d;
  int live;
} oface;

typedef struct {
  real vert[EPA_MAX_VERTS][3];
  int vsrc[EPA_MAX_VERTS][2];
  int nvert;
  oface face[EPA_MAX_FACES];
  int face_hi; /* one past the highest slot ever used */
} opoly;

/* EPA.c:307-344: Minkowski support from scratch, lowest index on ties */
static void epa_support(const obody* b1, const obody* b2, const real* d, real* out, int* out_id) {
  real m1 = -1e10f, m2 = -1e10f;
  int i1 = -1, i2 = -1;
  for (int i = 0; i < b1->n; ++i) {
    const real t = b1->xyz[3 * i] * d[0] + b1->xyz[3 * i + 1] * d[1] + b1->xyz[3 * i + 2] * d[2];
    if (t > m1) {
      m1 = t;
      i1 = i;
    }
  }
  for (int i = 0; i < b2->n; ++i) {
    const real t = b2->xyz[3 * i] * d[0] + b2->xyz[3 * i + 1] * d[1] + b2->xyz[3 * i + 2] * d[2];
    if (-t > m2) {
      m2 = -t;
      i2 = i;
    }
  }
  if (i1 >= 0 && i2 >= 0) {
    for (int t = 0; t < 3; ++t) out[t] = b1->xyz[3 * i1 + t] - b2->xyz[3 * i2 + t];
    out_id[0] = i1;
    out_id[1] = i2;
  }
}

/* EPA.c:350-360 */
static void normal_from_witnesses(const real* w1, const real* w2, real* nrm) {
  real d[3];
  sub3(w2, w1, d);
  const real len = R_SQRT(nrm2(d));
  if (len > R_EPS) {
    nrm[0] = d[0] / len;
    nrm[1] = d[1] / len;
    nrm[2] = d[2] / len;
  } else {
    nrm[0] = 1.0f;
    nrm[1] = 0.0f;
    nrm[2] = 0.0f;
  }
}

/* "no progress" exits of the regrow stage (EPA.c:409-418, 474-483, 559-567) */
static void touch_exit(const obody* b1, const obody* b2, osimplex* s, const int* id, real* dist, real* nrm) {
  *dist = 0.0f;
  for (int c = 0; c < 3; ++c) {
    s->witnesses[0][c] = b1->xyz[3 * id[0] + c];
    s->witnesses[1][c] = b2->xyz[3 * id[1] + c];
  }
  normal_from_witnesses(s->witnesses[0], s->witnesses[1], nrm);
}

/* is `p` at squared distance >= eps^2 from every current simplex vertex? */
static int is_new_point(const osimplex* s, const real* p) {
  const real eps_sq = R_EPS * R_EPS;
  for (int q = 0; q < s->nvrtx; ++q) {
    const real dx = p[0] - s->vrtx[q][0], dy = p[1] - s->vrtx[q][1], dz = p[2] - s->vrtx[q][2];
    if (dx * dx + dy * dy + dz * dz < eps_sq) return 0;
  }
  return 1;
}
static void push_point(osimplex* s, const real* p, const int* id) {
  const int m = s->nvrtx;
  set3(s->vrtx[m], p);
  s->vrtx_idx[m][0] = id[0];
  s->vrtx_idx[m][1] = id[1];
  s->nvrtx = m + 1;
}

/* EPA.c:92-129 */
static void face_plane(opoly* P, int f) {
  oface* F = &P->face[f];
  const real *v0 = P->vert[F->v[0]], *v1 = P->vert[F->v[1]], *v2 = P->vert[F->v[2]];
  real e0[3], e1[3];
  sub3(v1, v0, e0);
  sub3(v2, v0, e1);
  cross3(e0, e1, F->n);
  const real len2 = nrm2(F->n);
  if (len2 > R_EPS * R_EPS) {
    const real len = R_SQRT(len2);
    for (int t = 0; t < 3; ++t) F->n[t] /= len;
    F->d = dot3(F->n, v0);
    if (F->d < 0) {
      for (int t = 0; t < 3; ++t) F->n[t] = -F->n[t];
      F->d = -F->d;
    }
  } else {
    F->live = 0;
    F->d = (real)1e10;
  }
}

/* winding fix shared by EPA.c:203-232 and :791-819 */
static void orient_outward(opoly* P, int f, const real* centroid) {
  oface* F = &P->face[f];
  const real *v0 = P->vert[F->v[0]], *v1 = P->vert[F->v[1]], *v2 = P->vert[F->v[2]];
  real e0[3], e1[3], n[3], tc[3];
  sub3(v1, v0, e0);
  sub3(v2, v0, e1);
  cross3(e0, e1, n);
  sub3(centroid, v0, tc);
  if (dot3(n, tc) > 0) {
    int t = F->v[1];
    F->v[1] = F->v[2];
    F->v[2] = t;
    for (int b = 0; b < 2; ++b) {
      t = F->src[1][b];
      F->src[1][b] = F->src[2][b];
      F->src[2][b] = t;
    }
  }
}

/* EPA.c:238-304 */
static void origin_barycentric(const real* v0, const real* v1, const real* v2, real* a) {
  real e0[3], e1[3];
  sub3(v1, v0, e0);
  sub3(v2, v0, e1);
  const real d00 = dot3(e0, e0), d01 = dot3(e0, e1), d11 = dot3(e1, e1);
  const real d20 = -dot3(v0, e0), d21 = -dot3(v0, e1);
  const real denom = d00 * d11 - d01 * d01;
  if (R_FABS(denom) < R_EPS) {
    a[0] = a[1] = a[2] = (real)1.0 / (real)3.0;
    return;
  }
  const real inv = (real)1.0 / denom;
  const real u = (d11 * d20 - d01 * d21) * inv;
  const real vv = (d00 * d21 - d01 * d20) * inv;
  const real w = (real)1.0 - u - vv;
  if (w < 0) {
    real e12[3];
    sub3(v2, v1, e12);
    real t = -dot3(v1, e12) / dot3(e12, e12);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    a[0] = 0;
    a[1] = (real)1.0 - t;
    a[2] = t;
  } else if (u < 0) {
    real t = -dot3(v0, e1) / dot3(e1, e1);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    a[0] = (real)1.0 - t;
    a[1] = 0;
    a[2] = t;
  } else if (vv < 0) {
    real t = -dot3(v0, e0) / dot3(e0, e0);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    a[0] = (real)1.0 - t;
    a[1] = t;
    a[2] = 0;
  } else {
    a[0] = w;
    a[1] = u;
    a[2] = vv;
  }
}

/* outputs of a terminated expansion (EPA.c:636-651, 667-683, 846-861) */
static void report_face(const obody* b1, const obody* b2, const opoly* P, int f, osimplex* s, real* dist,
                        real* nrm) {
  const oface* F = &P->face[f];
  real a[3];
  origin_barycentric(P->vert[F->v[0]], P->vert[F->v[1]], P->vert[F->v[2]], a);
  for (int t = 0; t < 3; ++t) {
    s->witnesses[0][t] = b1->xyz[3 * F->src[0][0] + t] * a[0] + b1->xyz[3 * F->src[1][0] + t] * a[1] +
                         b1->xyz[3 * F->src[2][0] + t] * a[2];
    s->witnesses[1][t] = b2->xyz[3 * F->src[0][1] + t] * a[0] + b2->xyz[3 * F->src[1][1] + t] * a[1] +
                         b2->xyz[3 * F->src[2][1] + t] * a[2];
    nrm[t] = F->n[t];
  }
  *dist = -F->d;
}

/* recompute planes of all live faces, then argmin distance with lowest slot on ties
 * (EPA.c:599-617 and :831-844) */
static int closest_face(opoly* P) {
  for (int f = 0; f < P->face_hi; ++f)
    if (P->face[f].live) face_plane(P, f);
  int best = -1;
  real best_d = 1e10f;
  for (int f = 0; f < P->face_hi; ++f) {
    if (!P->face[f].live) continue;
    if (P->face[f].d >= 0.0f && P->face[f].d < best_d) {
      best_d = P->face[f].d;
      best = f;
    }
  }
  return best;
}

/* EPA.c:362-863.  Returns the number of expansion iterations (0 if the gate/regrow returned). */
int ogjk_oracle_epa(const real* xyz1, int n1, const real* xyz2, int n2, osimplex* s, real* dist, real* nrm) {
  const obody b1 = {xyz1, n1}, b2 = {xyz2, n2};

  /* 1. gate (EPA.c:369-373) */
  if (*dist > R_EPS) {
    normal_from_witnesses(s->witnesses[0], s->witnesses[1], nrm);
    return 0;
  }

  /* 2. regrow the simplex to a tetrahedron (EPA.c:375-583) */
  if (s->nvrtx != 4) {
    real p[3];
    int id[2];
    if (s->nvrtx == 1) {
      epa_support(&b1, &b2, s->vrtx[0], p, id);
      if (is_new_point(s, p)) push_point(s, p, id);
      else {
        touch_exit(&b1, &b2, s, id, dist, nrm);
        return 0;
      }
    }
    if (s->nvrtx == 2) {
      real edge[3], dir[3];
      sub3(s->vrtx[1], s->vrtx[0], edge);
      real axis[3] = {1.0f, 0.0f, 0.0f};
      const real len = R_SQRT(nrm2(edge));
      if (len > R_EPS && R_FABS(edge[0]) > 0.9f * len) {
        axis[0] = 0.0f;
        axis[1] = 1.0f;
      }
      cross3(edge, axis, dir);
      if (nrm2(dir) < R_EPS) {
        axis[0] = 0.0f;
        axis[1] = 0.0f;
        axis[2] = 1.0f;
        cross3(edge, axis, dir);
      }
      epa_support(&b1, &b2, dir, p, id);
      if (is_new_point(s, p)) push_point(s, p, id);
      else {
        touch_exit(&b1, &b2, s, id, dist, nrm);
        return 0;
      }
    }
    if (s->nvrtx == 3) {
      real e0[3], e1[3], dir[3];
      sub3(s->vrtx[1], s->vrtx[0], e0);
      sub3(s->vrtx[2], s->vrtx[0], e1);
      cross3(e0, e1, dir);
      epa_support(&b1, &b2, dir, p, id);
      if (is_new_point(s, p)) {
        push_point(s, p, id);
      } else {
        for (int t = 0; t < 3; ++t) dir[t] = -dir[t];
        epa_support(&b1, &b2, dir, p, id);
        if (is_new_point(s, p)) push_point(s, p, id);
        else {
          touch_exit(&b1, &b2, s, id, dist, nrm);
          return 0;
        }
      }
    }
    if (s->nvrtx != 4) { /* EPA.c:571-582 (nvrtx outside 1..4) */
      const int best = s->nvrtx > 0 ? s->nvrtx - 1 : 0;
      touch_exit(&b1, &b2, s, s->vrtx_idx[best], dist, nrm);
      return 0;
    }
  }

  /* 3. tetrahedron (EPA.c:144-235) */
  opoly P;
  memset(P.face, 0, sizeof(P.face));
  P.nvert = 4;
  for (int q = 0; q < 4; ++q) {
    set3(P.vert[q], s->vrtx[q]);
    P.vsrc[q][0] = s->vrtx_idx[q][0];
    P.vsrc[q][1] = s->vrtx_idx[q][1];
  }
  real centroid[3] = {0.0f, 0.0f, 0.0f};
  for (int q = 0; q < 4; ++q)
    for (int t = 0; t < 3; ++t) centroid[t] += P.vert[q][t] * 0.25f;
  static const int tetra[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
  for (int f = 0; f < 4; ++f) {
    for (int c = 0; c < 3; ++c) {
      P.face[f].v[c] = tetra[f][c];
      P.face[f].src[c][0] = P.vsrc[tetra[f][c]][0];
      P.face[f].src[c][1] = P.vsrc[tetra[f][c]][1];
    }
    P.face[f].live = 1;
    orient_outward(&P, f, centroid);
  }
  P.face_hi = 4;

  /* 4. expansion (EPA.c:596-826) */
  const real tol = ((real)(R_EPS) * (real)1e2);
  int iter = 0;
  while (iter < EPA_MAX_ITERS && P.nvert < EPA_MAX_VERTS - 1) {
    ++iter;
    const int cf = closest_face(&P);
    if (cf < 0) break;
    const real cd = P.face[cf].d;

    real w[3];
    int wid[2];
    epa_support(&b1, &b2, P.face[cf].n, w, wid);
    const real gain = dot3(P.face[cf].n, w) - cd;
    if (gain < tol) {
      report_face(&b1, &b2, &P, cf, s, dist, nrm);
      break;
    }
    int dup = 0;
    {
      const real eps_sq = R_EPS * R_EPS;
      for (int q = 0; q < P.nvert; ++q) {
        const real dx = w[0] - P.vert[q][0], dy = w[1] - P.vert[q][1], dz = w[2] - P.vert[q][2];
        if (dx * dx + dy * dy + dz * dz < eps_sq) {
          dup = 1;
          break;
        }
      }
    }
    if (dup) {
      report_face(&b1, &b2, &P, cf, s, dist, nrm);
      break;
    }

    const int nv = P.nvert;
    set3(P.vert[nv], w);
    P.vsrc[nv][0] = wid[0];
    P.vsrc[nv][1] = wid[1];
    P.nvert = nv + 1;
    const real inv_n = (real)1.0 / (real)P.nvert;
    for (int t = 0; t < 3; ++t) centroid[t] += (w[t] - centroid[t]) * inv_n;

    /* faces that see w die; their edges are candidates for the horizon (EPA.c:700-743) */
    struct {
      int a, b, sa[2], sb[2], keep;
    } edge[EPA_MAX_FACES * 3];
    int nedge = 0;
    for (int f = 0; f < P.face_hi; ++f) {
      oface* F = &P.face[f];
      if (!F->live) continue;
      real diff[3];
      sub3(w, P.vert[F->v[0]], diff);
      if (!(dot3(F->n, diff) > R_EPS)) continue;
      for (int c = 0; c < 3 && nedge < EPA_MAX_FACES * 3; ++c) {
        const int c2 = (c + 1) % 3;
        edge[nedge].a = F->v[c];
        edge[nedge].b = F->v[c2];
        edge[nedge].sa[0] = F->src[c][0];
        edge[nedge].sa[1] = F->src[c][1];
        edge[nedge].sb[0] = F->src[c2][0];
        edge[nedge].sb[1] = F->src[c2][1];
        edge[nedge].keep = 1;
        ++nedge;
      }
      F->live = 0;
    }
    /* an edge seen from two dead faces is interior (EPA.c:745-759) */
    for (int x = 0; x < nedge; ++x) {
      if (!edge[x].keep) continue;
      for (int y = x + 1; y < nedge; ++y) {
        if (!edge[y].keep) continue;
        if ((edge[x].a == edge[y].a && edge[x].b == edge[y].b) ||
            (edge[x].a == edge[y].b && edge[x].b == edge[y].a)) {
          edge[x].keep = 0;
          edge[y].keep = 0;
        }
      }
    }
    /* stitch the horizon to the new vertex, lowest free slot first (EPA.c:761-825) */
    for (int x = 0; x < nedge; ++x) {
      if (!edge[x].keep) continue;
      int slot = -1;
      for (int f = 0; f < EPA_MAX_FACES; ++f)
        if (!P.face[f].live) {
          slot = f;
          break;
        }
      if (slot < 0) break;
      oface* F = &P.face[slot];
      F->v[0] = edge[x].a;
      F->v[1] = edge[x].b;
      F->v[2] = nv;
      for (int b = 0; b < 2; ++b) {
        F->src[0][b] = edge[x].sa[b];
        F->src[1][b] = edge[x].sb[b];
        F->src[2][b] = wid[b];
      }
      F->live = 1;
      orient_outward(&P, slot, centroid);
      if (slot >= P.face_hi) P.face_hi = slot + 1;
    }
  }

  /* 5. iteration cap: report the currently closest face (EPA.c:828-863) */
  if (iter >= EPA_MAX_ITERS) {
    const int cf = closest_face(&P);
    if (cf >= 0) report_face(&b1, &b2, &P, cf, s, dist, nrm);
  }
  return iter;
}

/* ==========================================================================================
 * batch drivers (same flat format as oracle/ref_driver.c)
 * ======================================================================================== */
int ogjk_oracle_sizeof_real(void) { return (int)sizeof(real); }
int ogjk_oracle_sizeof_simplex(void) { return (int)sizeof(osimplex); }

static inline const real* body_ptr(const real* c, const long* off, int nv, long i, int* n) {
  const long first = off ? off[i] : i * (long)nv;
  *n = off ? (int)(off[i + 1] - off[i]) : nv;
  return c + 3 * first;
}

void ogjk_oracle_gjk_batch(long n, const real* c1, const long* off1, int nv1, const real* c2,
                           const long* off2, int nv2, osimplex* simplices, real* distances, int* iters,
                           int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < n; ++i) {
    int n1, n2;
    const real* a = body_ptr(c1, off1, nv1, i, &n1);
    const real* b = body_ptr(c2, off2, nv2, i, &n2);
    distances[i] = ogjk_oracle_gjk(a, n1, b, n2, &simplices[i], iters ? &iters[i] : NULL);
  }
}

void ogjk_oracle_epa_batch(long n, const real* c1, const long* off1, int nv1, const real* c2,
                           const long* off2, int nv2, osimplex* simplices, real* distances, real* normals,
                           int* iters, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < n; ++i) {
    int n1, n2;
    const real* a = body_ptr(c1, off1, nv1, i, &n1);
    const real* b = body_ptr(c2, off2, nv2, i, &n2);
    const int it = ogjk_oracle_epa(a, n1, b, n2, &simplices[i], &distances[i], &normals[3 * i]);
    if (iters) iters[i] = it;
  }
}

void ogjk_oracle_gjk_epa_indexed(long npairs, const real* pool, const long* off, int nv, const int* pairs,
                                 osimplex* simplices, real* distances, real* normals, int do_gjk,
                                 int do_epa, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < npairs; ++i) {
    int n1, n2;
    const real* a = body_ptr(pool, off, nv, pairs[2 * i], &n1);
    const real* b = body_ptr(pool, off, nv, pairs[2 * i + 1], &n2);
    if (do_gjk) distances[i] = ogjk_oracle_gjk(a, n1, b, n2, &simplices[i], NULL);
    if (do_epa) ogjk_oracle_epa(a, n1, b, n2, &simplices[i], &distances[i], &normals[3 * i]);
  }
}
