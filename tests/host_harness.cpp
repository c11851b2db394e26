// tests/host_harness.cpp -- TEST INFRASTRUCTURE: runs the product's per-thread GJK core
// (opengjk-gpu_b200/csrc/gjk_core.cuh + gjk_tables.h) on the CPU so that its logic can be compared with
// the oracle bit-for-bit in the `-m "not gpu"` suite.  The warp-level support search is emulated by the
// same "global max, lowest index, update only if strictly better than the current support" rule the
// kernels implement with shuffles.  Built by tests/conftest.py with g++ -ffp-contract=off.
#include <cstring>

#include "gjk_core.cuh"
#include "gjk_tables.h"

using namespace ogjk;

template <typename T>
struct SimplexOut {
  int nvrtx;
  T vrtx[4][3];
  int vrtx_idx[4][2];
  T witnesses[2][3];
};

template <typename T>
struct HostFetch {
  const T* c[2];
  V3<T> operator()(int body, int i) const { return mk<T>(c[body][3 * i], c[body][3 * i + 1], c[body][3 * i + 2]); }
};

template <typename T>
static void support(const T* c, int n, const V3<T>& d, V3<T>& sup, int& idx) {
  T best = dot(c[0], c[1], c[2], d);
  int bi = 0;
  for (int i = 1; i < n; ++i) {
    const T val = dot(c[3 * i], c[3 * i + 1], c[3 * i + 2], d);
    if (val > best) {
      best = val;
      bi = i;
    }
  }
  if (best > dot(sup, d)) {
    sup = mk<T>(c[3 * bi], c[3 * bi + 1], c[3 * bi + 2]);
    idx = bi;
  }
}

template <typename T>
static void put(SimplexOut<T>& o, int j, const SV<T>& s, bool live) {
  o.vrtx[j][0] = live ? s.p.x : T(0);
  o.vrtx[j][1] = live ? s.p.y : T(0);
  o.vrtx[j][2] = live ? s.p.z : T(0);
  o.vrtx_idx[j][0] = live ? s.i1 : 0;
  o.vrtx_idx[j][1] = live ? s.i2 : 0;
}

template <typename T>
static void run(long n, const T* c1, const long* off1, int nv1, const T* c2, const long* off2, int nv2,
                SimplexOut<T>* out, T* dist, int* iters, bool unified = false) {
  static LeafTables tabs;
  static uint16_t utab[kUnifiedSize];
  static bool built = false;
  if (!built) {
    build_leaf_tables(tabs);
    build_unified_table(utab);
    built = true;
  }
  for (long i = 0; i < n; ++i) {
    const T* a = c1 + 3 * (off1 ? off1[i] : i * (long)nv1);
    const T* b = c2 + 3 * (off2 ? off2[i] : i * (long)nv2);
    const int na = off1 ? (int)(off1[i + 1] - off1[i]) : nv1;
    const int nb = off2 ? (int)(off2[i + 1] - off2[i]) : nv2;
    GjkState<T> g;
    gjk_init(g, mk<T>(a[0], a[1], a[2]), mk<T>(b[0], b[1], b[2]));
    bool stop;
    do {
      ++g.k;
      support(a, na, vneg(g.v), g.sup1, g.idx1);
      support(b, nb, g.v, g.sup2, g.idx2);
      stop = unified ? gjk_advance_u(g, utab) : gjk_advance(g, tabs.t2, tabs.t3);
    } while (!stop);
    HostFetch<T> f{{a, b}};
    V3<T> w1, w2;
    gjk_witnesses(f, g.S, w1, w2);
    SimplexOut<T>& o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.nvrtx = g.S.n;
    put(o, 0, g.S.s0, g.S.n > 0);
    put(o, 1, g.S.s1, g.S.n > 1);
    put(o, 2, g.S.s2, g.S.n > 2);
    put(o, 3, g.S.s3, g.S.n > 3);
    o.witnesses[0][0] = w1.x; o.witnesses[0][1] = w1.y; o.witnesses[0][2] = w1.z;
    o.witnesses[1][0] = w2.x; o.witnesses[1][1] = w2.y; o.witnesses[1][2] = w2.z;
    dist[i] = sqrt_rn(norm2(g.v));
    if (iters) iters[i] = g.k;
  }
}

extern "C" {
void harness_gjk_f32(long n, const float* c1, const long* off1, int nv1, const float* c2, const long* off2, int nv2,
                     void* simplices, float* dist, int* iters) {
  run<float>(n, c1, off1, nv1, c2, off2, nv2, (SimplexOut<float>*)simplices, dist, iters);
}
// lane-uniform iteration of the persistent slot kernel (gjk_advance_u + the unified 16-bit table)
void harness_gjku_f32(long n, const float* c1, const long* off1, int nv1, const float* c2, const long* off2, int nv2,
                      void* simplices, float* dist, int* iters) {
  run<float>(n, c1, off1, nv1, c2, off2, nv2, (SimplexOut<float>*)simplices, dist, iters, true);
}
void harness_gjku_f64(long n, const double* c1, const long* off1, int nv1, const double* c2, const long* off2, int nv2,
                      void* simplices, double* dist, int* iters) {
  run<double>(n, c1, off1, nv1, c2, off2, nv2, (SimplexOut<double>*)simplices, dist, iters, true);
}
void harness_gjk_f64(long n, const double* c1, const long* off1, int nv1, const double* c2, const long* off2, int nv2,
                     void* simplices, double* dist, int* iters) {
  run<double>(n, c1, off1, nv1, c2, off2, nv2, (SimplexOut<double>*)simplices, dist, iters);
}
}
