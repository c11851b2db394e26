// tests/cxx_dropin.cpp -- a caller written against the REFERENCE's C++ API (same include paths, same calls as
// examples/usage/GJKUsage.c:140-160 and EPAUsage.cpp:85-125), compiled with plain g++ against include/ and linked
// to libopengjk_b200.so.  Prints the README lines; tests/test_cxx_dropin.py checks them.
#include <cmath>
#include <cstdio>
#include <vector>

#include "examples/gpu/example.h"
#include "GJK/common.h"

static const gkFloat P[9][3] = {{0.0, 5.5, 0.0}, {2.3, 1.0, -2.0}, {8.1, 4.0, 2.4}, {4.3, 5.0, 2.2}, {2.5, 1.0, 2.3},
                                {7.1, 1.0, 2.4}, {1.0, 1.5, 0.3}, {3.3, 0.5, 0.3}, {6.0, 1.4, 0.2}};

int main() {
  // ---- GJK on userP / userQ (reference README.md:111-115) ----
  std::vector<gkFloat> p(27), q(27);
  for (int i = 0; i < 9; ++i)
    for (int c = 0; c < 3; ++c) {
      p[3 * i + c] = P[i][c];
      q[3 * i + c] = -P[i][c];
    }
  q[2] = 0.0f;  // userQ.dat is the mirror image of userP except that its first z is +0.0, not -0.0
  gkPolytope bd1, bd2;
  bd1.coord = p.data();
  bd1.numpoints = 9;
  bd2.coord = q.data();
  bd2.numpoints = 9;
  gkSimplex s;
  s.nvrtx = 0;
  gkFloat dist[1];
  gkSimplex simplices[1] = {s};
  GJK::GPU::computeDistances(1, &bd1, &bd2, simplices, dist);
  if (ogjk_last_error()[0]) {
    std::printf("ERROR %s\n", ogjk_last_error());
    return 2;
  }
  std::printf("Distance between bodies %f\n", (double)dist[0]);
  std::printf("Witnesses: (%f, %f, %f) and (%f, %f, %f)\n", (double)simplices[0].witnesses[0][0],
              (double)simplices[0].witnesses[0][1], (double)simplices[0].witnesses[0][2],
              (double)simplices[0].witnesses[1][0], (double)simplices[0].witnesses[1][1],
              (double)simplices[0].witnesses[1][2]);

  // ---- EPA: unit cube vs cube rotated 45 deg about x, y, z and shifted +1 in x (README.md:131-137) ----
  const gkFloat pi = 3.14159265358979323846;
  const gkFloat angle = 45.0f * pi / 180.0f;
  const gkFloat ca = std::cos(angle), sa = std::sin(angle);
  std::vector<gkFloat> c1(24), c2(24);
  int idx = 0;
  for (int x = -1; x <= 1; x += 2)
    for (int y = -1; y <= 1; y += 2)
      for (int z = -1; z <= 1; z += 2) {
        c1[3 * idx] = (gkFloat)x;
        c1[3 * idx + 1] = (gkFloat)y;
        c1[3 * idx + 2] = (gkFloat)z;
        gkFloat px = (gkFloat)x, py = (gkFloat)y, pz = (gkFloat)z;
        gkFloat ty = py * ca - pz * sa, tz = py * sa + pz * ca;
        py = ty;
        pz = tz;
        gkFloat tx = px * ca + pz * sa;
        tz = -px * sa + pz * ca;
        px = tx;
        pz = tz;
        tx = px * ca - py * sa;
        ty = px * sa + py * ca;
        px = tx;
        py = ty;
        c2[3 * idx] = px + 1.0f;
        c2[3 * idx + 1] = py;
        c2[3 * idx + 2] = pz;
        ++idx;
      }
  gkPolytope a, b;
  a.numpoints = 8;
  a.coord = c1.data();
  b.numpoints = 8;
  b.coord = c2.data();
  gkFloat normal[3] = {0, 0, 0};
  simplices[0].nvrtx = 0;
  GJK::GPU::computeDistances(1, &a, &b, simplices, dist);
  GJK::GPU::computeEPA(1, &a, &b, simplices, dist, normal);
  std::printf("Penetration depth: %.6f\n", -(double)dist[0]);
  std::printf("Witness point on cube 1: (%.6f, %.6f, %.6f)\n", (double)simplices[0].witnesses[0][0],
              (double)simplices[0].witnesses[0][1], (double)simplices[0].witnesses[0][2]);
  std::printf("Witness point on cube 2: (%.6f, %.6f, %.6f)\n", (double)simplices[0].witnesses[1][0],
              (double)simplices[0].witnesses[1][1], (double)simplices[0].witnesses[1][2]);
  std::printf("Contact normal (from cube 1 to cube 2): (%.6f, %.6f, %.6f)\n", (double)normal[0], (double)normal[1],
              (double)normal[2]);

  // ---- README spelling with witness arrays, fused path, indexed path ----
  gkFloat w1[3], w2[3], n2[3];
  GJK::GPU::computeCollisionInformation(1, &a, &b, simplices, dist, w1, w2, n2);
  std::printf("README-API depth %.6f w1 (%.6f, %.6f, %.6f)\n", -(double)dist[0], (double)w1[0], (double)w1[1],
              (double)w1[2]);
  gkPolytope pool[2] = {a, b};
  gkCollisionPair pr = {0, 1};
  compute_gjk_epa_indexed(2, 1, pool, &pr, simplices, dist, n2);
  std::printf("indexed depth %.6f\n", -(double)dist[0]);
  GJK::GPU::computeGJKAndEPA(0, &a, &b, simplices, dist, n2);  // n <= 0: silent no-op
  return ogjk_last_error()[0] ? 2 : 0;
}
