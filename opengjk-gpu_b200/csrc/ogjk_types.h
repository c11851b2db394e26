// ogjk_types.h -- precision-templated mirrors of the reference's POD records.
// Layouts are byte-identical to reference GJK/common.h:68-89 and GJK/gpu/openGJK.h:309-312
// (fp32 32/108 B, fp64 48/184 B).
#pragma once
#include <stddef.h>

namespace ogjk {

template <typename T>
struct PolytopeT {
  int numpoints;
  T s[3];
  int s_idx;
  T* coord;
};

template <typename T>
struct SimplexT {
  int nvrtx;
  T vrtx[4][3];
  int vrtx_idx[4][2];
  T witnesses[2][3];
};

struct CollisionPair {
  int idx1, idx2;
};

static_assert(sizeof(PolytopeT<float>) == 32 && offsetof(PolytopeT<float>, coord) == 24, "gkPolytope fp32 layout");
static_assert(sizeof(PolytopeT<double>) == 48 && offsetof(PolytopeT<double>, coord) == 40, "gkPolytope fp64 layout");
static_assert(sizeof(SimplexT<float>) == 108 && offsetof(SimplexT<float>, witnesses) == 84, "gkSimplex fp32 layout");
static_assert(sizeof(SimplexT<double>) == 184 && offsetof(SimplexT<double>, witnesses) == 136, "gkSimplex fp64 layout");
static_assert(sizeof(CollisionPair) == 8, "gkCollisionPair layout");

}  // namespace ogjk
