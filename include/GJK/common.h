/*
 * GJK/common.h -- drop-in replacement for the reference header of the same path (reference GJK/common.h:35-91).
 *
 * Same include guard, same macro family, same POD layouts, so a translation unit written against the reference
 * compiles unchanged against include/ of this repository.  Precision switch as in the reference: USE_32BITS
 * selects float (the default, reference GJK/common.h:44); define OGJK_USE_64BITS (or remove USE_32BITS below)
 * for double.  Both precisions are served by the same shared library (symbols ogjk_f32_* / ogjk_f64_*).
 */
#ifndef COMMON_H__
#define COMMON_H__

#include <float.h>

#if !defined(OGJK_USE_64BITS) && !defined(USE_32BITS)
#define USE_32BITS
#endif
#if defined(OGJK_USE_64BITS) && defined(USE_32BITS)
#undef USE_32BITS
#endif

#ifdef USE_32BITS
#define gkFloat float
#define gkEpsilon FLT_EPSILON
#define gkSqrt sqrtf
#define gkFmax fmaxf
#define gkFmin fminf
#define gkFabs fabsf
#define OGJK_API(name) ogjk_f32_##name
#else
#define gkFloat double
#define gkEpsilon DBL_EPSILON
#define gkSqrt sqrt
#define gkFmax fmax
#define gkFmin fmin
#define gkFabs fabs
#define OGJK_API(name) ogjk_f64_##name
#endif

/* A convex polytope = a cloud of `numpoints` vertices, flattened x0 y0 z0 x1 y1 z1 ...; `coord` is owned by the
 * caller.  `s` / `s_idx` exist for layout compatibility; the GPU path ignores them on input. */
typedef struct gkPolytope {
  int numpoints;
  gkFloat s[3];
  int s_idx;
  gkFloat* coord;
} gkPolytope;

/* Result simplex of a GJK query: `nvrtx` Minkowski-difference points, their source vertex indices on body 1 / 2,
 * and the two witness points. */
typedef struct gkSimplex {
  int nvrtx;
  gkFloat vrtx[4][3];
  int vrtx_idx[4][2];
  gkFloat witnesses[2][3];
} gkSimplex;

#ifdef __cplusplus
static_assert(sizeof(gkPolytope) == (sizeof(gkFloat) == 4 ? 32 : 48), "gkPolytope layout (reference GJK/common.h:68-78)");
static_assert(sizeof(gkSimplex) == (sizeof(gkFloat) == 4 ? 108 : 184), "gkSimplex layout (reference GJK/common.h:84-89)");
#endif

#endif /* COMMON_H__ */
