/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin batch driver around the UNMODIFIED reference CPU sources
 *   /root/reference/GJK/cpu/openGJK.c  (compute_minimum_distance, :952)
 *   /root/reference/GJK/cpu/EPA.c      (computeCollisionInformation, :362)
 * which are compiled where they lie by oracle/build_ref.sh into oracle/_ref/.
 * The loop below is the same single `for` over pairs as the reference's
 * GJK::CPU::computeDistances / computeEPA (examples/cpu/example.cpp:13-40),
 * optionally spread over OpenMP threads (pairs are independent).
 *
 * Flat input format (shared with oracle/oracle_driver.c and the CUDA C-ABI):
 *   c1/c2 : concatenated xyz vertex coordinates of body 1 / body 2 of every pair
 *   off   : n+1 vertex offsets (NULL => uniform nv vertices per polytope)
 */
#include <stddef.h>
#include "GJK/cpu/openGJK.h"
#include "GJK/cpu/EPA.h"

static void make_body(gkPolytope* b, const gkFloat* c, const long* off, int nv, long i) {
  long first = off ? off[i] : (long)i * nv;
  b->numpoints = off ? (int)(off[i + 1] - off[i]) : nv;
  b->coord = (gkFloat*)(c + 3 * first);
  b->s[0] = b->s[1] = b->s[2] = 0;
  b->s_idx = 0;
}

int ogjk_ref_sizeof_real(void) { return (int)sizeof(gkFloat); }
int ogjk_ref_sizeof_simplex(void) { return (int)sizeof(gkSimplex); }
int ogjk_ref_sizeof_polytope(void) { return (int)sizeof(gkPolytope); }

void ogjk_ref_gjk_batch(long n, const gkFloat* c1, const long* off1, int nv1,
                        const gkFloat* c2, const long* off2, int nv2,
                        gkSimplex* simplices, gkFloat* distances, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < n; ++i) {
    gkPolytope a, b;
    make_body(&a, c1, off1, nv1, i);
    make_body(&b, c2, off2, nv2, i);
    distances[i] = compute_minimum_distance(a, b, &simplices[i]);
  }
}

void ogjk_ref_epa_batch(long n, const gkFloat* c1, const long* off1, int nv1,
                        const gkFloat* c2, const long* off2, int nv2,
                        gkSimplex* simplices, gkFloat* distances,
                        gkFloat* normals, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < n; ++i) {
    gkPolytope a, b;
    make_body(&a, c1, off1, nv1, i);
    make_body(&b, c2, off2, nv2, i);
    computeCollisionInformation(&a, &b, &simplices[i], &distances[i], &normals[3 * i]);
  }
}

/* indexed variant: polytopes come from one pool, pairs = (idx1, idx2) */
void ogjk_ref_gjk_epa_indexed(long npairs, const gkFloat* pool, const long* off, int nv,
                              const int* pairs, gkSimplex* simplices, gkFloat* distances,
                              gkFloat* normals, int do_gjk, int do_epa, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) if (nthreads > 1)
  for (long i = 0; i < npairs; ++i) {
    gkPolytope a, b;
    make_body(&a, pool, off, nv, pairs[2 * i]);
    make_body(&b, pool, off, nv, pairs[2 * i + 1]);
    if (do_gjk) distances[i] = compute_minimum_distance(a, b, &simplices[i]);
    if (do_epa) computeCollisionInformation(&a, &b, &simplices[i], &distances[i], &normals[3 * i]);
  }
}
