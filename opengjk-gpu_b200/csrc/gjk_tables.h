// gjk_tables.h -- the signed-volumes decision trees of the GJK sub-algorithm, as DATA.
//
// The reference walks a nested if/else tree per simplex (S2D / S3D, reference
// GJK/gpu/openGJK.cu:302-827 == GJK/cpu/openGJK.c:270-613).  Every decision in those trees is a
// function of a handful of sign predicates of the simplex, so the kernels here evaluate ALL
// predicates branch-free, pack them into an index and look the outcome ("leaf") up in a table
// built once on the host by the functions below.  Warp lanes that took different branches in the
// reference execute identical instructions here.
//
// Slot convention (reference openGJK.c:324-327): the newest simplex point `a` sits in the
// highest slot (slot 3 for a tetrahedron, slot 2 for a triangle); older points below it.
//
// Leaf encoding (uint32):
//   bits  0..2   number of vertices kept (1..4)
//   bits  4..11  source slot of output slot j at bits 4+2j..5+2j   (a full permutation)
//   bits 12..14  how the new search vector v is obtained:
//                  0 = v := a            1 = closest point on line (a, X)
//                  2 = closest point on plane (a, X, Y)   3 = v := 0 (origin enclosed)
//                  4 = v unchanged
//   bits 16..17  X (kind 1: slot of the other edge end; kind 2: face id 0:{0,1} 1:{0,2} 2:{1,2})
//   bit  20      the newest point `a` survives (drives the running max |w|^2, openGJK.c:1035-1040)
//
// Index of the 4-point table (12 bits):
//   bits 0..2  along[k]  = hff1(a, slot k)                       (openGJK.c:333-339)
//   bits 3..5  facing[k] = (hff3(face omitting slot k) - sss)^2  (openGJK.c:345-353)
//   bits 6..11 hff2(a, x, y) for (x,y) = (0,1) (1,0) (0,2) (2,0) (1,2) (2,1)
// Index of the 3-point table (4 bits): along[1], along[0], hff2(a,1,0), hff2(a,0,1).
#pragma once
#include <stdint.h>

namespace ogjk {

enum : int { VK_A = 0, VK_LINE = 1, VK_PLANE = 2, VK_ZERO = 3, VK_KEEP = 4 };

struct LeafSpec {
  int nv;
  int src[4];
  int vkind;
  int x;
  int keeps_a;
};

inline uint32_t encode_leaf(const LeafSpec& l) {
  uint32_t e = (uint32_t)l.nv;
  for (int j = 0; j < 4; ++j) e |= (uint32_t)(l.src[j] & 3) << (4 + 2 * j);
  e |= (uint32_t)l.vkind << 12;
  e |= (uint32_t)(l.x & 3) << 16;
  e |= (uint32_t)(l.keeps_a & 1) << 20;
  return e;
}

// complete src[] to a permutation of {0,1,2,3}: dropped slots go to the unused output slots
inline void fill_perm(LeafSpec& l) {
  bool used[4] = {false, false, false, false};
  for (int j = 0; j < l.nv; ++j) used[l.src[j]] = true;
  int next = 0;
  for (int j = l.nv; j < 4; ++j) {
    while (used[next]) ++next;
    l.src[j] = next;
    used[next] = true;
  }
}

inline int face_id(int x, int y) {  // unordered pair of slots {0,1,2} -> 0,1,2
  const int lo = x < y ? x : y, hi = x < y ? y : x;
  return lo == 0 ? (hi == 1 ? 0 : 1) : 2;
}

namespace detail {
// leaves expressed with `top` = slot of the newest point
inline LeafSpec leaf_vertex(int top) {
  LeafSpec l = {1, {top, 0, 0, 0}, VK_A, 0, 1};
  fill_perm(l);
  return l;
}
// select_1x macros (openGJK.c:80-99): slot0 := x, slot1 := a
inline LeafSpec leaf_edge(int top, int low) {
  LeafSpec l = {2, {low, top, 0, 0}, VK_LINE, low, 1};
  fill_perm(l);
  return l;
}
// select_1xy macros (openGJK.c:53-78): slot0 := low, slot1 := mid, slot2 := a
inline LeafSpec leaf_face(int top, int mid, int low) {
  LeafSpec l = {3, {low, mid, top, 0}, VK_PLANE, face_id(mid, low), 1};
  fill_perm(l);
  return l;
}

// 3-point tree (openGJK.c:270-313) for a triangle whose newest point is `a` (slot `top`),
// second point `b` (in slot sb) and third point `c` (in slot sc); hbc = hff2(a,b,c), hcb = hff2(a,c,b).
// The S2Dregion macros (openGJK.c:129-154) overwrite slots in place: region12 puts `a` in slot 0
// (b stays in slot 1), region13 puts `a` in slot 1 (c stays in slot 0).
inline LeafSpec tree_2d(int top, int sb, int sc, int ab, int ac, int hbc, int hcb) {
  enum { FACE, EDGE_AB, EDGE_AC, VERT } leaf;
  if (ab) {
    if (!hbc) leaf = ac ? (!hcb ? FACE : EDGE_AC) : FACE;
    else leaf = EDGE_AB;
  } else if (ac) {
    leaf = !hcb ? FACE : EDGE_AC;
  } else {
    leaf = VERT;
  }
  LeafSpec l;
  switch (leaf) {
    case FACE:
      l = LeafSpec{3, {sc, sb, top, 0}, VK_PLANE, face_id(sb, sc), 1};
      break;
    case EDGE_AB:
      l = LeafSpec{2, {top, sb, 0, 0}, VK_LINE, sb, 1};
      break;
    case EDGE_AC:
      l = LeafSpec{2, {sc, top, 0, 0}, VK_LINE, sc, 1};
      break;
    default:
      l = LeafSpec{1, {top, 0, 0, 0}, VK_A, 0, 1};
      break;
  }
  fill_perm(l);
  return l;
}
}  // namespace detail

inline int h2_bit(int x, int y) {
  static const int t[3][3] = {{-1, 0, 2}, {1, -1, 4}, {3, 5, -1}};
  return t[x][y];
}

// 4-point tree (openGJK.c:315-613)
inline LeafSpec tree_3d(uint32_t index) {
  using namespace detail;
  int along[3], facing[3];
  for (int k = 0; k < 3; ++k) {
    along[k] = (index >> k) & 1;
    facing[k] = (index >> (3 + k)) & 1;
  }
  auto h2 = [&](int x, int y) { return (int)((index >> (6 + h2_bit(x, y))) & 1); };
  const int n_along = along[0] + along[1] + along[2];
  const int n_facing = facing[0] + facing[1] + facing[2];
  const int A = 3;

  if (n_along == 0) return leaf_vertex(A);  // openGJK.c:340-343
  if (n_facing == 3) {                       // origin enclosed, openGJK.c:356-358
    LeafSpec l = {4, {0, 1, 2, 3}, VK_ZERO, 0, 1};
    return l;
  }
  if (n_facing == 2) {  // openGJK.c:360-396: drop the slot opposite the outward face, then S2D
    int x, y;           // surviving old slots, x -> new slot 0, y -> new slot 1
    if (!facing[2]) { x = 0; y = 1; }
    else if (!facing[1]) { x = 0; y = 2; }
    else { x = 1; y = 2; }
    // after the drop: slot0 = x, slot1 = y, slot2 = a.  tree_2d leaves are in that 3-slot frame,
    // so build them with the OLD slot numbers as sources.
    LeafSpec l = tree_2d(A, y, x, along[y], along[x], h2(y, x), h2(x, y));
    // tree_2d wrote sources assuming b keeps "slot 1" and c keeps "slot 0" of the 3-frame:
    // FACE -> {x, y, a}; EDGE_AB -> {a, y}; EDGE_AC -> {x, a}; VERT -> {a}.  Already old-slot ids.
    return l;
  }
  int k, i, j;
  auto roles = [&](int kk) { k = kk; i = (kk + 2) % 3; j = (kk + 1) % 3; };
  if (n_facing == 1) {  // openGJK.c:397-536
    if (facing[2]) roles(2);
    else if (facing[1]) roles(1);
    else roles(0);
    if (n_along == 1) {
      if (along[k]) {
        if (!h2(k, i)) return leaf_face(A, i, k);
        if (!h2(k, j)) return leaf_face(A, j, k);
        return leaf_edge(A, k);
      }
      if (along[i]) return !h2(i, k) ? leaf_face(A, i, k) : leaf_edge(A, i);
      return !h2(j, k) ? leaf_face(A, j, k) : leaf_edge(A, j);
    }
    if (n_along == 2) {
      if (along[i]) {
        if (!h2(k, i)) return !h2(i, k) ? leaf_face(A, i, k) : leaf_edge(A, k);
        return !h2(k, j) ? leaf_face(A, j, k) : leaf_edge(A, k);
      }
      if (along[j]) {
        if (!h2(k, j)) return !h2(j, k) ? leaf_face(A, j, k) : leaf_edge(A, j);
        return !h2(k, i) ? leaf_face(A, i, k) : leaf_edge(A, k);
      }
      // unreachable ("ERROR", openGJK.c:494-496): nvrtx := 3, slots and v untouched
      LeafSpec l = {3, {0, 1, 2, 3}, VK_KEEP, 0, 0};
      return l;
    }
    // n_along == 3, openGJK.c:498-535
    const int ik = h2(i, k), jk = h2(j, k), ki = h2(k, i), kj = h2(k, j);
    if (ki && kj) return leaf_edge(A, k);
    if (ki) return jk ? leaf_edge(A, j) : leaf_face(A, j, k);
    return ik ? leaf_edge(A, i) : leaf_face(A, i, k);
  }
  // n_facing == 0, openGJK.c:538-608
  if (n_along == 1) {
    if (along[1]) roles(2);
    else if (along[0]) roles(1);
    else roles(0);
    if (!h2(i, j)) return leaf_face(A, i, j);
    if (!h2(i, k)) return leaf_face(A, i, k);
    return leaf_edge(A, i);
  }
  if (n_along == 2) {
    if (!along[1]) roles(2);
    else if (!along[0]) roles(1);
    else roles(0);
    if (!h2(j, k)) {
      if (!h2(k, j)) return leaf_face(A, j, k);
      if (!h2(k, i)) return leaf_face(A, i, k);
      return leaf_edge(A, k);
    }
    if (!h2(j, i)) return leaf_face(A, i, j);
    return leaf_edge(A, j);
  }
  // n_along == 3 with no face test firing: the reference leaves simplex and v untouched
  LeafSpec l = {4, {0, 1, 2, 3}, VK_KEEP, 0, 1};
  return l;
}

// 3-point table index: bit0 = along[1] (a->b), bit1 = along[0] (a->c), bit2 = hff2(a,b,c), bit3 = hff2(a,c,b)
inline LeafSpec tree_2d_index(uint32_t index) {
  return detail::tree_2d(2, 1, 0, index & 1, (index >> 1) & 1, (index >> 2) & 1, (index >> 3) & 1);
}

struct LeafTables {
  uint32_t t3[4096];
  uint32_t t2[16];
};

inline void build_leaf_tables(LeafTables& t) {
  for (uint32_t i = 0; i < 4096; ++i) t.t3[i] = encode_leaf(tree_3d(i));
  for (uint32_t i = 0; i < 16; ++i) t.t2[i] = encode_leaf(tree_2d_index(i));
}

// ---- unified 16-bit table (one lookup for 2-, 3- and 4-point simplices) --------------------------------------
// The lane-uniform iteration (gjk_advance_u, gjk_core.cuh) always treats the newest point `a` as slot 3 and the
// older points as slots 0..m-1 (m = 1..3), evaluates the 12 predicate bits of the 4-point index (bits that involve
// absent slots are garbage) and looks the leaf up at  kUnifiedBase[m] + (index & kUnifiedMask[m]):
//   m = 3 : the 4-point tree, all 12 bits;
//   m = 2 : the 3-point tree with b = slot 1, c = slot 0: along[1] = bit 1, along[0] = bit 0,
//           hff2(a,b,c) = bit 7 (pair (1,0)), hff2(a,c,b) = bit 6 (pair (0,1)); bits 2..5 ignored;
//   m = 1 : the 2-point rule (S1D, openGJK.c:257-268): bit 0 = hff1(a, slot 0).
// Leaf16 encoding: bits 0..1 vertices kept - 1 | bits 2..9 source slot of output slot j at bits 2+2j |
//                  bits 10..12 how v is obtained (VK_*) | bits 13..14 X | bit 15 the newest point survives.
constexpr int kUnifiedSize = 4096 + 256 + 2;
constexpr int kUnifiedBase1 = 4096 + 256, kUnifiedBase2 = 4096, kUnifiedBase3 = 0;

inline uint16_t encode_leaf16(const LeafSpec& l) {
  uint32_t e = (uint32_t)(l.nv - 1) & 3u;
  for (int j = 0; j < 4; ++j) e |= (uint32_t)(l.src[j] & 3) << (2 + 2 * j);
  e |= (uint32_t)(l.vkind & 7) << 10;
  e |= (uint32_t)(l.x & 3) << 13;
  e |= (uint32_t)(l.keeps_a & 1) << 15;
  return (uint16_t)e;
}

inline void build_unified_table(uint16_t* t /* [kUnifiedSize] */) {
  for (uint32_t i = 0; i < 4096; ++i) t[kUnifiedBase3 + i] = encode_leaf16(tree_3d(i));
  for (uint32_t i = 0; i < 256; ++i) {
    const int along0 = i & 1, along1 = (i >> 1) & 1, h01 = (i >> 6) & 1, h10 = (i >> 7) & 1;
    t[kUnifiedBase2 + i] = encode_leaf16(detail::tree_2d(3, 1, 0, along1, along0, h10, h01));
  }
  {
    LeafSpec keep = {2, {0, 3, 0, 0}, VK_LINE, 0, 1};  // both points stay, v = closest point on the segment
    fill_perm(keep);
    t[kUnifiedBase1 + 1] = encode_leaf16(keep);
    t[kUnifiedBase1 + 0] = encode_leaf16(detail::leaf_vertex(3));  // only the new point stays, v = a
  }
}

}  // namespace ogjk
