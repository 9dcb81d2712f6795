// dtrav2.cuh — resumable (flat state machine) form of the two-level 4-wide BVH traversal of dtrav.cuh, for
// persistent warps with dynamic ray fetch: a lane that finishes its ray pulls the next one instead of idling until
// the slowest ray of its warp is done, and the work is split in a NODE phase and a LEAF phase ("while-while") so that
// the lanes of a warp run the slab tests together and the triangle tests together (ncu on the nested-call if-if
// version: 6-15 of 32 lanes active per instruction).
// Same reference semantics as dtrav.cuh (qbvh_x86.rs:231-343,373-399; transformed_primitive.rs:26-46): identical
// node visit order, leaf order, t_max shrinking and float operations, so primitive ids and the node / triangle visit
// counters still equal the CPU path's.
//
// Stack entries: >= 0 node index (relative to the current object), < 0 inline leaf ~((first << 3) | (count - 1))
// (bit 30 = PT_LEAF_REMAINDER marks the rest of a top-level leaf that an instance interrupted: not a node visit),
// PT_STACK_BARRIER = "the instanced object's traversal ends here, return to the top level".  When an instance is
// met inside a top-level leaf, the leaf's remaining primitives are pushed back as inline-leaf entries below the
// barrier, so a leaf is processed in order even though the nested traversal runs in between.
#pragma once
#include "dtrav.cuh"

// tuning switches (A/B measured on B200, see profiles/r1_trace_ab.md)
#ifndef PT_OPT_TRI_PRED
#define PT_OPT_TRI_PRED 0
#endif
#ifndef PT_OPT_HOIST
#define PT_OPT_HOIST 1
#endif
#ifndef PT_OPT_FMNMX
#define PT_OPT_FMNMX 1
#endif
#ifndef PT_OPT_POS
#define PT_OPT_POS 1
#endif
#ifndef PT_OPT_POPLUT
#define PT_OPT_POPLUT 0
#endif
// conservative world-space cull of instance records before the exact (transformed) object-bounds test
#ifndef PT_INST_CULL
#define PT_INST_CULL 1
#endif
// triangle records in the three axis orders of the watertight test (dscene.cuh): no per-triangle component selects
#ifndef PT_TRI_PERM
#define PT_TRI_PERM 1
#endif

namespace ptd {

struct Trav {
  f3 o, d;              // ray in the current space (world, or the instanced object's space)
  float idx, idy, idz;  // 1 / d
  float tmin, tmax;     // slab range of the current level (qbvh_x86.rs:246-252, 279-283)
  float ray_tmax;       // Ray.t_max (Cell): shrinks on every accepted hit in closest-hit mode
  float time;
  int cur;              // entry being processed
  int sg;               // sign bits of d: bit0 x, bit1 y, bit2 z
  bool nonan;           // no slab product of this ray can be NaN: max_ps/min_ps == fmaxf/fminf (see trav_nodes)
  int kz;               // watertight test: dominant axis of d and the shear (pure functions of the ray, triangle.rs:258-277)
  float shx, shy, shz;
  int lf, lc;           // pending leaf: first record, count (lc = -1: the pending entry is the barrier)
  uint32_t node_base, rec_base;
  int inst;             // >= 0 while traversing an instanced object (the world-space ray state is parked in the stack's save area)
  bool leaf_hit;        // a primitive of the leaf being processed (at the current level) was accepted
  bool outer_hit;       // ... of the top-level leaf that was interrupted by the instance
  bool nested_hit;      // ... of some leaf of the instanced object
  bool found;
  bool any;             // any-hit query (intersect_p)
  bool done;
  HitD hit;
};

// trav_set_ray with the reciprocals already known (the instance entry computes them once for the object-bounds test)
PT_HD void trav_set_ray_idir(Trav& T, f3 o, f3 d, float idx, float idy, float idz);
PT_HD void trav_set_ray(Trav& T, f3 o, f3 d) { trav_set_ray_idir(T, o, d, 1.0f / d.x, 1.0f / d.y, 1.0f / d.z); }
PT_HD void trav_set_ray_idir(Trav& T, f3 o, f3 d, float idx, float idy, float idz) {
  T.o = o; T.d = d;
  T.idx = idx; T.idy = idy; T.idz = idz;
  T.sg = (signbit_(d.x) ? 1 : 0) | (signbit_(d.y) ? 2 : 0) | (signbit_(d.z) ? 4 : 0);
  // (box - o) * idir is NaN only for 0 * inf or a NaN operand: excluded when every 1/d is finite and non-zero and o is finite
  {
    const float big = 3.0e38f;
    T.nonan = fabsf(T.idx) < big && fabsf(T.idy) < big && fabsf(T.idz) < big && fabsf(T.idx) > 0.0f && fabsf(T.idy) > 0.0f && fabsf(T.idz) > 0.0f &&
              fabsf(o.x) < big && fabsf(o.y) < big && fabsf(o.z) < big;
  }
#if !PT_OPT_HOIST
  T.kz = 0; T.shx = T.shy = T.shz = 0.0f;
  return;
#endif
  // Triangle::intersect recomputes these for every triangle; they only depend on the ray
  int kz = max_dim(abs3(d));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float dx = comp(d, kx), dy = comp(d, ky), dz = comp(d, kz);
  T.kz = kz;
  T.shx = -dx / dz; T.shy = -dy / dz;
  T.shz = kz == 0 ? idx : (kz == 1 ? idy : idz);  // 1 / d[kz]: the same quotient as the slab reciprocal of that axis
}

// everything trav_set_ray derives from a ray <-> 15 words of the stack's save area
template <typename Stack>
PT_HD void trav_park_ray(const Trav& T, Stack& st) {
  st.save(0, T.o.x); st.save(1, T.o.y); st.save(2, T.o.z);
  st.save(3, T.d.x); st.save(4, T.d.y); st.save(5, T.d.z);
  st.save(6, T.tmin); st.save(7, T.tmax);
  st.save(8, T.idx); st.save(9, T.idy); st.save(10, T.idz);
  st.save(11, bitsf((uint32_t)T.sg | ((uint32_t)T.kz << 3) | (T.nonan ? 32u : 0u)));
  st.save(12, T.shx); st.save(13, T.shy); st.save(14, T.shz);
}
template <typename Stack>
PT_HD void trav_unpark_ray(Trav& T, const Stack& st) {
  T.o = mk3(st.saved(0), st.saved(1), st.saved(2));
  T.d = mk3(st.saved(3), st.saved(4), st.saved(5));
  T.tmin = st.saved(6); T.tmax = st.saved(7);
  T.idx = st.saved(8); T.idy = st.saved(9); T.idz = st.saved(10);
  const uint32_t pk = fbits(st.saved(11));
  T.sg = (int)(pk & 7u); T.kz = (int)((pk >> 3) & 3u); T.nonan = (pk & 32u) != 0u;
  T.shx = st.saved(12); T.shy = st.saved(13); T.shz = st.saved(14);
}
// bounds_range (dtrav.cuh) with the reciprocals given
PT_HD bool bounds_range_idir(const float* bmin, const float* bmax, f3 o, f3 d, float idx, float idy, float idz, float tmax, float* t0o, float* t1o) {
  float t0 = 0.0f, t1 = tmax;
  bool nx = signbit_(d.x), ny = signbit_(d.y), nz = signbit_(d.z);
  t0 = fmaxf(t0, ((nx ? bmax[0] : bmin[0]) - o.x) * idx);
  t1 = fminf(t1, ((nx ? bmin[0] : bmax[0]) - o.x) * idx);
  t0 = fmaxf(t0, ((ny ? bmax[1] : bmin[1]) - o.y) * idy);
  t1 = fminf(t1, ((ny ? bmin[1] : bmax[1]) - o.y) * idy);
  t0 = fmaxf(t0, ((nz ? bmax[2] : bmin[2]) - o.z) * idz);
  t1 = fminf(t1, ((nz ? bmin[2] : bmax[2]) - o.z) * idz);
  *t0o = t0;
  *t1o = t1;
  return t0 <= t1;
}

#if PT_TRI_PERM
// tri_test_ray on a record of copy kz: q0, q1, q2 hold (p[kx], p[ky], p[kz]) and op the ray origin in the same order, so the
// permutation of triangle.rs:258-268 is already done.  A cyclic axis rotation keeps every operand pair of the cross product,
// and the culling dot product is summed in the original x, y, z order: same operations, same results.
PT_HD bool tri_test_perm(f3 q0, f3 q1, f3 q2, uint32_t flags, const Trav& T, f3 op, TriHitD* out) {
  if (!(flags & PT_TRI_TWO_SIDED)) {
    f3 n = cross3(q0 - q2, q1 - q2);  // = (n[kx], n[ky], n[kz])
    if (((flags & PT_TRI_REVERSE_ORIENTATION) != 0) != ((flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    const int kz = T.kz;
    const int px = kz == 2 ? 0 : (kz == 0 ? 2 : 1), py = kz == 2 ? 1 : (kz == 0 ? 0 : 2), pz = kz == 2 ? 2 : (kz == 0 ? 1 : 0);  // where x, y, z are stored
    if (comp(n, px) * T.d.x + comp(n, py) * T.d.y + comp(n, pz) * T.d.z >= 0.0f) return false;
  }
  f3 a = q0 - op, b = q1 - op, c = q2 - op;
  float ax = a.x, ay = a.y, az = a.z;
  float bx = b.x, by = b.y, bz = b.z;
  float cx = c.x, cy = c.y, cz = c.z;
  const float sx = T.shx, sy = T.shy, sz = T.shz;
  ax += sx * az; ay += sy * az;
  bx += sx * bz; by += sy * bz;
  cx += sx * cz; cy += sy * cz;
  float e0 = bx * cy - by * cx;
  float e1 = cx * ay - cy * ax;
  float e2 = ax * by - ay * bx;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // rare: exact edge hit, f64 fallback (triangle.rs:283-293)
    e0 = (float)((double)cy * (double)bx - (double)cx * (double)by);
    e1 = (float)((double)ay * (double)cx - (double)ax * (double)cy);
    e2 = (float)((double)by * (double)ax - (double)bx * (double)ay);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  az *= sz; bz *= sz; cz *= sz;
  float ts = e0 * az + e1 * bz + e2 * cz;
  const float tmax = T.ray_tmax;
  if (det < 0.0f && (ts >= 0.0f || ts < tmax * det)) return false;
  else if (det > 0.0f && (ts <= 0.0f || ts > tmax * det)) return false;
  float inv_det = 1.0f / det;
  float t = ts * inv_det;
  float max_zt = fmaxf(fabsf(az), fmaxf(fabsf(bz), fabsf(cz)));
  float delta_z = PT_GAMMA(3.0f) * max_zt;
  float max_xt = fmaxf(fabsf(ax), fmaxf(fabsf(bx), fabsf(cx)));
  float max_yt = fmaxf(fabsf(ay), fmaxf(fabsf(by), fabsf(cy)));
  float delta_x = PT_GAMMA(5.0f) * (max_xt + max_zt);
  float delta_y = PT_GAMMA(5.0f) * (max_yt + max_zt);
  float delta_e = 2.0f * (PT_GAMMA(2.0f) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
  float max_e = fmaxf(fabsf(e0), fmaxf(fabsf(e1), fabsf(e2)));
  float delta_t = 3.0f * (PT_GAMMA(3.0f) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
  if (t <= delta_t) return false;
  out->t = t;
  out->b0 = e0 * inv_det;
  out->b1 = e1 * inv_det;
  out->b2 = e2 * inv_det;
  return true;
}
// the ray origin in the component order of copy T.kz
PT_HD f3 trav_origin_perm(const Trav& T) {
  const int kz = T.kz;
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  return mk3(comp(T.o, kx), comp(T.o, ky), comp(T.o, kz));
}
#endif

// tri_test (dtrav.cuh) with the per-ray constants hoisted and the rejection tests folded into one predicate, so that
// the lanes of a warp stay converged through the leaf loop.  Same operations, same order, same results.
PT_HD bool tri_test_ray(f3 p0, f3 p1, f3 p2, uint32_t flags, const Trav& T, TriHitD* out) {
  if (!(flags & PT_TRI_TWO_SIDED)) {
    f3 n = cross3(p0 - p2, p1 - p2);
    if (((flags & PT_TRI_REVERSE_ORIENTATION) != 0) != ((flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    if (dot3(n, T.d) >= 0.0f) return false;
  }
  f3 a = p0 - T.o, b = p1 - T.o, c = p2 - T.o;
  const int kz = T.kz;
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float ax = comp(a, kx), ay = comp(a, ky), az = comp(a, kz);
  float bx = comp(b, kx), by = comp(b, ky), bz = comp(b, kz);
  float cx = comp(c, kx), cy = comp(c, ky), cz = comp(c, kz);
  const float sx = T.shx, sy = T.shy, sz = T.shz;
  ax += sx * az; ay += sy * az;
  bx += sx * bz; by += sy * bz;
  cx += sx * cz; cy += sy * cz;
  float e0 = bx * cy - by * cx;
  float e1 = cx * ay - cy * ax;
  float e2 = ax * by - ay * bx;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // rare: exact edge hit, f64 fallback (triangle.rs:283-293)
    e0 = (float)((double)cy * (double)bx - (double)cx * (double)by);
    e1 = (float)((double)ay * (double)cx - (double)ax * (double)cy);
    e2 = (float)((double)by * (double)ax - (double)bx * (double)ay);
  }
#if PT_OPT_TRI_PRED
  bool ok = !((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f));
  float det = e0 + e1 + e2;
  ok = ok && det != 0.0f;
  az *= sz; bz *= sz; cz *= sz;
  float ts = e0 * az + e1 * bz + e2 * cz;
  const float tmax = T.ray_tmax;
  if (det < 0.0f && (ts >= 0.0f || ts < tmax * det)) ok = false;
  else if (det > 0.0f && (ts <= 0.0f || ts > tmax * det)) ok = false;
  if (!ok) return false;
#else
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  az *= sz; bz *= sz; cz *= sz;
  float ts = e0 * az + e1 * bz + e2 * cz;
  const float tmax = T.ray_tmax;
  if (det < 0.0f && (ts >= 0.0f || ts < tmax * det)) return false;
  else if (det > 0.0f && (ts <= 0.0f || ts > tmax * det)) return false;
#endif
  float inv_det = 1.0f / det;
  float t = ts * inv_det;
  float max_zt = fmaxf(fabsf(az), fmaxf(fabsf(bz), fabsf(cz)));
  float delta_z = PT_GAMMA(3.0f) * max_zt;
  float max_xt = fmaxf(fabsf(ax), fmaxf(fabsf(bx), fabsf(cx)));
  float max_yt = fmaxf(fabsf(ay), fmaxf(fabsf(by), fabsf(cy)));
  float delta_x = PT_GAMMA(5.0f) * (max_xt + max_zt);
  float delta_y = PT_GAMMA(5.0f) * (max_yt + max_zt);
  float delta_e = 2.0f * (PT_GAMMA(2.0f) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
  float max_e = fmaxf(fabsf(e0), fmaxf(fabsf(e1), fabsf(e2)));
  float delta_t = 3.0f * (PT_GAMMA(3.0f) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
  if (t <= delta_t) return false;
  out->t = t;
  out->b0 = e0 * inv_det;
  out->b1 = e1 * inv_det;
  out->b2 = e2 * inv_det;
  return true;
}

// QBVHAccel::intersect[_p] prologue on the top-level object
PT_HD void trav_begin(const DScene& S, Trav& T, f3 o, f3 d, float tmax, float time, bool any) {
  T.any = any; T.found = false; T.done = false; T.inst = -1; T.leaf_hit = false; T.outer_hit = false; T.nested_hit = false;
  T.ray_tmax = tmax; T.time = time;
  T.hit.tri = -1; T.hit.instance = -1;
  T.lf = 0; T.lc = 0;
  const ObjectRec& O = S.objects[0];
  if (O.n_prims == 0) { T.done = true; return; }
  trav_set_ray(T, o, d);
  if (!bounds_range(O.bmin, O.bmax, o, d, tmax, &T.tmin, &T.tmax)) { T.done = true; return; }
  T.node_base = O.node_base; T.rec_base = O.rec_base;
  T.cur = 0;
}

template <typename Stack>
PT_HD void trav_pop(Trav& T, Stack& st) {
  if (!st.pop(&T.cur)) T.done = true;
}

// one accepted-or-not triangle record in the current space; returns true when the query is finished (any-hit)
template <bool COUNT>
PT_HD bool trav_triangle(Trav& T, f4 v0, f4 v1, f4 v2, TravCounters* cnt
#if PT_TRI_PERM
                         , f3 op
#endif
) {
  if (COUNT) cnt->tris++;
  TriHitD th;
#if PT_TRI_PERM
  if (!tri_test_perm(xyz(v0), xyz(v1), xyz(v2), fbits(v1.w), T, op, &th)) return false;
#elif PT_OPT_HOIST
  if (!tri_test_ray(xyz(v0), xyz(v1), xyz(v2), fbits(v1.w), T, &th)) return false;
#else
  if (!tri_test(xyz(v0), xyz(v1), xyz(v2), fbits(v1.w), T.o, T.d, T.ray_tmax, &th)) return false;
#endif
  T.found = true;
  if (T.any) return true;
  T.ray_tmax = th.t;
  T.leaf_hit = true;
  T.hit.tri = (int)fbits(v0.w);
  T.hit.instance = T.inst;
  T.hit.t = th.t; T.hit.b0 = th.b0; T.hit.b1 = th.b1; T.hit.b2 = th.b2;
  return false;
}

// NODE phase, one step: processes the current entry - an inner node (slab tests, ordered pushes, pop; returns true:
// stay in the node phase), or a leaf, a leaf remainder or the barrier (left pending in T.lf / T.lc), or finds the
// stack empty (T.done) (returns false).
//
// Child order.  The reference pushes the hit children in ORDER_TABLE order and pops the last one pushed
// (qbvh_x86.rs:187-204,264-274).  With s = sign[axis_top] << 2 | sign[axis_left] << 1 | sign[axis_right] that order is
//   s_t = 0: [3 ^ s_r, 2 ^ s_r, 1 ^ s_l, 0 ^ s_l]        s_t = 1: [1 ^ s_l, 0 ^ s_l, 3 ^ s_r, 2 ^ s_r]
// restricted to the hit lanes (checked against all 128 table entries by tests/test_host_logic.py).  So the stack
// position of child k is popc(mask & before_k), before_k = the lanes pushed before k: every hit child is stored with
// one predicated st.shared, the one at the top goes straight to T.cur - no table lookup, no data-dependent loop.
template <bool INST, bool COUNT, typename Stack>
PT_HD bool trav_node_step(const DScene& S, Trav& T, Stack& st, int* overflow, TravCounters* cnt) {
  const uint32_t sx = (T.sg & 1) ? 48u : 0u, sy = (T.sg & 2) ? 48u : 0u, sz = (T.sg & 4) ? 48u : 0u;
  const int sg3 = 3 * T.sg;
  const char* nbase = reinterpret_cast<const char*>(S.nodes + T.node_base);
  {
    if (T.cur < 0) {  // inline leaf or leaf remainder
      unsigned e = (unsigned)~T.cur;
      if (COUNT && !(INST && (e & PT_LEAF_REMAINDER))) cnt->nodes++;  // a leaf remainder is not a node visit of the reference traversal
      if (INST) e &= ~PT_LEAF_REMAINDER;
      T.lf = (int)(e >> PT_LEAF_BITS);
      T.lc = (int)(e & ((1u << PT_LEAF_BITS) - 1u)) + 1;
      return false;
    }
    if (INST && T.cur == PT_STACK_BARRIER) { T.lc = -1; return false; }
    if (COUNT) cnt->nodes++;
    // the whole 128-byte node is requested at once (one L2 round trip); bboxes[sign][axis] / bboxes[1 - sign][axis]
    const char* nd = nbase + (size_t)(uint32_t)T.cur * sizeof(QNode);
    const f4 nx = ldg4(reinterpret_cast<const f4*>(nd + sx)), fx = ldg4(reinterpret_cast<const f4*>(nd + (48u - sx)));
    const f4 ny = ldg4(reinterpret_cast<const f4*>(nd + (16u + sy))), fy = ldg4(reinterpret_cast<const f4*>(nd + (64u - sy)));
    const f4 nz = ldg4(reinterpret_cast<const f4*>(nd + (32u + sz))), fz = ldg4(reinterpret_cast<const f4*>(nd + (80u - sz)));
    const i4 ch = ldgi4(reinterpret_cast<const i4*>(nd + 96));
    const i4 meta = ldgi4(reinterpret_cast<const i4*>(nd + 112));
    if (meta.x & 0x100) {  // leaf node in the reference's own form
      T.lf = meta.y;
      T.lc = meta.z;
      return false;
    }
    unsigned mask = 0;
#if PT_OPT_FMNMX
    if (T.nonan) {
      // without NaNs `a > b ? a : b` (max_ps) and fmaxf(a, b) differ only in the sign of a zero, which `tx >= tn` ignores:
      // FMNMX3 instead of FSETP + FSEL per slab
#define PT_LANE3(L, BIT)                                                                                              \
  {                                                                                                                   \
    float tn = fmaxf(fmaxf(T.tmin, (nx.L - T.o.x) * T.idx), fmaxf((ny.L - T.o.y) * T.idy, (nz.L - T.o.z) * T.idz));   \
    float tx = fminf(fminf(T.tmax, (fx.L - T.o.x) * T.idx), fminf((fy.L - T.o.y) * T.idy, (fz.L - T.o.z) * T.idz));   \
    if (tx >= tn) mask |= BIT;                                                                                        \
  }
      PT_LANE3(x, 1u) PT_LANE3(y, 2u) PT_LANE3(z, 4u) PT_LANE3(w, 8u)
#undef PT_LANE3
    } else
#endif
    {
#define PT_LANE2(L, BIT)                                                  \
  {                                                                       \
    float tn = T.tmin, tx = T.tmax, v;                                    \
    v = (nx.L - T.o.x) * T.idx; tn = PT_SSE_MAX(tn, v);                   \
    v = (fx.L - T.o.x) * T.idx; tx = PT_SSE_MIN(tx, v);                   \
    v = (ny.L - T.o.y) * T.idy; tn = PT_SSE_MAX(tn, v);                   \
    v = (fy.L - T.o.y) * T.idy; tx = PT_SSE_MIN(tx, v);                   \
    v = (nz.L - T.o.z) * T.idz; tn = PT_SSE_MAX(tn, v);                   \
    v = (fz.L - T.o.z) * T.idz; tx = PT_SSE_MIN(tx, v);                   \
    if (tx >= tn) mask |= BIT;                                            \
  }
      PT_LANE2(x, 1u) PT_LANE2(y, 2u) PT_LANE2(z, 4u) PT_LANE2(w, 8u)
#undef PT_LANE2
    }
#if PT_OPT_POS
    // NOTE (measured, profiles/r1_trace_ab.md): no `continue` / early loop-back in this body.  A lane that jumps to the
    // loop tail ahead of the others is never merged back (independent thread scheduling reconverges only at the
    // compiler's BSSY/BSYNC pairs), and the warp then runs the loop once per lane group.
    const unsigned s = ((unsigned)meta.w >> sg3) & 7u;
    const unsigned bl = (s & 4u) ? 0u : 0xcu, br = (s & 4u) ? 0x3u : 0u;  // lanes of the other pair that are pushed first
#if PT_OPT_POPLUT
    // popc of a 4-bit value, 2 bits per entry (POPC issues at 1/4 of the ALU rate); popc(15) wraps to 0
#define PT_POP4(v) ((int)((0x3ee9e994u >> (2u * (v))) & 3u))
    const int p0 = PT_POP4(mask & (bl | ((s & 2u) ? 0u : 2u))), p1 = PT_POP4(mask & (bl | ((s & 2u) ? 1u : 0u)));
    const int p2 = PT_POP4(mask & (br | ((s & 1u) ? 0u : 8u))), p3 = PT_POP4(mask & (br | ((s & 1u) ? 4u : 0u)));
    const int last = (mask == 15u ? 4 : PT_POP4(mask)) - 1;
#undef PT_POP4
#else
    const int p0 = popc_(mask & (bl | ((s & 2u) ? 0u : 2u))), p1 = popc_(mask & (bl | ((s & 2u) ? 1u : 0u)));
    const int p2 = popc_(mask & (br | ((s & 1u) ? 0u : 8u))), p3 = popc_(mask & (br | ((s & 1u) ? 4u : 0u)));
    const int last = popc_(mask) - 1;  // -1: no child hit
#endif
    const bool h0 = (mask & 1u) != 0u, h1 = (mask & 2u) != 0u, h2 = (mask & 4u) != 0u, h3 = (mask & 8u) != 0u;
    if (st.room(last)) {
      if (h0 && p0 < last) st.put(p0, ch.x);
      if (h1 && p1 < last) st.put(p1, ch.y);
      if (h2 && p2 < last) st.put(p2, ch.z);
      if (h3 && p3 < last) st.put(p3, ch.w);
      if (last > 0) st.advance(last);
    } else {
#pragma unroll 1
      for (int pos = 0; pos < last; pos++) {
        int c = (h0 && p0 == pos) ? ch.x : ((h1 && p1 == pos) ? ch.y : ((h2 && p2 == pos) ? ch.z : ch.w));
        if (!st.push(c)) *overflow = 1;
      }
    }
    int nxt = (h0 && p0 == last) ? ch.x : ((h1 && p1 == last) ? ch.y : ((h2 && p2 == last) ? ch.z : ch.w));
    bool out = false;
    if (last < 0) out = !st.pop(&nxt);
    if (out) { T.done = true; return false; }
    T.cur = nxt;
#else
    {
      unsigned si = ((unsigned)meta.w >> sg3) & 7u;
      uint32_t order = kOrderTable[mask * 8 + si];  // mask 0 -> empty list
      while ((order & 0x4) == 0) {
        int k = (int)(order & 3);
        int c = k == 0 ? ch.x : (k == 1 ? ch.y : (k == 2 ? ch.z : ch.w));
        if (!st.push(c)) *overflow = 1;
        order >>= 4;
      }
    }
    if (!st.pop(&T.cur)) { T.done = true; return false; }
#endif
  }
  return true;
}

// the node phase of one ray on its own (tests/hostcheck and the fallback driver)
template <bool INST, bool COUNT, typename Stack>
PT_HD void trav_nodes(const DScene& S, Trav& T, Stack& st, int* overflow, TravCounters* cnt) {
  while (trav_node_step<INST, COUNT>(S, T, st, overflow, cnt)) {}
}

// LEAF phase: the pending leaf (intersect_primitives[_p] over recs[lf .. lf + lc) in order) or the barrier, then pop.
template <bool INST, bool COUNT, typename Stack>
PT_HD void trav_leaf(const DScene& S, Trav& T, Stack& st, int* overflow, TravCounters* cnt) {
  if (INST && T.lc < 0) {  // TransformedPrimitive::intersect returns: back to world space
    trav_unpark_ray(T, st);
    const ObjectRec& O = S.objects[0];
    T.node_base = O.node_base; T.rec_base = O.rec_base;
    T.inst = -1;
    // the interrupted leaf reported a hit if any of its primitives (the instance included) did: tmax = r.t_max.
    // Its remainder, if any, is the next entry, so no node test can observe the early update.
    if (T.outer_hit || T.nested_hit) T.tmax = T.ray_tmax;
    T.outer_hit = false; T.nested_hit = false; T.leaf_hit = false;
    trav_pop(T, st);
    return;
  }
  const int first = T.lf, count = T.lc;
#if PT_TRI_PERM
  const f4* r = S.recs_perm + (size_t)T.kz * S.rec_stride + ((size_t)T.rec_base + (size_t)first) * 3;
  f3 op = trav_origin_perm(T);
#else
  const f4* r = S.recs + ((size_t)T.rec_base + (size_t)first) * 3;
#endif
  for (int i = 0; i < count; i++, r += 3) {
    f4 v0 = ldg4(r), v1 = ldg4(r + 1), v2 = ldg4(r + 2);
    if (INST && (fbits(v1.w) & PT_REC_INSTANCE)) {
      // TransformedPrimitive::intersect[_p] (only top-level leaves hold instances)
#if PT_INST_CULL
      // 5 of 6 instance records met by a ray of the 1000-instance scene fail the object-bounds test below, after the ray
      // transform and four IEEE divisions.  The record carries a CONSERVATIVE world-space box of the instance (grown far
      // beyond the rounding of the transform, device_api.cu): a world ray that misses it cannot pass the exact test, so
      // it is skipped without touching the instance (63 % of the records).  NaN slab products are dropped by
      // fmaxf / fminf, i.e. they loosen the test.  (Measured and rejected, profiles/r2b_ab_cfg3.txt: culling all records of
      // the leaf first and then looping over the survivors - k_trace 125.6 -> 140.9 ms.)
      {
        const bool nx = (T.sg & 1) != 0, ny = (T.sg & 2) != 0, nz = (T.sg & 4) != 0;
        const float t0 = fmaxf(fmaxf(((nx ? v2.x : v0.x) - T.o.x) * T.idx, ((ny ? v2.y : v0.y) - T.o.y) * T.idy), fmaxf(((nz ? v2.z : v0.z) - T.o.z) * T.idz, 0.0f));
        const float t1 = fminf(fminf(((nx ? v0.x : v2.x) - T.o.x) * T.idx, ((ny ? v0.y : v2.y) - T.o.y) * T.idy), fminf(((nz ? v0.z : v2.z) - T.o.z) * T.idz, T.ray_tmax));
        if (t0 > t1) continue;
      }
#endif
      int inst = (int)fbits(v0.w);
      DRay wr;
      wr.o = T.o; wr.d = T.d; wr.tmax = T.ray_tmax; wr.time = T.time;
      const InstanceRec& I = S.instances[inst];
      const ObjectRec& O = S.objects[I.object];
      DRay q = xform_ray(I.minv, wr);
      if (O.n_prims == 0) continue;
      if (O.single) {  // one primitive, no accelerator: test it in object space right here
        trav_park_ray(T, st);
        trav_set_ray(T, q.o, q.d);
        T.inst = inst;
#if PT_TRI_PERM
        const f4* pr = S.recs_perm + (size_t)T.kz * S.rec_stride + (size_t)O.rec_base * 3;
        bool fin = trav_triangle<COUNT>(T, ldg4(pr), ldg4(pr + 1), ldg4(pr + 2), cnt, trav_origin_perm(T));
#else
        const f4* pr = S.recs + (size_t)O.rec_base * 3;
        bool fin = trav_triangle<COUNT>(T, ldg4(pr), ldg4(pr + 1), ldg4(pr + 2), cnt);
#endif
        trav_unpark_ray(T, st);
        T.inst = -1;
        if (fin) { T.done = true; return; }
        continue;
      }
      float tn, tx;
      const float qix = 1.0f / q.d.x, qiy = 1.0f / q.d.y, qiz = 1.0f / q.d.z;  // shared by the bounds test and the traversal
      if (!bounds_range_idir(O.bmin, O.bmax, q.o, q.d, qix, qiy, qiz, T.ray_tmax, &tn, &tx)) continue;
      // push back the rest of this leaf (in order, chunks of <= 8), then the barrier, and descend
      int rest = count - (i + 1), rfirst = first + i + 1;
      while (rest > 0) {  // last chunk first so that the chunks pop in order
        int nchunks = (rest + 7) >> 3;
        int cfirst = rfirst + (nchunks - 1) * 8, ccount = rest - (nchunks - 1) * 8;
        if (!st.push(~(int)(PT_LEAF_REMAINDER | ((unsigned)cfirst << PT_LEAF_BITS) | (unsigned)(ccount - 1)))) *overflow = 1;
        rest -= ccount;
      }
      if (!st.push(PT_STACK_BARRIER)) *overflow = 1;
      trav_park_ray(T, st);
      trav_set_ray_idir(T, q.o, q.d, qix, qiy, qiz);
      T.tmin = tn; T.tmax = tx;
      T.node_base = O.node_base; T.rec_base = O.rec_base;
      T.inst = inst;
      T.outer_hit = T.leaf_hit;
      T.leaf_hit = false;
      T.nested_hit = false;
      T.cur = 0;
      return;
    }
#if PT_TRI_PERM
    if (trav_triangle<COUNT>(T, v0, v1, v2, cnt, op)) { T.done = true; return; }
#else
    if (trav_triangle<COUNT>(T, v0, v1, v2, cnt)) { T.done = true; return; }
#endif
  }
  if (T.leaf_hit) {  // qbvh_x86.rs:279-283: tmax = r.t_max after a leaf that reported a hit
    T.tmax = T.ray_tmax;
    T.leaf_hit = false;
    if (INST && T.inst >= 0) T.nested_hit = true;
  }
  trav_pop(T, st);
}

}  // namespace ptd
