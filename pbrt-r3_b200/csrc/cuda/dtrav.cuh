// dtrav.cuh — watertight ray/triangle test and 4-wide BVH traversal (closest hit and any hit), two levels
// for ObjectInstance.
// Replaces reference src/shapes/triangle.rs:226-352,452-577 (Triangle::intersect / intersect_p decision part),
// src/accelerators/bvh/accel/qbvh/qbvh_x86.rs:26-67,187-343,373-399 (test_aabb, ORDER_TABLE, intersect_simd[_p],
// QBVHAccel::intersect[_p]), src/core/geometry/intersect.rs:46-65 (world-bound slab test),
// src/core/primitive/transformed_primitive.rs:26-46.
// Traversal order, tie-breaking and every float operation follow the reference so that primitive ids match
// bit for bit: separate sub and mul (no FMA), SSE max/min operand order, far-to-near push, leaf order.
#pragma once
#include "dscene.cuh"

namespace ptd {

struct TriHitD {
  float t, b0, b1, b2;
};

PT_HD bool tri_test(f3 p0, f3 p1, f3 p2, uint32_t flags, f3 ro, f3 rd, float tmax, TriHitD* out) {
  f3 dp02 = p0 - p2, dp12 = p1 - p2;
  if (!(flags & PT_TRI_TWO_SIDED)) {
    f3 n = cross3(dp02, dp12);
    if (((flags & PT_TRI_REVERSE_ORIENTATION) != 0) != ((flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    if (dot3(n, rd) >= 0.0f) return false;
  }
  f3 a = p0 - ro, b = p1 - ro, c = p2 - ro;
  int kz = max_dim(abs3(rd));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float dx = comp(rd, kx), dy = comp(rd, ky), dz = comp(rd, kz);
  float ax = comp(a, kx), ay = comp(a, ky), az = comp(a, kz);
  float bx = comp(b, kx), by = comp(b, ky), bz = comp(b, kz);
  float cx = comp(c, kx), cy = comp(c, ky), cz = comp(c, kz);
  float sx = -dx / dz, sy = -dy / dz, sz = 1.0f / dz;
  ax += sx * az; ay += sy * az;
  bx += sx * bz; by += sy * bz;
  cx += sx * cz; cy += sy * cz;
  float e0 = bx * cy - by * cx;
  float e1 = cx * ay - cy * ax;
  float e2 = ax * by - ay * bx;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)((double)cy * (double)bx - (double)cx * (double)by);
    e1 = (float)((double)ay * (double)cx - (double)ax * (double)cy);
    e2 = (float)((double)by * (double)ax - (double)bx * (double)ay);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  az *= sz; bz *= sz; cz *= sz;
  float ts = e0 * az + e1 * bz + e2 * cz;
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

// little-endian nibble lists, terminated by nibble 4 (qbvh_x86.rs:187-204); index = hit_mask * 8 + sign bits
#if defined(__CUDACC__)
__device__ __constant__
#endif
static const uint32_t kOrderTable[128] = {
    0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44440, 0x44440, 0x44440, 0x44440, 0x44440, 0x44440,
    0x44440, 0x44440, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44401, 0x44401, 0x44410, 0x44410,
    0x44401, 0x44401, 0x44410, 0x44410, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44402, 0x44402,
    0x44402, 0x44402, 0x44420, 0x44420, 0x44420, 0x44420, 0x44412, 0x44412, 0x44412, 0x44412, 0x44421, 0x44421, 0x44421, 0x44421,
    0x44012, 0x44012, 0x44102, 0x44102, 0x44201, 0x44201, 0x44210, 0x44210, 0x44443, 0x44443, 0x44443, 0x44443, 0x44443, 0x44443,
    0x44443, 0x44443, 0x44403, 0x44403, 0x44403, 0x44403, 0x44430, 0x44430, 0x44430, 0x44430, 0x44413, 0x44413, 0x44413, 0x44413,
    0x44431, 0x44431, 0x44431, 0x44431, 0x44013, 0x44013, 0x44103, 0x44103, 0x44301, 0x44301, 0x44310, 0x44310, 0x44423, 0x44432,
    0x44423, 0x44432, 0x44423, 0x44432, 0x44423, 0x44432, 0x44023, 0x44032, 0x44023, 0x44032, 0x44230, 0x44320, 0x44230, 0x44320,
    0x44123, 0x44132, 0x44123, 0x44132, 0x44231, 0x44321, 0x44231, 0x44321, 0x40123, 0x40132, 0x41023, 0x41032, 0x42301, 0x43201,
    0x42310, 0x43210};

struct HitD {
  int tri;       // global triangle id, -1 = miss
  int instance;  // -1 = top level
  float t, b0, b1, b2;
};

struct TravCounters {
  unsigned long long nodes, tris;
};

// world-bound slab test with f32::max / f32::min semantics (intersect.rs:46-65, bounds3.rs:154-163)
PT_HD bool bounds_range(const float* bmin, const float* bmax, f3 o, f3 d, float tmax, float* t0o, float* t1o) {
  float t0 = 0.0f, t1 = tmax;
  float idx = 1.0f / d.x, idy = 1.0f / d.y, idz = 1.0f / d.z;
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

#define PT_SSE_MAX(a, b) ((a) > (b) ? (a) : (b))
#define PT_SSE_MIN(a, b) ((a) < (b) ? (a) : (b))

// Stack policy: `Stack` provides push(int)/pop(int&)/empty(); see kernels for the shared-memory backed one.
struct LocalStack {
  int v[PT_MAX_STACK];
  int n;
  PT_HD void reset() { n = 0; }
  PT_HD bool push(int x) { if (n >= PT_MAX_STACK) return false; v[n++] = x; return true; }
  PT_HD bool pop(int* x) { if (n == 0) return false; *x = v[--n]; return true; }
};

// One level of traversal.  LEVEL 0 = top-level object (may contain instances when INST), LEVEL 1 = instanced object.
// ANY: any-hit (intersect_p) — returns at the first accepted primitive and never shrinks tmax.
// The ray's tmax is passed by reference (Cell semantics of the reference's Ray.t_max).
template <bool ANY, bool INST, bool COUNT, int LEVEL, typename Stack>
PT_HD bool traverse_object(const DScene& S, uint32_t obj, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, Stack& st, int* overflow,
                           TravCounters* cnt);

template <bool ANY, bool INST, bool COUNT, int LEVEL, typename Stack>
PT_HD bool test_record(const DScene& S, const f4* rec, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, Stack& st, int* overflow,
                       TravCounters* cnt) {
  f4 v0 = ldg4(rec), v1 = ldg4(rec + 1), v2 = ldg4(rec + 2);
  uint32_t flags = fbits(v1.w);
  if (INST && LEVEL == 0 && (flags & PT_REC_INSTANCE)) {
    // TransformedPrimitive::intersect: inverse-transform the ray, traverse the object, keep its t_max
    int inst = (int)fbits(v0.w);
    const InstanceRec& I = S.instances[inst];
    DRay r;
    r.o = ro; r.d = rd; r.tmax = ray_tmax; r.time = rtime;
    DRay q = xform_ray(I.minv, r);
    float tm = q.tmax;
    HitD h;
    // the nested traversal continues on the same stack above a barrier entry
    bool found = traverse_object<ANY, false, COUNT, 1, Stack>(S, (uint32_t)I.object, q.o, q.d, rtime, tm, &h, st, overflow, cnt);
    if (!found) return false;
    if (!ANY) {
      ray_tmax = tm;
      *hit = h;
      hit->instance = inst;
    }
    return true;
  }
  if (COUNT) cnt->tris++;
  TriHitD th;
  if (!tri_test(xyz(v0), xyz(v1), xyz(v2), flags, ro, rd, ray_tmax, &th)) return false;
  if (!ANY) {
    ray_tmax = th.t;
    hit->tri = (int)fbits(v0.w);
    hit->instance = -1;
    hit->t = th.t; hit->b0 = th.b0; hit->b1 = th.b1; hit->b2 = th.b2;
  }
  return true;
}

#define PT_STACK_BARRIER ((int)0x7fffffff)

template <bool ANY, bool INST, bool COUNT, int LEVEL, typename Stack>
PT_HD bool traverse_object(const DScene& S, uint32_t obj, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, Stack& st, int* overflow,
                           TravCounters* cnt) {
  const ObjectRec& O = S.objects[obj];
  const f4* recs = S.recs + (size_t)O.rec_base * 3;
  if (O.n_prims == 0) return false;
  if (O.single) return test_record<ANY, INST, COUNT, LEVEL, Stack>(S, recs, ro, rd, rtime, ray_tmax, hit, st, overflow, cnt);
  float tmin, tmax;
  if (!bounds_range(O.bmin, O.bmax, ro, rd, ray_tmax, &tmin, &tmax)) return false;
  const QNode* nodes = S.nodes + O.node_base;
  const float idx = 1.0f / rd.x, idy = 1.0f / rd.y, idz = 1.0f / rd.z;
  const int sgx = signbit_(rd.x) ? 1 : 0, sgy = signbit_(rd.y) ? 1 : 0, sgz = signbit_(rd.z) ? 1 : 0;
  bool found = false;
  if (LEVEL == 1 && !st.push(PT_STACK_BARRIER)) { *overflow = 1; return false; }
  int cur = 0;  // node 0 = root
  bool have = true;
  while (have) {
    if (COUNT) cnt->nodes++;
    int first = -1, count = 0;
    if (cur >= 0) {
      const QNode* nd = nodes + cur;
      i4 meta = ldgi4(&nd->meta);
      if (meta.x & 0x100) {  // leaf node in the reference's own form (root leaf, or > 8 primitives)
        first = meta.y;
        count = meta.z;
      } else {
        // test_aabb: lane k, per axis: tmin = max(tmin, (near - o) * idir); tmax = min(tmax, (far - o) * idir)
        f4 nx = ldg4(sgx ? &nd->hi[0] : &nd->lo[0]), fx = ldg4(sgx ? &nd->lo[0] : &nd->hi[0]);
        f4 ny = ldg4(sgy ? &nd->hi[1] : &nd->lo[1]), fy = ldg4(sgy ? &nd->lo[1] : &nd->hi[1]);
        f4 nz = ldg4(sgz ? &nd->hi[2] : &nd->lo[2]), fz = ldg4(sgz ? &nd->lo[2] : &nd->hi[2]);
        i4 ch = ldgi4(&nd->child);
        unsigned mask = 0;
#define PT_LANE(L, BIT)                                                   \
  {                                                                       \
    float tn = tmin, tx = tmax, v;                                        \
    v = (nx.L - ro.x) * idx; tn = PT_SSE_MAX(tn, v);                      \
    v = (fx.L - ro.x) * idx; tx = PT_SSE_MIN(tx, v);                      \
    v = (ny.L - ro.y) * idy; tn = PT_SSE_MAX(tn, v);                      \
    v = (fy.L - ro.y) * idy; tx = PT_SSE_MIN(tx, v);                      \
    v = (nz.L - ro.z) * idz; tn = PT_SSE_MAX(tn, v);                      \
    v = (fz.L - ro.z) * idz; tx = PT_SSE_MIN(tx, v);                      \
    if (tx >= tn) mask |= BIT;                                            \
  }
        PT_LANE(x, 1u) PT_LANE(y, 2u) PT_LANE(z, 4u) PT_LANE(w, 8u)
#undef PT_LANE
        if (mask) {
          int at = meta.x & 3, al = (meta.x >> 2) & 3, ar = (meta.x >> 4) & 3;
          int s_t = at == 0 ? sgx : (at == 1 ? sgy : sgz);
          int s_l = al == 0 ? sgx : (al == 1 ? sgy : sgz);
          int s_r = ar == 0 ? sgx : (ar == 1 ? sgy : sgz);
          uint32_t order = kOrderTable[mask * 8 + (unsigned)((s_t << 2) | (s_l << 1) | s_r)];
          while ((order & 0x4) == 0) {
            int k = (int)(order & 3);
            int c = k == 0 ? ch.x : (k == 1 ? ch.y : (k == 2 ? ch.z : ch.w));
            if (c != PT_EMPTY) {
              if (!st.push(c)) { *overflow = 1; }
            }
            order >>= 4;
          }
        }
      }
    } else {  // inline leaf reference
      unsigned e = (unsigned)~cur;
      first = (int)(e >> PT_LEAF_BITS);
      count = (int)(e & ((1u << PT_LEAF_BITS) - 1u)) + 1;
    }
    if (count > 0) {
      bool any = false;
      const f4* r = recs + (size_t)first * 3;
      for (int i = 0; i < count; i++, r += 3) {
        if (test_record<ANY, INST, COUNT, LEVEL, Stack>(S, r, ro, rd, rtime, ray_tmax, hit, st, overflow, cnt)) {
          if (ANY) {
            if (LEVEL == 1) {  // unwind this level's entries
              int x;
              while (st.pop(&x) && x != PT_STACK_BARRIER) {}
            }
            return true;
          }
          any = true;
        }
      }
      if (any) { tmax = ray_tmax; found = true; }
    }
    have = st.pop(&cur);
    if (LEVEL == 1 && have && cur == PT_STACK_BARRIER) have = false;
  }
  return found;
}

// Scene::intersect / Scene::intersect_p on the top-level object
template <bool INST, bool COUNT, typename Stack>
PT_HD bool scene_intersect(const DScene& S, f3 ro, f3 rd, float rtime, float& tmax, HitD* hit, Stack& st, int* overflow, TravCounters* cnt) {
  hit->tri = -1;
  hit->instance = -1;
  return traverse_object<false, INST, COUNT, 0, Stack>(S, 0u, ro, rd, rtime, tmax, hit, st, overflow, cnt);
}
template <bool INST, bool COUNT, typename Stack>
PT_HD bool scene_intersect_p(const DScene& S, f3 ro, f3 rd, float rtime, float tmax, Stack& st, int* overflow, TravCounters* cnt) {
  HitD h;
  return traverse_object<true, INST, COUNT, 0, Stack>(S, 0u, ro, rd, rtime, tmax, &h, st, overflow, cnt);
}

}  // namespace ptd
