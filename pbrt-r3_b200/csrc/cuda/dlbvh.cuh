// dlbvh.cuh — the binary LinearBVHNode layout (`Accelerator "bvh" "string layout" "lbvh"`): closest-hit and any-hit
// traversal, two levels for ObjectInstance.
// Replaces reference src/accelerators/bvh/accel/lbvh.rs:14-19 (LinearBVHNode), :55-73 (flatten_bvh_tree: depth-first,
// first child = node + 1, second child = offset), :128-241 (LBVHAccel::intersect / intersect_p) and
// src/core/geometry/intersect.rs:46-65 (intersect_box_array_i).  The reference selects this accelerator on targets other
// than x86_64 / aarch64 (accel/mod.rs:4-10); here it is a second node layout next to the 4-wide one.
//
// Node, 32 B (the reference's is 48 B: usize offset / n_primitives): two 16-byte loads
//   f4 a = bounds.min.xyz, bits(offset)        f4 b = bounds.max.xyz, bits(n_primitives | axis << 30)
// offset = first primitive record (leaf) or second child (interior), relative to the object.
// Stack entries carry the parent's clipped slab range (node, t0, t1), exactly like the reference's Vec<(usize, f32, f32)>:
//   closest hit: t1 = min(t_max, t1) on pop, entry skipped when t1 < t0; any hit: the range is used as pushed.
// The visit order (near child first by the sign of d[axis]), the leaf order and every float operation follow the
// reference, so primitive ids, t and the node / triangle counters equal the CPU path's.
#pragma once
#include "dtrav.cuh"

namespace ptd {

struct LStack {  // per-thread, local memory: the LBVH layout is the compatibility path, the 4-wide layout is the fast one
  int node[PT_MAX_STACK];
  float t0[PT_MAX_STACK], t1[PT_MAX_STACK];
  int n;
  PT_HD bool push(int x, float a, float b) {
    if (n >= PT_MAX_STACK) return false;
    node[n] = x; t0[n] = a; t1[n] = b;
    n++;
    return true;
  }
};

template <bool ANY, bool INST, bool COUNT, int LEVEL>
PT_HD bool lbvh_traverse(const DScene& S, uint32_t obj, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, LStack& st, int* overflow, TravCounters* cnt);

template <bool ANY, bool INST, bool COUNT, int LEVEL>
PT_HD bool lbvh_record(const DScene& S, const f4* rec, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, LStack& st, int* overflow, TravCounters* cnt) {
  f4 v0 = ldg4(rec), v1 = ldg4(rec + 1), v2 = ldg4(rec + 2);
  uint32_t flags = fbits(v1.w);
  if (INST && LEVEL == 0 && (flags & PT_REC_INSTANCE)) {  // TransformedPrimitive::intersect[_p] (transformed_primitive.rs:26-46)
    int inst = (int)fbits(v0.w);
    const InstanceRec& I = S.instances[inst];
    DRay r;
    r.o = ro; r.d = rd; r.tmax = ray_tmax; r.time = rtime;
    DRay q = xform_ray(I.minv, r);
    float tm = q.tmax;
    HitD h;
    bool found = lbvh_traverse<ANY, false, COUNT, 1>(S, (uint32_t)I.object, q.o, q.d, rtime, tm, &h, st, overflow, cnt);
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

template <bool ANY, bool INST, bool COUNT, int LEVEL>
PT_HD bool lbvh_traverse(const DScene& S, uint32_t obj, f3 ro, f3 rd, float rtime, float& ray_tmax, HitD* hit, LStack& st, int* overflow, TravCounters* cnt) {
  const ObjectRec& O = S.objects[obj];
  const f4* recs = S.recs + (size_t)O.rec_base * 3;
  if (O.n_prims == 0) return false;
  if (O.single) return lbvh_record<ANY, INST, COUNT, LEVEL>(S, recs, ro, rd, rtime, ray_tmax, hit, st, overflow, cnt);
  const f4* nodes = S.lnodes + (size_t)O.node_base * 2;
  const float idx = 1.0f / rd.x, idy = 1.0f / rd.y, idz = 1.0f / rd.z;  // Float::recip
  // sign(idir[i]) (lbvh.rs:75-81): is_sign_negative of the reciprocal = of the component, -0.0 included
  const bool ngx = signbit_(idx), ngy = signbit_(idy), ngz = signbit_(idz);
  const int base = st.n;  // entries below belong to the enclosing level
  bool found = false;
  if (!st.push(0, 0.0f, ray_tmax)) { *overflow = 1; return false; }
  while (st.n > base) {
    st.n--;
    const int cur = st.node[st.n];
    float t0 = st.t0[st.n], t1 = st.t1[st.n];
    if (!ANY) {
      t1 = fminf(ray_tmax, t1);
      if (t1 < t0) continue;
    }
    if (COUNT) cnt->nodes++;
    const f4 a = ldg4(nodes + 2 * (size_t)cur), b = ldg4(nodes + 2 * (size_t)cur + 1);
    // intersect_box_array_i: f32::max / f32::min (NaN-ignoring, like fmaxf / fminf), x then y then z
    t0 = fmaxf(t0, ((ngx ? b.x : a.x) - ro.x) * idx);
    t1 = fminf(t1, ((ngx ? a.x : b.x) - ro.x) * idx);
    t0 = fmaxf(t0, ((ngy ? b.y : a.y) - ro.y) * idy);
    t1 = fminf(t1, ((ngy ? a.y : b.y) - ro.y) * idy);
    t0 = fmaxf(t0, ((ngz ? b.z : a.z) - ro.z) * idz);
    t1 = fminf(t1, ((ngz ? a.z : b.z) - ro.z) * idz);
    if (!(t0 <= t1)) continue;
    const uint32_t meta = fbits(b.w), np = meta & 0x3fffffffu, axis = meta >> 30;
    const int offset = (int)fbits(a.w);
    if (np > 0) {
      const f4* r = recs + (size_t)offset * 3;
      for (uint32_t i = 0; i < np; i++, r += 3) {
        if (lbvh_record<ANY, INST, COUNT, LEVEL>(S, r, ro, rd, rtime, ray_tmax, hit, st, overflow, cnt)) {
          if (ANY) { st.n = base; return true; }
          found = true;
        }
      }
    } else {
      const bool neg = axis == 0 ? ngx : (axis == 1 ? ngy : ngz);
      // index_children = [node + 1, offset]; the near one ([dir_is_neg]) is pushed last, so it is popped first
      const int near_c = neg ? offset : cur + 1, far_c = neg ? cur + 1 : offset;
      if (!st.push(far_c, t0, t1) || !st.push(near_c, t0, t1)) { *overflow = 1; st.n = base; return found; }
    }
  }
  return found;
}

}  // namespace ptd
