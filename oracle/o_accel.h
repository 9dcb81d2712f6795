// ORACLE — TEST INFRASTRUCTURE ONLY (see o_math.h header).
// o_accel.h — triangle intersection, SAH BVH build, QBVH flatten + traversal, instances.
// Follows reference src/shapes/triangle.rs:226-577, src/accelerators/bvh/build/{node,sah,equal_counts,
// middle,types}.rs, src/accelerators/bvh/accel/qbvh/qbvh_x86.rs, src/core/geometry/intersect.rs,
// src/core/primitive/{geometric_primitive,transformed_primitive}.rs.
#pragma once
#include "../include/pbrt_r3_b200.h"
#include "o_math.h"

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

namespace orc {

struct Counters {
  uint64_t nodes = 0, tris = 0, regular_rays = 0, shadow_rays = 0, camera_rays = 0, vertices = 0;
};

struct TriRef {  // view of one triangle of the scene description
  Vec3 p0, p1, p2;
  uint32_t flags;
};
inline TriRef get_tri(const pt_scene_desc* d, int t) {
  const float* p = d->tri_p + (size_t)t * 9;
  TriRef r;
  r.p0 = Vec3(p[0], p[1], p[2]);
  r.p1 = Vec3(p[3], p[4], p[5]);
  r.p2 = Vec3(p[6], p[7], p[8]);
  r.flags = d->tri_flags[t];
  return r;
}

struct TriHit {
  Float t, b0, b1, b2;
};

// Triangle::intersect / intersect_p decision part (triangle.rs:226-352 / 452-577): identical arithmetic.
inline bool tri_intersect(const TriRef& T, const Ray& r, TriHit* out) {
  Vec3 p0 = T.p0, p1 = T.p1, p2 = T.p2;
  Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
  Vec3 n = cross(dp02, dp12);
  bool flip = ((T.flags & PT_TRI_REVERSE_ORIENTATION) != 0) ^ ((T.flags & PT_TRI_SWAPS_HANDEDNESS) != 0);
  if (flip) n = n * -1.0f;
  if (!(T.flags & PT_TRI_TWO_SIDED)) {
    if (dot(n, r.d) >= 0.0f) return false;
  }
  Vec3 p0t = p0 - r.o, p1t = p1 - r.o, p2t = p2 - r.o;
  int kz = max_dimension(vabs(r.d));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  Vec3 d = permute(r.d, kx, ky, kz);
  p0t = permute(p0t, kx, ky, kz);
  p1t = permute(p1t, kx, ky, kz);
  p2t = permute(p2t, kx, ky, kz);
  Float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
  p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
  p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
  p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
  Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
    e0 = (Float)(p2typ1tx - p2txp1ty);
    double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
    e1 = (Float)(p0typ2tx - p0txp2ty);
    double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
    e2 = (Float)(p1typ0tx - p1txp0ty);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  Float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
  Float t_max = r.t_max;
  Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
  else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
  Float inv_det = 1.0f / det;
  Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
  Float t = t_scaled * inv_det;
  Float max_zt = max_component(vabs(Vec3(p0t.z, p1t.z, p2t.z)));
  Float delta_z = GAMMA3 * max_zt;
  Float max_xt = max_component(vabs(Vec3(p0t.x, p1t.x, p2t.x)));
  Float max_yt = max_component(vabs(Vec3(p0t.y, p1t.y, p2t.y)));
  Float delta_x = GAMMA5 * (max_xt + max_zt);
  Float delta_y = GAMMA5 * (max_yt + max_zt);
  Float delta_e = 2.0f * (GAMMA2 * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
  Float max_e = max_component(vabs(Vec3(e0, e1, e2)));
  Float delta_t = 3.0f * (GAMMA3 * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
  if (t <= delta_t) return false;
  out->t = t; out->b0 = b0; out->b1 = b1; out->b2 = b2;
  return true;
}

// ---- BVH build (build/node.rs, sah.rs, equal_counts.rs, middle.rs) ---------------------------------
struct PrimInfo {
  size_t number;
  Bounds3 bounds;
  Vec3 centroid;
};
struct BuildNode {
  Bounds3 bounds;
  std::unique_ptr<BuildNode> c[2];
  int split_axis = 0;
  size_t first = 0, n = 0;
};

// ---- HLBVH (build/hlbvh.rs): Morton codes of the centroids, LBVH treelets over the top 12 bits, SAH over the treelets ----
struct MortonPrim { uint32_t index, code; };
inline uint32_t left_shift3(uint32_t x) {  // hlbvh.rs:23-37
  if (x >= 1024u) x = 1023u;
  x = (x | (x << 16)) & 0x30000ffu;
  x = (x | (x << 8)) & 0x300f00fu;
  x = (x | (x << 4)) & 0x30c30c3u;
  x = (x | (x << 2)) & 0x9249249u;
  return x;
}
struct HlbvhBuilder {
  size_t max_prims;
  const PrimInfo* info;
  std::vector<size_t>* ordered;
  std::unique_ptr<BuildNode> leaf(const MortonPrim* mp, size_t n) {
    auto nd = std::make_unique<BuildNode>();
    nd->first = ordered->size();
    nd->n = n;
    Bounds3 b = info[mp[0].index].bounds;
    ordered->push_back(mp[0].index);
    for (size_t i = 1; i < n; i++) { ordered->push_back(mp[i].index); b = bunion(b, info[mp[i].index].bounds); }
    nd->bounds = b;
    return nd;
  }
  static std::unique_ptr<BuildNode> interior(int axis, std::unique_ptr<BuildNode> a, std::unique_ptr<BuildNode> b) {
    auto nd = std::make_unique<BuildNode>();
    nd->bounds = bunion(a->bounds, b->bounds);
    nd->split_axis = axis;
    nd->c[0] = std::move(a);
    nd->c[1] = std::move(b);
    return nd;
  }
  std::unique_ptr<BuildNode> split_node(MortonPrim* mp, size_t n, int dim) {  // :86-137: equal counts, axes cycling z, y, x
    if (n <= max_prims) return leaf(mp, n);
    std::stable_sort(mp, mp + n, [&](const MortonPrim& a, const MortonPrim& b) { return info[a.index].centroid[dim] < info[b.index].centroid[dim]; });
    int next_dim = (dim + 3 - 1) % 3;
    size_t split = n / 2;
    auto n0 = split_node(mp, split, next_dim);
    auto n1 = split_node(mp + split, n - split, next_dim);
    return interior(dim, std::move(n0), std::move(n1));
  }
  std::unique_ptr<BuildNode> emit_lbvh(const MortonPrim* mp, size_t n, int bit_index) {  // :139-226
    if (n <= max_prims) return leaf(mp, n);
    if (bit_index == -1) {
      std::vector<MortonPrim> copy(mp, mp + n);
      return split_node(copy.data(), n, 2);
    }
    uint32_t mask = 1u << bit_index;
    if ((mp[0].code & mask) == (mp[n - 1].code & mask)) return emit_lbvh(mp, n, bit_index - 1);
    size_t s0 = 0, s1 = n - 1;
    while (s0 + 1 != s1) {
      size_t mid = (s0 + s1) / 2;
      if ((mp[s0].code & mask) == (mp[mid].code & mask)) s0 = mid; else s1 = mid;
    }
    auto n0 = emit_lbvh(mp, s1, bit_index - 1);
    auto n1 = emit_lbvh(mp + s1, n - s1, bit_index - 1);
    return interior(bit_index % 3, std::move(n0), std::move(n1));
  }
  static std::unique_ptr<BuildNode> upper_sah(std::vector<std::unique_ptr<BuildNode>> roots) {  // :234-330
    if (roots.empty()) throw std::runtime_error("hlbvh: build_upper_sah on an empty side (the reference panics here)");
    if (roots.size() <= 1) return std::move(roots[0]);
    Bounds3 bounds = roots[0]->bounds;
    for (size_t i = 1; i < roots.size(); i++) bounds = bunion(bounds, roots[i]->bounds);
    Vec3 center = (roots[0]->bounds.mn + roots[0]->bounds.mx) * 0.5f;
    Bounds3 cb(center, center);
    for (size_t i = 1; i < roots.size(); i++) cb = bunion_p(cb, (roots[i]->bounds.mn + roots[i]->bounds.mx) * 0.5f);
    int dim = cb.maximum_extent();
    const int NB = 12;
    uint32_t count[NB];
    Bounds3 bb[NB];
    for (int i = 0; i < NB; i++) count[i] = 0;
    auto bucket = [&](const BuildNode& nd) {
      Float c = (nd.bounds.mn[dim] + nd.bounds.mx[dim]) * 0.5f;
      size_t b = (size_t)f2u((Float)NB * ((c - cb.mn[dim]) / (cb.mx[dim] - cb.mn[dim])));
      return (int)std::min(b, (size_t)NB - 1);
    };
    for (auto& r : roots) {
      int b = bucket(*r);
      count[b]++;
      if (count[b] == 1) bb[b] = r->bounds; else bb[b] = bunion(bb[b], r->bounds);
    }
    Float cost[NB - 1];
    for (int i = 0; i < NB - 1; i++) {  // as written in the reference: both sides start from bucket i, count1 starts at 1
      Bounds3 b0 = bb[i], b1 = bb[i];
      uint32_t c0 = 0, c1 = 1;
      for (int j = 0; j < i; j++) { b0 = bunion(b0, bb[j]); c0 += count[j]; }
      for (int j = i + 1; j < NB; j++) { b1 = bunion(b1, bb[j]); c1 += count[j]; }
      cost[i] = 0.125f + (((Float)c0 * b0.surface_area() + (Float)c1 * b1.surface_area()) / bounds.surface_area());
    }
    Float min_cost = cost[0];
    int min_b = 0;
    for (int i = 1; i < NB - 1; i++)
      if (cost[i] < min_cost) { min_cost = cost[i]; min_b = i; }
    std::vector<std::unique_ptr<BuildNode>> c0, c1;
    for (auto& r : roots) (bucket(*r) <= min_b ? c0 : c1).push_back(std::move(r));
    auto n0 = upper_sah(std::move(c0));
    auto n1 = upper_sah(std::move(c1));
    return interior(dim, std::move(n0), std::move(n1));
  }
  std::unique_ptr<BuildNode> build(size_t n) {  // hlbvh_build (:332-428)
    Bounds3 bounds = info[0].bounds;
    for (size_t i = 1; i < n; i++) bounds = bunion(bounds, info[i].bounds);
    std::vector<MortonPrim> mp(n);
    for (size_t i = 0; i < n; i++) {
      mp[i].index = (uint32_t)info[i].number;
      Vec3 o = bounds.offset(info[i].centroid);
      Float sc[3] = {rclamp(o.x, 0.0f, 1.0f) * 1024.0f, rclamp(o.y, 0.0f, 1.0f) * 1024.0f, rclamp(o.z, 0.0f, 1.0f) * 1024.0f};
      mp[i].code = (left_shift3(f2u(ceilf(sc[2]))) << 2) | (left_shift3(f2u(ceilf(sc[1]))) << 1) | left_shift3(f2u(ceilf(sc[0])));
    }
    std::stable_sort(mp.begin(), mp.end(), [](const MortonPrim& a, const MortonPrim& b) { return a.code < b.code; });  // LSD radix sort = stable
    std::vector<std::unique_ptr<BuildNode>> treelets;
    const uint32_t MASK = 0x3ffc0000u;
    for (size_t start = 0, end = 1; end <= n; end++)
      if (end == n || (mp[start].code & MASK) != (mp[end].code & MASK)) {
        treelets.push_back(emit_lbvh(mp.data() + start, end - start, 29 - 12));
        start = end;
      }
    return upper_sah(std::move(treelets));
  }
};

struct Builder {
  size_t max_prims;
  int split;  // pt_split_method
  std::vector<size_t> ordered;

  std::unique_ptr<BuildNode> leaf(const PrimInfo* pi, size_t n, const Bounds3& b) {
    auto nd = std::make_unique<BuildNode>();
    nd->first = ordered.size();
    nd->n = n;
    nd->bounds = b;
    for (size_t i = 0; i < n; i++) ordered.push_back(pi[i].number);
    return nd;
  }
  std::unique_ptr<BuildNode> interior(int axis, std::unique_ptr<BuildNode> a, std::unique_ptr<BuildNode> b) {
    auto nd = std::make_unique<BuildNode>();
    nd->bounds = bunion(a->bounds, b->bounds);
    nd->split_axis = axis;
    nd->c[0] = std::move(a);
    nd->c[1] = std::move(b);
    return nd;
  }
  static void sort_dim(PrimInfo* pi, size_t n, int dim) {
    std::stable_sort(pi, pi + n, [dim](const PrimInfo& a, const PrimInfo& b) { return a.centroid[dim] < b.centroid[dim]; });
  }
  std::unique_ptr<BuildNode> equal_counts(int dim, PrimInfo* pi, size_t n) {
    sort_dim(pi, n, dim);
    size_t mid = n / 2;
    auto c0 = build(pi, mid);
    auto c1 = build(pi + mid, n - mid);
    return interior(dim, std::move(c0), std::move(c1));
  }
  std::unique_ptr<BuildNode> middle(int dim, Float p_mid, PrimInfo* pi, size_t n) {
    sort_dim(pi, n, dim);
    size_t mid = 0;
    while (mid < n && !(p_mid <= pi[mid].centroid[dim])) mid++;
    if (mid == 0 || mid == n) return equal_counts(dim, pi, n);
    auto c0 = build(pi, mid);
    auto c1 = build(pi + mid, n - mid);
    return interior(dim, std::move(c0), std::move(c1));
  }
  std::unique_ptr<BuildNode> sah(int dim, PrimInfo* pi, size_t n) {
    if (n == 1) return leaf(pi, n, pi[0].bounds);
    if (n == 2) {
      sort_dim(pi, n, dim);
      auto c0 = build(pi, 1);
      auto c1 = build(pi + 1, 1);
      return interior(dim, std::move(c0), std::move(c1));
    }
    Bounds3 bounds = pi[0].bounds;
    for (size_t i = 0; i < n; i++) bounds = bunion(bounds, pi[i].bounds);
    Bounds3 cb(pi[0].centroid, pi[0].centroid);
    for (size_t i = 0; i < n; i++) cb = bunion_p(cb, pi[i].centroid);
    const int NB = 12;
    struct Bucket { int count = 0; Bounds3 bounds; } buckets[NB];
    auto bucket_of = [&](const PrimInfo& p) {
      int b = f2i(floorf((Float)NB * cb.offset(p.centroid)[dim]));
      b = std::min(b, NB - 1);
      b = std::max(b, 0);
      return b;
    };
    for (size_t i = 0; i < n; i++) {
      int b = bucket_of(pi[i]);
      buckets[b].count += 1;
      buckets[b].bounds = bunion(buckets[b].bounds, pi[i].bounds);
    }
    Float cost[NB - 1];
    for (int i = 0; i < NB - 1; i++) {
      int count0 = 0, count1 = 0;
      Bounds3 b0 = buckets[i].bounds, b1 = buckets[i + 1].bounds;
      for (int j = 0; j <= i; j++) { b0 = bunion(b0, buckets[j].bounds); count0 += buckets[j].count; }
      for (int j = i + 1; j < NB; j++) { b1 = bunion(b1, buckets[j].bounds); count1 += buckets[j].count; }
      cost[i] = 1.0f + ((Float)count0 * b0.surface_area() + (Float)count1 * b1.surface_area()) / bounds.surface_area();
    }
    Float min_cost = cost[0];
    int min_bucket = 0;
    for (int i = 1; i < NB - 1; i++)
      if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
    Float leaf_cost = (Float)n;
    if (n > max_prims || min_cost < leaf_cost) {
      std::vector<PrimInfo> left, right;
      for (size_t i = 0; i < n; i++) (bucket_of(pi[i]) <= min_bucket ? left : right).push_back(pi[i]);
      if (left.empty() || right.empty()) return equal_counts(dim, pi, n);
      auto c0 = build(left.data(), left.size());
      auto c1 = build(right.data(), right.size());
      return interior(dim, std::move(c0), std::move(c1));
    }
    return leaf(pi, n, bounds);
  }
  std::unique_ptr<BuildNode> build(PrimInfo* pi, size_t n) {  // recursive_build
    Bounds3 bounds = pi[0].bounds;
    for (size_t i = 1; i < n; i++) bounds = bunion(bounds, pi[i].bounds);
    if (n <= max_prims) return leaf(pi, n, bounds);
    Bounds3 cb(pi[0].centroid, pi[0].centroid);
    for (size_t i = 1; i < n; i++) cb = bunion_p(cb, pi[i].centroid);
    int dim = cb.maximum_extent();
    if (cb.mn[dim] == cb.mx[dim]) return leaf(pi, n, bounds);
    if (split == PT_SPLIT_MIDDLE) return middle(dim, (cb.mn[dim] + cb.mx[dim]) / 2.0f, pi, n);
    if (split == PT_SPLIT_EQUAL) return equal_counts(dim, pi, n);
    return sah(dim, pi, n);
  }
};

// ---- QBVH (qbvh_x86.rs) -----------------------------------------------------------------------------
struct QNode {
  float bboxes[2][3][4];
  int64_t children[4];  // -1 = EMPTY_MASK
  uint8_t axis_top, axis_left, axis_right, is_leaf;
};
static const uint32_t ORDER_TABLE[128] = {
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

inline size_t flatten_qbvh(std::vector<QNode>& nodes, const BuildNode* node) {
  size_t offset = nodes.size();
  nodes.push_back(QNode());
  memset(&nodes[offset], 0, sizeof(QNode));
  if (node->n > 0) {
    nodes[offset].is_leaf = 1;
    nodes[offset].children[0] = (int64_t)node->first;
    nodes[offset].children[1] = (int64_t)node->n;
    return offset;
  }
  int64_t idx[4] = {0, 0, 0, 0};
  Vec3 boxes[4][2];
  const BuildNode* c0 = node->c[0].get();
  const BuildNode* c1 = node->c[1].get();
  if (c0->n > 0) {
    idx[0] = (int64_t)flatten_qbvh(nodes, c0);
    idx[1] = -1;
    boxes[0][0] = c0->bounds.mn; boxes[0][1] = c0->bounds.mx;
  } else {
    idx[0] = (int64_t)flatten_qbvh(nodes, c0->c[0].get());
    idx[1] = (int64_t)flatten_qbvh(nodes, c0->c[1].get());
    boxes[0][0] = c0->c[0]->bounds.mn; boxes[0][1] = c0->c[0]->bounds.mx;
    boxes[1][0] = c0->c[1]->bounds.mn; boxes[1][1] = c0->c[1]->bounds.mx;
  }
  if (c1->n > 0) {
    idx[2] = (int64_t)flatten_qbvh(nodes, c1);
    idx[3] = -1;
    boxes[2][0] = c1->bounds.mn; boxes[2][1] = c1->bounds.mx;
  } else {
    idx[2] = (int64_t)flatten_qbvh(nodes, c1->c[0].get());
    idx[3] = (int64_t)flatten_qbvh(nodes, c1->c[1].get());
    boxes[2][0] = c1->c[0]->bounds.mn; boxes[2][1] = c1->c[0]->bounds.mx;
    boxes[3][0] = c1->c[1]->bounds.mn; boxes[3][1] = c1->c[1]->bounds.mx;
  }
  QNode& q = nodes[offset];
  for (int j = 0; j < 3; j++)
    for (int k = 0; k < 4; k++) {
      q.bboxes[0][j][k] = boxes[k][0][j];
      q.bboxes[1][j][k] = boxes[k][1][j];
    }
  for (int k = 0; k < 4; k++) q.children[k] = idx[k];
  q.axis_top = (uint8_t)node->split_axis;
  q.axis_left = (uint8_t)c0->split_axis;
  q.axis_right = (uint8_t)c1->split_axis;
  return offset;
}

// SSE semantics: _mm_max_ps(a,b) = a > b ? a : b ; _mm_min_ps(a,b) = a < b ? a : b (NaN -> b)
inline float sse_max(float a, float b) { return a > b ? a : b; }
inline float sse_min(float a, float b) { return a < b ? a : b; }

// core/geometry/intersect.rs:46-65 with Float::max/min (NaN-ignoring) semantics
inline bool box_range(const Bounds3& b, const Ray& r, Float* t0o, Float* t1o) {
  Float t0 = 0.0f, t1 = r.t_max;
  for (int i = 0; i < 3; i++) {
    Float idir = 1.0f / r.d[i];
    int sign = std::signbit(r.d[i]) ? 1 : 0;
    Float lo = sign ? b.mx[i] : b.mn[i], hi = sign ? b.mn[i] : b.mx[i];
    t0 = rmax(t0, (lo - r.o[i]) * idir);
    t1 = rmin(t1, (hi - r.o[i]) * idir);
  }
  *t0o = t0; *t1o = t1;
  return t0 <= t1;
}

struct Accel;
struct SceneGeom {
  const pt_scene_desc* d;
  std::vector<std::unique_ptr<Accel>> objects;  // object k's accelerator
};

struct Hit {
  int tri = -1;       // global triangle index
  int instance = -1;  // instance index or -1
  Float t = 0, b0 = 0, b1 = 0, b2 = 0;
  Ray local_ray;      // the ray in the triangle's space (object space for instances)
};

// LinearBVHNode (accel/lbvh.rs:14-19)
struct LNode {
  Bounds3 bounds;
  size_t offset = 0;        // leaf: first primitive; interior: second child
  size_t n_primitives = 0;  // 0 -> interior node
  uint8_t axis = 0;
};
// flatten_bvh_tree (accel/lbvh.rs:55-73)
inline size_t flatten_lbvh(std::vector<LNode>& nodes, const BuildNode* node) {
  LNode ln;
  ln.bounds = node->bounds;
  ln.n_primitives = node->n;
  ln.axis = (uint8_t)node->split_axis;
  size_t offset = nodes.size();
  if (node->n > 0) {
    ln.offset = node->first;
    nodes.push_back(ln);
  } else {
    ln.n_primitives = 0;
    nodes.push_back(ln);
    flatten_lbvh(nodes, node->c[0].get());
    size_t second = flatten_lbvh(nodes, node->c[1].get());
    nodes[offset].offset = second;
  }
  return offset;
}
// intersect_box_array_i (core/geometry/intersect.rs:46-65)
inline bool box_array_i(const Bounds3& b, const Float* org, const Float* idir, const int* sign, Float* t0, Float* t1) {
  for (int i = 0; i < 3; i++) {
    Float lo = sign[i] ? b.mx[i] : b.mn[i], hi = sign[i] ? b.mn[i] : b.mx[i];
    *t0 = rmax(*t0, (lo - org[i]) * idir[i]);
    *t1 = rmin(*t1, (hi - org[i]) * idir[i]);
  }
  return *t0 <= *t1;
}

struct Accel {
  std::vector<int32_t> prims;  // ordered primitive list (triangle index >= 0, ~instance < 0)
  std::vector<QNode> nodes;
  std::vector<LNode> lnodes;   // Accelerator layout "lbvh" (LBVHAccel) instead of the 4-wide nodes
  bool lbvh = false;
  bool intersect_lbvh(const SceneGeom& g, const Ray& r, Hit* hit, Counters* cnt) const;
  bool intersect_p_lbvh(const SceneGeom& g, const Ray& r, Counters* cnt) const;
  Bounds3 bounds;
  bool single = false;  // object with exactly one primitive: no accelerator is built (scene_context.rs:1372-1379)
  size_t max_depth = 0;

  // closest hit: mutates r.t_max (Cell semantics); returns whether something was hit
  bool intersect(const SceneGeom& g, const Ray& r, Hit* hit, Counters* cnt) const;
  bool intersect_p(const SceneGeom& g, const Ray& r, Counters* cnt) const;
  bool prim_intersect(const SceneGeom& g, int32_t prim, const Ray& r, Hit* hit, Counters* cnt) const;
  bool prim_intersect_p(const SceneGeom& g, int32_t prim, const Ray& r, Counters* cnt) const;
};

inline Mat4 mat_of(const float* m) { Mat4 r; memcpy(r.m, m, sizeof(r.m)); return r; }

inline bool Accel::prim_intersect(const SceneGeom& g, int32_t prim, const Ray& r, Hit* hit, Counters* cnt) const {
  if (prim >= 0) {  // GeometricPrimitive::intersect
    if (cnt) cnt->tris++;
    TriHit th;
    if (!tri_intersect(get_tri(g.d, prim), r, &th)) return false;
    r.t_max = th.t;
    hit->tri = prim; hit->instance = -1; hit->t = th.t; hit->b0 = th.b0; hit->b1 = th.b1; hit->b2 = th.b2;
    hit->local_ray = r;
    return true;
  }
  // TransformedPrimitive::intersect (transformed_primitive.rs:26-39)
  int inst = ~prim;
  const pt_instance& I = g.d->instances[inst];
  Ray ray = transform_ray(mat_of(I.minv), r);
  Hit h;
  if (!g.objects[(size_t)I.object]->intersect(g, ray, &h, cnt)) return false;
  r.t_max = ray.t_max;
  *hit = h;
  hit->instance = inst;
  return true;
}
inline bool Accel::prim_intersect_p(const SceneGeom& g, int32_t prim, const Ray& r, Counters* cnt) const {
  if (prim >= 0) {
    if (cnt) cnt->tris++;
    TriHit th;
    return tri_intersect(get_tri(g.d, prim), r, &th);
  }
  const pt_instance& I = g.d->instances[~prim];
  Ray ray = transform_ray(mat_of(I.minv), r);
  return g.objects[(size_t)I.object]->intersect_p(g, ray, cnt);
}

// LBVHAccel::intersect (accel/lbvh.rs:128-186)
inline bool Accel::intersect_lbvh(const SceneGeom& g, const Ray& r, Hit* hit, Counters* cnt) const {
  if (lnodes.empty()) return false;
  bool found = false;
  struct E { size_t node; Float t0, t1; };
  std::vector<E> stack;
  stack.reserve(64);
  Float t_max = r.t_max;
  Float org[3] = {r.o.x, r.o.y, r.o.z};
  Float idir[3] = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  int neg[3] = {std::signbit(idir[0]) ? 1 : 0, std::signbit(idir[1]) ? 1 : 0, std::signbit(idir[2]) ? 1 : 0};
  stack.push_back(E{0, 0.0f, t_max});
  while (!stack.empty()) {
    E e = stack.back();
    stack.pop_back();
    Float t0 = e.t0, t1 = rmin(t_max, e.t1);
    if (t1 < t0) continue;
    const LNode& node = lnodes[e.node];
    if (cnt) cnt->nodes++;
    if (!box_array_i(node.bounds, org, idir, neg, &t0, &t1)) continue;
    if (node.n_primitives > 0) {
      for (size_t i = node.offset; i < node.offset + node.n_primitives; i++)
        if (prim_intersect(g, prims[i], r, hit, cnt)) { t_max = r.t_max; found = true; }
    } else {
      size_t child[2] = {e.node + 1, node.offset};
      int idx[2] = {neg[node.axis], 1 - neg[node.axis]};
      stack.push_back(E{child[idx[1]], t0, t1});
      stack.push_back(E{child[idx[0]], t0, t1});
    }
  }
  return found;
}
// LBVHAccel::intersect_p (accel/lbvh.rs:188-241): the popped range is not clipped against t_max
inline bool Accel::intersect_p_lbvh(const SceneGeom& g, const Ray& r, Counters* cnt) const {
  if (lnodes.empty()) return false;
  struct E { size_t node; Float t0, t1; };
  std::vector<E> stack;
  stack.reserve(64);
  Float org[3] = {r.o.x, r.o.y, r.o.z};
  Float idir[3] = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  int neg[3] = {std::signbit(idir[0]) ? 1 : 0, std::signbit(idir[1]) ? 1 : 0, std::signbit(idir[2]) ? 1 : 0};
  stack.push_back(E{0, 0.0f, r.t_max});
  while (!stack.empty()) {
    E e = stack.back();
    stack.pop_back();
    Float t0 = e.t0, t1 = e.t1;
    const LNode& node = lnodes[e.node];
    if (cnt) cnt->nodes++;
    if (!box_array_i(node.bounds, org, idir, neg, &t0, &t1)) continue;
    if (node.n_primitives > 0) {
      for (size_t i = node.offset; i < node.offset + node.n_primitives; i++)
        if (prim_intersect_p(g, prims[i], r, cnt)) return true;
    } else {
      size_t child[2] = {e.node + 1, node.offset};
      int idx[2] = {neg[node.axis], 1 - neg[node.axis]};
      stack.push_back(E{child[idx[1]], t0, t1});
      stack.push_back(E{child[idx[0]], t0, t1});
    }
  }
  return false;
}

inline bool Accel::intersect(const SceneGeom& g, const Ray& r, Hit* hit, Counters* cnt) const {
  if (single) return prim_intersect(g, prims[0], r, hit, cnt);
  if (lbvh) return intersect_lbvh(g, r, hit, cnt);
  Float tmin, tmax;
  if (!box_range(bounds, r, &tmin, &tmax)) return false;
  bool found = false;
  std::vector<size_t> stack;
  stack.reserve(64);
  Float org[3] = {r.o.x, r.o.y, r.o.z};
  Float idir[3] = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  int sign[3] = {std::signbit(r.d.x) ? 1 : 0, std::signbit(r.d.y) ? 1 : 0, std::signbit(r.d.z) ? 1 : 0};
  stack.push_back(0);
  while (!stack.empty()) {
    size_t cur = stack.back();
    stack.pop_back();
    const QNode& node = nodes[cur];
    if (cnt) cnt->nodes++;
    if (!node.is_leaf) {
      uint32_t mask = 0;
      for (int k = 0; k < 4; k++) {  // test_aabb (qbvh_x86.rs:26-67)
        Float tn = tmin, tx = tmax;
        for (int a = 0; a < 3; a++) {
          tn = sse_max(tn, (node.bboxes[sign[a]][a][k] - org[a]) * idir[a]);
          tx = sse_min(tx, (node.bboxes[1 - sign[a]][a][k] - org[a]) * idir[a]);
        }
        if (tx >= tn) mask |= 1u << k;
      }
      if (mask != 0) {
        uint32_t node_idx = (uint32_t)((sign[node.axis_top] << 2) | (sign[node.axis_left] << 1) | sign[node.axis_right]);
        uint32_t order = ORDER_TABLE[mask * 8 + node_idx];
        while ((order & 0x4) == 0) {
          int64_t c = node.children[order & 0x3];
          if (c != -1) stack.push_back((size_t)c);
          order >>= 4;
        }
      }
    } else {
      size_t start = (size_t)node.children[0], end = start + (size_t)node.children[1];
      bool any = false;
      for (size_t i = start; i < end; i++)
        if (prim_intersect(g, prims[i], r, hit, cnt)) any = true;
      if (any) { tmax = r.t_max; found = true; }
    }
  }
  return found;
}
inline bool Accel::intersect_p(const SceneGeom& g, const Ray& r, Counters* cnt) const {
  if (single) return prim_intersect_p(g, prims[0], r, cnt);
  if (lbvh) return intersect_p_lbvh(g, r, cnt);
  Float tmin, tmax;
  if (!box_range(bounds, r, &tmin, &tmax)) return false;
  std::vector<size_t> stack;
  stack.reserve(64);
  Float org[3] = {r.o.x, r.o.y, r.o.z};
  Float idir[3] = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  int sign[3] = {std::signbit(r.d.x) ? 1 : 0, std::signbit(r.d.y) ? 1 : 0, std::signbit(r.d.z) ? 1 : 0};
  stack.push_back(0);
  while (!stack.empty()) {
    size_t cur = stack.back();
    stack.pop_back();
    const QNode& node = nodes[cur];
    if (cnt) cnt->nodes++;
    if (!node.is_leaf) {
      uint32_t mask = 0;
      for (int k = 0; k < 4; k++) {
        Float tn = tmin, tx = tmax;
        for (int a = 0; a < 3; a++) {
          tn = sse_max(tn, (node.bboxes[sign[a]][a][k] - org[a]) * idir[a]);
          tx = sse_min(tx, (node.bboxes[1 - sign[a]][a][k] - org[a]) * idir[a]);
        }
        if (tx >= tn) mask |= 1u << k;
      }
      if (mask != 0) {
        uint32_t node_idx = (uint32_t)((sign[node.axis_top] << 2) | (sign[node.axis_left] << 1) | sign[node.axis_right]);
        uint32_t order = ORDER_TABLE[mask * 8 + node_idx];
        while ((order & 0x4) == 0) {
          int64_t c = node.children[order & 0x3];
          if (c != -1) stack.push_back((size_t)c);
          order >>= 4;
        }
      }
    } else {
      size_t start = (size_t)node.children[0], end = start + (size_t)node.children[1];
      for (size_t i = start; i < end; i++)
        if (prim_intersect_p(g, prims[i], r, cnt)) return true;
    }
  }
  return false;
}

// Triangle::world_bound (triangle.rs:214-224, union3 uses f32::min/max)
inline Bounds3 tri_bounds(const TriRef& T) {
  Vec3 mn = T.p0, mx = T.p0;
  const Vec3 a[2] = {T.p1, T.p2};
  for (int j = 0; j < 2; j++) {
    mn = Vec3(rmin(mn.x, a[j].x), rmin(mn.y, a[j].y), rmin(mn.z, a[j].z));
    mx = Vec3(rmax(mx.x, a[j].x), rmax(mx.y, a[j].y), rmax(mx.z, a[j].z));
  }
  return Bounds3(mn, mx);
}
// Transform::transform_bounds (transform.rs:134-181)
inline Bounds3 transform_bounds(const Mat4& M, const Bounds3& b) {
  Vec3 mn, mx;
  bool first = true;
  for (int ix = 0; ix < 2; ix++)
    for (int iy = 0; iy < 2; iy++)
      for (int iz = 0; iz < 2; iz++) {
        Vec3 q = M.point(Vec3(ix ? b.mx.x : b.mn.x, iy ? b.mx.y : b.mn.y, iz ? b.mx.z : b.mn.z));
        if (first) { mn = q; mx = q; first = false; }
        else {
          mn = Vec3(rmin(mn.x, q.x), rmin(mn.y, q.y), rmin(mn.z, q.z));
          mx = Vec3(rmax(mx.x, q.x), rmax(mx.y, q.y), rmax(mx.z, q.z));
        }
      }
  return Bounds3(mn, mx);
}

inline size_t build_depth(const BuildNode* n) {
  if (n->n > 0) return 1;
  return 1 + std::max(build_depth(n->c[0].get()), build_depth(n->c[1].get()));
}

inline Bounds3 prim_world_bound(const SceneGeom& g, int32_t prim);

inline std::unique_ptr<Accel> build_accel(const SceneGeom& g, const int32_t* prims, size_t n, const pt_accel& cfg, bool top) {
  auto A = std::make_unique<Accel>();
  if (n == 0) { A->single = false; A->bounds = Bounds3(); return A; }
  if (!top && n == 1) {
    A->single = true;
    A->prims.assign(prims, prims + 1);
    A->bounds = prim_world_bound(g, prims[0]);
    return A;
  }
  const Float BOUND_EPS = std::numeric_limits<float>::epsilon() * 2.0f;
  std::vector<PrimInfo> info(n);
  for (size_t i = 0; i < n; i++) {
    Bounds3 b = prim_world_bound(g, prims[i]);
    Vec3 mn(b.mn.x - BOUND_EPS, b.mn.y - BOUND_EPS, b.mn.z - BOUND_EPS);
    Vec3 mx(b.mx.x + BOUND_EPS, b.mx.y + BOUND_EPS, b.mx.z + BOUND_EPS);
    Bounds3 bi(mn, mx);
    info[i].number = i;
    info[i].bounds = bi;
    info[i].centroid = (bi.mn + bi.mx) * 0.5f;
  }
  Builder B;
  B.max_prims = (size_t)std::min(std::max(cfg.max_prims_in_node, 0), 255);
  B.split = cfg.split_method;
  std::unique_ptr<BuildNode> root;
  if (cfg.split_method == PT_SPLIT_HLBVH) {
    HlbvhBuilder H{B.max_prims, info.data(), &B.ordered};
    root = H.build(n);
  } else root = B.build(info.data(), n);
  A->prims.resize(n);
  for (size_t i = 0; i < n; i++) A->prims[i] = prims[B.ordered[i]];
  if (cfg.type == PT_ACCEL_LBVH) { A->lbvh = true; flatten_lbvh(A->lnodes, root.get()); }
  else flatten_qbvh(A->nodes, root.get());
  A->bounds = root->bounds;
  A->max_depth = build_depth(root.get());
  return A;
}

// AnimatedTransform::motion_bounds for a static transform = transform_bounds (animated_transform.rs:154-157)
inline Bounds3 prim_world_bound(const SceneGeom& g, int32_t prim) {
  if (prim >= 0) return tri_bounds(get_tri(g.d, prim));
  const pt_instance& I = g.d->instances[~prim];
  return transform_bounds(mat_of(I.m), g.objects[(size_t)I.object]->bounds);
}

}  // namespace orc
