// bvh_build.cpp — see bvh_build.h.  Reference: src/accelerators/bvh/build/node.rs:27-151, build/sah.rs:31-170,
// build/equal_counts.rs:6-30, build/middle.rs:10-44, accel/qbvh/qbvh_x86.rs:93-176.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace pth {
namespace {

struct Info {
  uint32_t number;
  BBox b;
  float c[3];
};

// Bounds3f helpers with the reference's PartialOrd min_/max_ (bounds3.rs:38-44)
inline float mn_(float a, float b) { return a <= b ? a : b; }
inline float mx_(float a, float b) { return a >= b ? a : b; }
inline BBox b_union(const BBox& a, const BBox& b) {
  BBox r;
  for (int i = 0; i < 3; i++) { r.mn[i] = mn_(a.mn[i], b.mn[i]); r.mx[i] = mx_(a.mx[i], b.mx[i]); }
  return r;
}
inline BBox b_empty() {  // Bounds3f::default (bounds3.rs:224-231)
  BBox r;
  for (int i = 0; i < 3; i++) { r.mn[i] = FLT_MAX; r.mx[i] = -FLT_MAX; }
  return r;
}
inline float b_area(const BBox& b) {
  float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  return 2.0f * (dx * dy + dx * dz + dy * dz);
}
inline int b_max_extent(const float mn[3], const float mx[3]) {
  float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  if (dx > dy && dx > dz) return 0;
  if (dy > dz) return 1;
  return 2;
}

// The build is deterministic AND parallel: a subtree over n primitives owns the slice [ord, ord + n) of the ordered
// primitive list (the reference appends leaves in depth-first order, so the left subtree's primitives come first) and
// takes its nodes from a pre-sized arena through an atomic counter, so the two children of a large node can be built by
// different threads.  Node numbers then depend on the thread schedule, but nothing reads them: the tree is only ever
// walked through its child links (flattening, parity views), and the root is node 0.
struct Ctx {
  BvhResult* out;
  size_t max_prims;
  int split;
  std::atomic<int32_t> next_node{0};
  std::atomic<int> spare_threads{0};
  static constexpr size_t kParallelMin = 1u << 14;  // primitives below which a subtree is not worth a thread

  int32_t reserve() { return next_node.fetch_add(1); }
  int32_t leaf_at(int32_t self, const Info* pi, size_t n, const BBox& b, size_t ord) {
    if (self < 0) self = reserve();
    BuildNodeH& nd = out->tree[(size_t)self];
    nd.bounds = b;
    nd.child[0] = nd.child[1] = -1;
    nd.axis = 0;
    nd.first = (uint32_t)ord;
    nd.n = (uint32_t)n;
    for (size_t i = 0; i < n; i++) out->ordered[ord + i] = pi[i].number;
    return self;
  }
  void set_interior(int32_t id, int axis, int32_t c0, int32_t c1) {
    BuildNodeH& nd = out->tree[(size_t)id];
    nd.bounds = b_union(out->tree[(size_t)c0].bounds, out->tree[(size_t)c1].bounds);
    nd.child[0] = c0;
    nd.child[1] = c1;
    nd.axis = axis;
    nd.first = 0;
    nd.n = 0;
  }
  // the two children of `self`: [pi, pi + nl) then [pi + nl, pi + n); the larger halves of big nodes on their own threads
  int32_t children(int32_t self, int dim, Info* pi, size_t nl, size_t n, size_t ord) {
    int32_t c0, c1;
    bool forked = false;
    if (nl >= kParallelMin && n - nl >= kParallelMin) {
      int avail = spare_threads.load();
      while (avail > 0 && !spare_threads.compare_exchange_weak(avail, avail - 1)) {}
      forked = avail > 0;
    }
    if (forked) {
      std::thread t([&]() { c0 = build(pi, nl, -1, ord); });
      c1 = build(pi + nl, n - nl, -1, ord + nl);
      t.join();
      spare_threads.fetch_add(1);
    } else {
      c0 = build(pi, nl, -1, ord);
      c1 = build(pi + nl, n - nl, -1, ord + nl);
    }
    set_interior(self, dim, c0, c1);
    return self;
  }
  static void sort_dim(Info* pi, size_t n, int dim) {
    std::stable_sort(pi, pi + n, [dim](const Info& a, const Info& b) { return a.c[dim] < b.c[dim]; });
  }
  int32_t equal_counts(int dim, Info* pi, size_t n, int32_t self, size_t ord) {
    sort_dim(pi, n, dim);
    return children(self, dim, pi, n / 2, n, ord);
  }

  // ---- splitmethod "hlbvh" (build/hlbvh.rs:332-428): Morton codes of the centroids (10 bits per axis), one LBVH treelet
  // per run of equal top-12 bits, SAH over the treelet roots.  Sequential: leaves are appended to the ordered list in the
  // order the reference creates them.
  struct MortonPrim { uint32_t index, code; };
  const Info* hl_info = nullptr;
  size_t hl_ord = 0;
  static uint32_t left_shift3(uint32_t x) {  // :23-37
    if (x >= 1024u) x = 1023u;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
  }
  static uint32_t f2u(float v) { return v != v ? 0u : (v <= 0.0f ? 0u : (v >= 4294967296.0f ? 0xffffffffu : (uint32_t)v)); }  // `as u32`
  int32_t hl_leaf(const MortonPrim* mp, size_t n) {
    const int32_t self = reserve();
    BuildNodeH& nd = out->tree[(size_t)self];
    BBox b = hl_info[mp[0].index].b;
    for (size_t i = 1; i < n; i++) b = b_union(b, hl_info[mp[i].index].b);
    nd.bounds = b;
    nd.child[0] = nd.child[1] = -1;
    nd.axis = 0;
    nd.first = (uint32_t)hl_ord;
    nd.n = (uint32_t)n;
    for (size_t i = 0; i < n; i++) out->ordered[hl_ord + i] = mp[i].index;
    hl_ord += n;
    return self;
  }
  int32_t hl_interior(int axis, int32_t c0, int32_t c1) {
    const int32_t self = reserve();
    set_interior(self, axis, c0, c1);
    return self;
  }
  int32_t hl_split_node(MortonPrim* mp, size_t n, int dim) {  // :86-137
    if (n <= max_prims) return hl_leaf(mp, n);
    const Info* info = hl_info;
    std::stable_sort(mp, mp + n, [info, dim](const MortonPrim& a, const MortonPrim& b) { return info[a.index].c[dim] < info[b.index].c[dim]; });
    const int next_dim = (dim + 3 - 1) % 3;
    const size_t split = n / 2;
    const int32_t n0 = hl_split_node(mp, split, next_dim);
    const int32_t n1 = hl_split_node(mp + split, n - split, next_dim);
    return hl_interior(dim, n0, n1);
  }
  int32_t hl_emit(const MortonPrim* mp, size_t n, int bit_index) {  // emit_lbvh (:139-226)
    if (n <= max_prims) return hl_leaf(mp, n);
    if (bit_index == -1) {
      std::vector<MortonPrim> copy(mp, mp + n);
      return hl_split_node(copy.data(), n, 2);
    }
    const uint32_t mask = 1u << bit_index;
    if ((mp[0].code & mask) == (mp[n - 1].code & mask)) return hl_emit(mp, n, bit_index - 1);
    size_t s0 = 0, s1 = n - 1;
    while (s0 + 1 != s1) {
      const size_t mid = (s0 + s1) / 2;
      if ((mp[s0].code & mask) == (mp[mid].code & mask)) s0 = mid; else s1 = mid;
    }
    const int32_t n0 = hl_emit(mp, s1, bit_index - 1);
    const int32_t n1 = hl_emit(mp + s1, n - s1, bit_index - 1);
    return hl_interior(bit_index % 3, n0, n1);
  }
  int32_t hl_upper_sah(const std::vector<int32_t>& roots) {  // build_upper_sah (:234-330), with the reference's cost formula as written
    if (roots.empty()) throw std::runtime_error("splitmethod \"hlbvh\": an empty side in the SAH over the treelets (the reference panics here)");
    if (roots.size() <= 1) return roots[0];
    auto B = [&](int32_t id) -> const BBox& { return out->tree[(size_t)id].bounds; };
    BBox bounds = B(roots[0]);
    for (size_t i = 1; i < roots.size(); i++) bounds = b_union(bounds, B(roots[i]));
    float cmn[3], cmx[3];
    for (int a = 0; a < 3; a++) cmn[a] = cmx[a] = (B(roots[0]).mn[a] + B(roots[0]).mx[a]) * 0.5f;
    for (size_t i = 1; i < roots.size(); i++)
      for (int a = 0; a < 3; a++) { float c = (B(roots[i]).mn[a] + B(roots[i]).mx[a]) * 0.5f; cmn[a] = mn_(cmn[a], c); cmx[a] = mx_(cmx[a], c); }
    const int dim = b_max_extent(cmn, cmx);
    const int NB = 12;
    uint32_t count[NB];
    BBox bb[NB];
    for (int i = 0; i < NB; i++) { count[i] = 0; bb[i] = b_empty(); }
    auto bucket = [&](int32_t id) {
      const float c = (B(id).mn[dim] + B(id).mx[dim]) * 0.5f;
      const uint32_t b = f2u((float)NB * ((c - cmn[dim]) / (cmx[dim] - cmn[dim])));
      return (int)std::min<uint32_t>(b, (uint32_t)NB - 1);
    };
    for (int32_t r : roots) {
      const int b = bucket(r);
      count[b]++;
      bb[b] = count[b] == 1 ? B(r) : b_union(bb[b], B(r));
    }
    float cost[NB - 1];
    for (int i = 0; i < NB - 1; i++) {
      BBox b0 = bb[i], b1 = bb[i];
      uint32_t c0 = 0, c1 = 1;
      for (int j = 0; j < i; j++) { b0 = b_union(b0, bb[j]); c0 += count[j]; }
      for (int j = i + 1; j < NB; j++) { b1 = b_union(b1, bb[j]); c1 += count[j]; }
      cost[i] = 0.125f + (((float)c0 * b_area(b0) + (float)c1 * b_area(b1)) / b_area(bounds));
    }
    float min_cost = cost[0];
    int min_b = 0;
    for (int i = 1; i < NB - 1; i++)
      if (cost[i] < min_cost) { min_cost = cost[i]; min_b = i; }
    std::vector<int32_t> c0, c1;
    for (int32_t r : roots) (bucket(r) <= min_b ? c0 : c1).push_back(r);
    const int32_t n0 = hl_upper_sah(c0);
    const int32_t n1 = hl_upper_sah(c1);
    return hl_interior(dim, n0, n1);
  }
  int32_t hlbvh(const Info* info, size_t n) {
    hl_info = info;
    hl_ord = 0;
    BBox bounds = info[0].b;
    for (size_t i = 1; i < n; i++) bounds = b_union(bounds, info[i].b);
    std::vector<MortonPrim> mp(n);
    auto clamp01 = [](float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); };
    for (size_t i = 0; i < n; i++) {
      mp[i].index = info[i].number;
      float o[3];
      for (int a = 0; a < 3; a++) {  // Bounds3f::offset
        o[a] = info[i].c[a] - bounds.mn[a];
        if (bounds.mx[a] > bounds.mn[a]) o[a] = o[a] / (bounds.mx[a] - bounds.mn[a]);
      }
      mp[i].code = (left_shift3(f2u(ceilf(clamp01(o[2]) * 1024.0f))) << 2) | (left_shift3(f2u(ceilf(clamp01(o[1]) * 1024.0f))) << 1) |
                   left_shift3(f2u(ceilf(clamp01(o[0]) * 1024.0f)));
    }
    std::stable_sort(mp.begin(), mp.end(), [](const MortonPrim& a, const MortonPrim& b) { return a.code < b.code; });  // the LSD radix sort is stable
    std::vector<int32_t> treelets;
    const uint32_t MASK = 0x3ffc0000u;
    for (size_t start = 0, end = 1; end <= n; end++)
      if (end == n || (mp[start].code & MASK) != (mp[end].code & MASK)) {
        treelets.push_back(hl_emit(mp.data() + start, end - start, 29 - 12));
        start = end;
      }
    return hl_upper_sah(treelets);
  }

  int32_t build(Info* pi, size_t n, int32_t self, size_t ord) {  // recursive_build (node.rs:27-105)
    BBox bounds = pi[0].b;
    for (size_t i = 1; i < n; i++) bounds = b_union(bounds, pi[i].b);
    if (n <= max_prims) return leaf_at(self, pi, n, bounds, ord);
    float cmn[3] = {pi[0].c[0], pi[0].c[1], pi[0].c[2]}, cmx[3] = {pi[0].c[0], pi[0].c[1], pi[0].c[2]};
    for (size_t i = 1; i < n; i++)
      for (int a = 0; a < 3; a++) { cmn[a] = mn_(cmn[a], pi[i].c[a]); cmx[a] = mx_(cmx[a], pi[i].c[a]); }
    int dim = b_max_extent(cmn, cmx);
    if (cmn[dim] == cmx[dim]) return leaf_at(self, pi, n, bounds, ord);
    if (self < 0) self = reserve();
    if (split == 2) {  // middle.rs
      float p_mid = (cmn[dim] + cmx[dim]) / 2.0f;
      sort_dim(pi, n, dim);
      size_t mid = 0;
      while (mid < n && !(p_mid <= pi[mid].c[dim])) mid++;
      if (mid == 0 || mid == n) return equal_counts(dim, pi, n, self, ord);
      return children(self, dim, pi, mid, n, ord);
    }
    if (split == 3) return equal_counts(dim, pi, n, self, ord);
    // split_sah (sah.rs:31-170); n >= 3 here unless maxnodeprims < 2
    if (n == 2) {
      sort_dim(pi, n, dim);
      return children(self, dim, pi, 1, n, ord);
    }
    const int NB = 12;
    int count[NB];
    BBox bb[NB];
    for (int i = 0; i < NB; i++) { count[i] = 0; bb[i] = b_empty(); }
    const float ext = cmx[dim] - cmn[dim];
    auto bucket_of = [&](const Info& p) {
      float o = p.c[dim] - cmn[dim];
      if (cmx[dim] > cmn[dim]) o = o / ext;
      float f = floorf((float)NB * o);
      int b = (f != f) ? 0 : (f >= 2147483648.0f ? INT32_MAX : (f <= -2147483648.0f ? INT32_MIN : (int)f));
      if (b > NB - 1) b = NB - 1;
      if (b < 0) b = 0;
      return b;
    };
    for (size_t i = 0; i < n; i++) {
      int b = bucket_of(pi[i]);
      count[b]++;
      bb[b] = b_union(bb[b], pi[i].b);
    }
    float cost[NB - 1];
    const float inv_parent = b_area(bounds);
    for (int i = 0; i < NB - 1; i++) {
      int c0 = 0, c1 = 0;
      BBox b0 = bb[i], b1 = bb[i + 1];
      for (int j = 0; j <= i; j++) { b0 = b_union(b0, bb[j]); c0 += count[j]; }
      for (int j = i + 1; j < NB; j++) { b1 = b_union(b1, bb[j]); c1 += count[j]; }
      cost[i] = 1.0f + ((float)c0 * b_area(b0) + (float)c1 * b_area(b1)) / inv_parent;
    }
    float min_cost = cost[0];
    int min_b = 0;
    for (int i = 1; i < NB - 1; i++)
      if (cost[i] < min_cost) { min_cost = cost[i]; min_b = i; }
    float leaf_cost = (float)n;
    if (n > max_prims || min_cost < leaf_cost) {
      // order-preserving partition (Iterator::partition), in place via a scratch copy of the right side
      size_t nl = 0;
      static thread_local std::vector<Info> scratch;  // reused: the copy is over before the children are built
      scratch.clear();
      for (size_t i = 0; i < n; i++) {
        if (bucket_of(pi[i]) <= min_b) pi[nl++] = pi[i];
        else scratch.push_back(pi[i]);
      }
      if (nl == 0 || nl == n) {
        // nothing moved (one side empty): the slice is still in its original order
        if (nl == 0) std::copy(scratch.begin(), scratch.end(), pi);
        return equal_counts(dim, pi, n, self, ord);
      }
      std::copy(scratch.begin(), scratch.end(), pi + nl);
      return children(self, dim, pi, nl, n, ord);
    }
    return leaf_at(self, pi, n, bounds, ord);
  }
};

const int32_t kEmpty = (int32_t)0x80000000;
const uint32_t kLeafBits = 3;

struct Flat {
  const BvhResult* r;
  std::vector<GpuNodeH>* nodes;
  uint32_t depth = 0;

  // reference to a build node from its parent: inline leaf when it fits, otherwise a node index
  int32_t ref(int32_t b, uint32_t d) {
    const BuildNodeH& n = r->tree[(size_t)b];
    if (n.n > 0 && n.n <= (1u << kLeafBits) && n.first < (1u << (31 - kLeafBits - 1))) {
      if (d + 1 > depth) depth = d + 1;
      return ~(int32_t)((n.first << kLeafBits) | (n.n - 1));
    }
    return emit(b, d + 1);
  }
  int32_t emit(int32_t b, uint32_t d) {
    if (d > depth) depth = d;
    size_t off = nodes->size();
    nodes->emplace_back();
    memset(&(*nodes)[off], 0, sizeof(GpuNodeH));
    const BuildNodeH& n = r->tree[(size_t)b];
    if (n.n > 0) {  // leaf node kept in the reference's form (root leaf, or more than 8 primitives)
      GpuNodeH& g = (*nodes)[off];
      g.meta[0] = 0x100;
      g.meta[1] = (int32_t)n.first;
      g.meta[2] = (int32_t)n.n;
      for (int k = 0; k < 4; k++) g.child[k] = kEmpty;
      return (int32_t)off;
    }
    int32_t idx[4] = {kEmpty, kEmpty, kEmpty, kEmpty};
    BBox box[4];
    memset(box, 0, sizeof(box));  // empty lanes keep a zero box (qbvh_x86.rs:69-74,101)
    const BuildNodeH& c0 = r->tree[(size_t)n.child[0]];
    const BuildNodeH& c1 = r->tree[(size_t)n.child[1]];
    int32_t kids[4] = {-1, -1, -1, -1};
    if (c0.n > 0) kids[0] = n.child[0]; else { kids[0] = c0.child[0]; kids[1] = c0.child[1]; }
    if (c1.n > 0) kids[2] = n.child[1]; else { kids[2] = c1.child[0]; kids[3] = c1.child[1]; }
    for (int k = 0; k < 4; k++)
      if (kids[k] >= 0) {
        box[k] = r->tree[(size_t)kids[k]].bounds;
        idx[k] = ref(kids[k], d);
      }
    GpuNodeH& g = (*nodes)[off];
    // The reference keeps a zero box in an empty lane and skips the lane by its child index (qbvh_x86.rs:69-74,267-274);
    // here the lane gets the box (+inf, -inf), which no slab test can hit, so the traversal needs no index check.
    for (int a = 0; a < 3; a++)
      for (int k = 0; k < 4; k++) {
        g.lo[a][k] = kids[k] >= 0 ? box[k].mn[a] : INFINITY;
        g.hi[a][k] = kids[k] >= 0 ? box[k].mx[a] : -INFINITY;
      }
    for (int k = 0; k < 4; k++) g.child[k] = idx[k];
    g.meta[0] = (n.axis & 3) | ((c0.axis & 3) << 2) | ((c1.axis & 3) << 4);
    // meta.w: for each of the 8 ray sign octants sg (bit a = sign of d[a]) the ORDER_TABLE column
    // sign[axis_top] << 2 | sign[axis_left] << 1 | sign[axis_right] (qbvh_x86.rs:264-266), 3 bits each
    uint32_t tab = 0;
    for (uint32_t sg = 0; sg < 8; sg++) {
      uint32_t st = (sg >> (n.axis & 3)) & 1u, sl = (sg >> (c0.axis & 3)) & 1u, sr = (sg >> (c1.axis & 3)) & 1u;
      tab |= ((st << 2) | (sl << 1) | sr) << (3 * sg);
    }
    g.meta[3] = (int32_t)tab;
    return (int32_t)off;
  }
};

size_t ref_flatten(const BvhResult& r, int32_t b, std::vector<RefNodeH>* out) {
  size_t off = out->size();
  out->emplace_back();
  memset(&(*out)[off], 0, sizeof(RefNodeH));
  const BuildNodeH& n = r.tree[(size_t)b];
  if (n.n > 0) {
    (*out)[off].is_leaf = 1;
    (*out)[off].children[0] = n.first;
    (*out)[off].children[1] = n.n;
    return off;
  }
  const BuildNodeH& c0 = r.tree[(size_t)n.child[0]];
  const BuildNodeH& c1 = r.tree[(size_t)n.child[1]];
  int32_t kids[4] = {-1, -1, -1, -1};
  if (c0.n > 0) kids[0] = n.child[0]; else { kids[0] = c0.child[0]; kids[1] = c0.child[1]; }
  if (c1.n > 0) kids[2] = n.child[1]; else { kids[2] = c1.child[0]; kids[3] = c1.child[1]; }
  int64_t idx[4];
  for (int k = 0; k < 4; k++) idx[k] = kids[k] >= 0 ? (int64_t)ref_flatten(r, kids[k], out) : -1;
  RefNodeH& q = (*out)[off];
  for (int k = 0; k < 4; k++) {
    q.children[k] = idx[k];
    if (kids[k] >= 0) {
      const BBox& bx = r.tree[(size_t)kids[k]].bounds;
      for (int a = 0; a < 3; a++) { q.bboxes[0][a][k] = bx.mn[a]; q.bboxes[1][a][k] = bx.mx[a]; }
    }
  }
  q.axes[0] = n.axis; q.axes[1] = c0.axis; q.axes[2] = c1.axis;
  return off;
}

}  // namespace

void build_bvh(const std::vector<BBox>& prim_bounds, int max_prims_in_node, int split_method, BvhResult* out) {
  out->tree.clear();
  out->ordered.clear();
  out->nodes.clear();
  size_t n = prim_bounds.size();
  if (n == 0) { out->bounds = b_empty(); return; }
  const float BOUND_EPS = FLT_EPSILON * 2.0f;  // node.rs:13
  std::vector<Info> info(n);
  for (size_t i = 0; i < n; i++) {
    Info& f = info[i];
    f.number = (uint32_t)i;
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = prim_bounds[i].mn[a] - BOUND_EPS; mx[a] = prim_bounds[i].mx[a] + BOUND_EPS; }
    for (int a = 0; a < 3; a++) { f.b.mn[a] = mn_(mn[a], mx[a]); f.b.mx[a] = mx_(mn[a], mx[a]); }
    for (int a = 0; a < 3; a++) f.c[a] = (f.b.mn[a] + f.b.mx[a]) * 0.5f;
  }
  Ctx c;
  c.out = out;
  c.max_prims = (size_t)std::min(std::max(max_prims_in_node, 0), 255);
  c.split = split_method;
  out->tree.assign(2 * n, BuildNodeH());   // a binary tree over n primitives has at most 2n - 1 nodes
  out->ordered.assign(n, 0u);
  {
    unsigned hw = std::thread::hardware_concurrency();
    c.spare_threads.store((int)std::min(std::max(hw, 1u), 64u) - 1);
  }
  if (split_method == 1) {  // PT_SPLIT_HLBVH: bottom-up, the root is made last; it is copied into node 0 (nothing reads node numbers)
    const int32_t slot0 = c.reserve();
    const int32_t root = c.hlbvh(info.data(), n);
    out->tree[(size_t)slot0] = out->tree[(size_t)root];
  } else {
    int32_t root = c.build(info.data(), n, c.reserve(), 0);
    (void)root;
  }
  out->tree.resize((size_t)c.next_node.load());
  out->bounds = out->tree[0].bounds;
  Flat f;
  f.r = out;
  f.nodes = &out->nodes;
  f.emit(0, 1);
  out->qdepth = f.depth;
  // node count of the reference's flattening = one per reachable build node after the 2-level collapse
  std::vector<RefNodeH> tmp;
  out->n_ref_nodes = 0;
}

namespace {
uint32_t lbvh_emit(const BvhResult& r, int32_t b, size_t base, std::vector<LNodeH>* out) {
  const size_t off = out->size();
  out->emplace_back();
  const BuildNodeH& n = r.tree[(size_t)b];
  LNodeH g;
  for (int a = 0; a < 3; a++) { g.mn[a] = n.bounds.mn[a]; g.mx[a] = n.bounds.mx[a]; }
  uint32_t depth = 1;
  if (n.n > 0) {
    g.offset = (int32_t)n.first;
    g.meta = n.n & 0x3fffffffu;
    (*out)[off] = g;
    return depth;
  }
  g.meta = ((uint32_t)n.axis & 3u) << 30;
  uint32_t d0 = lbvh_emit(r, n.child[0], base, out);
  g.offset = (int32_t)(out->size() - base);  // node indices are relative to the object's first node
  uint32_t d1 = lbvh_emit(r, n.child[1], base, out);
  (*out)[off] = g;
  return 1 + std::max(d0, d1);
}
}  // namespace

uint32_t flatten_lbvh(const BvhResult& r, std::vector<LNodeH>* out) {
  if (r.tree.empty()) return 0;
  return lbvh_emit(r, 0, out->size(), out);
}

void flatten_reference(const BvhResult& r, std::vector<RefNodeH>* out) {
  out->clear();
  if (!r.tree.empty()) ref_flatten(r, 0, out);
}

}  // namespace pth
