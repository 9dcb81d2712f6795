// bvh_build.h — host-side SAH BVH construction and flattening to the GPU wide-node layout.
// Replaces reference src/accelerators/bvh/build/{node,sah,equal_counts,middle,types}.rs (binary build) and
// src/accelerators/bvh/accel/qbvh/qbvh_x86.rs:93-176 (two-level collapse into 4-wide nodes).
// The tree must be IDENTICAL to the reference's (leaf order decides equal-t ties), so the build keeps its exact
// rules: leaf as soon as n <= maxnodeprims, 12-bucket SAH in f32, order-preserving partition, stable sorts,
// bounds inflated by 2*f32::EPSILON.  Unlike the reference it partitions in place (no Vec copies per split).
#pragma once
#include <cstdint>
#include <vector>

namespace pth {

struct BBox {
  float mn[3], mx[3];
};

struct BuildNodeH {
  BBox bounds;
  int32_t child[2];   // -1 for leaves
  int32_t axis;
  uint32_t first, n;  // leaf: range in the ordered primitive list
};

struct GpuNodeH {     // 128 B, mirrors ptd::QNode
  float lo[3][4];
  float hi[3][4];
  int32_t child[4];
  int32_t meta[4];
};
static_assert(sizeof(GpuNodeH) == 128, "GpuNodeH must be 128 bytes");

struct LNodeH {       // 32 B, the LinearBVHNode layout of cuda/dlbvh.cuh (accel/lbvh.rs:14-19)
  float mn[3];
  int32_t offset;     // leaf: first primitive of the ordered list; interior: second child (the first one is this + 1)
  float mx[3];
  uint32_t meta;      // n_primitives | axis << 30
};
static_assert(sizeof(LNodeH) == 32, "LNodeH must be 32 bytes");

struct RefNodeH {     // the reference's SIMDBVHNode view (for build-parity checks only)
  float bboxes[2][3][4];
  int64_t children[4];
  int32_t axes[3];
  int32_t is_leaf;
};

struct BvhResult {
  std::vector<BuildNodeH> tree;    // binary tree, root = 0
  std::vector<uint32_t> ordered;   // ordered[i] = index into the input primitive list
  std::vector<GpuNodeH> nodes;     // GPU layout, root = 0
  BBox bounds;
  uint32_t qdepth = 0;             // depth of the 4-wide tree (for the traversal stack bound)
  uint32_t n_ref_nodes = 0;        // node count of the reference's own flattening (leaf nodes included)
};

// prim_bounds: world bounds of each primitive BEFORE inflation.
void build_bvh(const std::vector<BBox>& prim_bounds, int max_prims_in_node, int split_method, BvhResult* out);
// flatten_bvh_tree (accel/lbvh.rs:55-73): depth-first, root = 0; returns the depth of the binary tree
uint32_t flatten_lbvh(const BvhResult& r, std::vector<LNodeH>* out);
// the reference's flatten_qbvh_tree, for parity checks
void flatten_reference(const BvhResult& r, std::vector<RefNodeH>* out);

}  // namespace pth
