// scene_accel.cpp — see scene_accel.h.
#include "scene_accel.h"

#include <stdexcept>

#include <cfloat>
#include <cmath>
#include <cstring>

namespace pth {
namespace {

// Triangle::world_bound (shapes/triangle.rs:188-224, union3 with f32::min/max then Bounds3f::from)
BBox tri_bbox(const float* p) {
  BBox b;
  for (int a = 0; a < 3; a++) {
    float mn = p[a], mx = p[a];
    mn = fminf(mn, p[3 + a]); mx = fmaxf(mx, p[3 + a]);
    mn = fminf(mn, p[6 + a]); mx = fmaxf(mx, p[6 + a]);
    b.mn[a] = mn <= mx ? mn : mx;
    b.mx[a] = mn >= mx ? mn : mx;
  }
  return b;
}
// Transform::transform_bounds (core/transform/transform.rs:134-181); static instances: motion_bounds == this
BBox xform_bbox(const float* m, const BBox& b) {
  BBox r;
  bool first = true;
  for (int ix = 0; ix < 2; ix++)
    for (int iy = 0; iy < 2; iy++)
      for (int iz = 0; iz < 2; iz++) {
        float x = ix ? b.mx[0] : b.mn[0], y = iy ? b.mx[1] : b.mn[1], z = iz ? b.mx[2] : b.mn[2];
        float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
        float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
        float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
        float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
        float v[3] = {xp, yp, zp};
        if (wp != 1.0f) { v[0] = xp / wp; v[1] = yp / wp; v[2] = zp / wp; }
        for (int a = 0; a < 3; a++) {
          if (first) { r.mn[a] = v[a]; r.mx[a] = v[a]; }
          else { r.mn[a] = fminf(r.mn[a], v[a]); r.mx[a] = fmaxf(r.mx[a], v[a]); }
        }
        first = false;
      }
  return r;
}

}  // namespace

BBox instance_world_bounds(const float* instance_to_world, const BBox& b) { return xform_bbox(instance_to_world, b); }

std::string build_objects(const pt_scene_desc* D, std::vector<ObjectHost>* out) {
  out->clear();
  out->resize(D->n_objects);
  // instanced objects first (they are referenced by the top level), then object 0
  for (uint32_t k = D->n_objects; k-- > 0;) {
    ObjectHost& O = (*out)[k];
    const int32_t* pr = D->prims + D->obj_first[k];
    size_t n = D->obj_first[k + 1] - D->obj_first[k];
    std::vector<BBox> pb(n);
    for (size_t i = 0; i < n; i++) {
      if (pr[i] >= 0) {
        if ((uint32_t)pr[i] >= D->n_tris) return "primitive index out of range";
        pb[i] = tri_bbox(D->tri_p + (size_t)pr[i] * 9);
      } else {
        uint32_t ii = (uint32_t)~pr[i];
        if (ii >= D->n_instances) return "instance index out of range";
        const pt_instance& I = D->instances[ii];
        if (k != 0 || I.object <= 0 || (uint32_t)I.object >= D->n_objects) return "bad instance object index";
        pb[i] = xform_bbox(I.m, (*out)[(size_t)I.object].bounds);
      }
    }
    if (n == 0) {
      for (int a = 0; a < 3; a++) { O.bounds.mn[a] = FLT_MAX; O.bounds.mx[a] = -FLT_MAX; }
    } else if (k != 0 && n == 1) {
      O.single = true;
      O.bounds = pb[0];
      O.ordered.assign(pr, pr + 1);
    } else {
      try { build_bvh(pb, D->accel.max_prims_in_node, D->accel.split_method, &O.bvh); }
      catch (const std::exception& e) { return std::string("BVH build: ") + e.what(); }
      O.bounds = O.bvh.bounds;
      O.ordered.resize(n);
      for (size_t i = 0; i < n; i++) O.ordered[i] = pr[O.bvh.ordered[i]];
    }
  }
  return "";
}

}  // namespace pth

// ---- host-only build-parity entry points -------------------------------------------------------------------------
void pt_set_error(const std::string& s);

struct pt_host_bvh {
  std::vector<pth::ObjectHost> objects;
  std::vector<std::vector<pth::RefNodeH>> ref;
};

extern "C" {

int pt_host_bvh_build(const pt_scene_desc* desc, pt_host_bvh** out) {
  if (!desc || !out) { pt_set_error("pt_host_bvh_build: null argument"); return 1; }
  pt_host_bvh* h = new pt_host_bvh();
  std::string e = pth::build_objects(desc, &h->objects);
  if (!e.empty()) { delete h; pt_set_error(e); return 1; }
  h->ref.resize(h->objects.size());
  for (size_t k = 0; k < h->objects.size(); k++) pth::flatten_reference(h->objects[k].bvh, &h->ref[k]);
  *out = h;
  return 0;
}
int pt_host_bvh_info(const pt_host_bvh* h, uint32_t object, uint32_t* n_nodes, uint32_t* n_prims, int32_t* ordered, float* bounds6,
                     uint32_t* gpu_nodes, uint32_t* qdepth) {
  if (!h || object >= h->objects.size()) { pt_set_error("pt_host_bvh_info: bad object"); return 1; }
  const pth::ObjectHost& O = h->objects[object];
  if (n_nodes) *n_nodes = (uint32_t)h->ref[object].size();
  if (n_prims) *n_prims = (uint32_t)O.ordered.size();
  if (ordered) for (size_t i = 0; i < O.ordered.size(); i++) ordered[i] = O.ordered[i];
  if (bounds6) for (int a = 0; a < 3; a++) { bounds6[a] = O.bounds.mn[a]; bounds6[3 + a] = O.bounds.mx[a]; }
  if (gpu_nodes) *gpu_nodes = (uint32_t)O.bvh.nodes.size();
  if (qdepth) *qdepth = O.bvh.qdepth;
  return 0;
}
int pt_host_bvh_node(const pt_host_bvh* h, uint32_t object, uint32_t node, float* boxes24, int64_t* children4, int32_t* axes3, int32_t* is_leaf) {
  if (!h || object >= h->objects.size() || node >= h->ref[object].size()) { pt_set_error("pt_host_bvh_node: bad index"); return 1; }
  const pth::RefNodeH& q = h->ref[object][node];
  memcpy(boxes24, q.bboxes, sizeof(q.bboxes));
  for (int k = 0; k < 4; k++) children4[k] = q.children[k];
  for (int k = 0; k < 3; k++) axes3[k] = q.axes[k];
  *is_leaf = q.is_leaf;
  return 0;
}
void pt_host_bvh_destroy(pt_host_bvh* h) { delete h; }

}  // extern "C"
