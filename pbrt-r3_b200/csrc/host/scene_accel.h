// scene_accel.h — builds the accelerators of every object of a scene description on the host
// (SceneContext::make_scene -> create_accelerator("bvh") -> QBVHAccel::new, scene_context.rs:718-731,
// qbvh_x86.rs:352-371; ObjectInstance accelerators, scene_context.rs:1349-1391).
#pragma once
#include <string>
#include <vector>

#include "../../../include/pbrt_r3_b200.h"
#include "bvh_build.h"

namespace pth {

struct ObjectHost {
  BvhResult bvh;
  std::vector<int32_t> ordered;  // primitive ids (triangle >= 0, ~instance < 0) in leaf order
  bool single = false;           // object with one primitive: no accelerator is built (scene_context.rs:1372-1379)
  BBox bounds;
};

// TransformedPrimitive::world_bound of an instance of an object with bounds `b` (Transform::transform_bounds)
BBox instance_world_bounds(const float* instance_to_world, const BBox& b);

// returns "" on success, otherwise an error message
std::string build_objects(const pt_scene_desc* D, std::vector<ObjectHost>* out);

}  // namespace pth
