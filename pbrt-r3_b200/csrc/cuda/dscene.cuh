// dscene.cuh — HBM layout of the flattened scene (what pt_scene_create uploads) and of the wavefront state.
//
// QBVH node, 128 B, 16-byte aligned (reference SIMDBVHNode is 144 B: accel/qbvh/qbvh_x86.rs:15-24):
//   f4 lo[3], hi[3]  : bboxes[0][axis][lane], bboxes[1][axis][lane]                          96 B
//   i4 child         : lane k -> >= 0 inner/leaf NODE index | < 0 inline leaf ~((first<<3)|(count-1)) | PT_EMPTY
//   i4 meta          : x = axis_top | axis_left<<2 | axis_right<<4 | is_leaf<<8 ; y = first ; z = count (leaf nodes)
// Triangle record, 48 B, stored in BVH leaf order (one coalesced run per leaf):
//   f4 v0 = p0.xyz, bits(id)   id = global triangle index, or instance index when flag PT_REC_INSTANCE
//   f4 v1 = p1.xyz, bits(flags)
//   f4 v2 = p2.xyz, 0
#pragma once
#include "../../../include/pbrt_r3_b200.h"
#include "dmath.cuh"

namespace ptd {

#define PT_EMPTY ((int)0x80000000)
#define PT_LEAF_BITS 3
#define PT_REC_INSTANCE 0x100u
#define PT_LEAF_REMAINDER 0x40000000u  /* inline-leaf stack entries of two-level scenes: first record < 2^27 */
#define PT_MAX_STACK 96

struct alignas(16) QNode {
  f4 lo[3];
  f4 hi[3];
  i4 child;
  i4 meta;
};
static_assert(sizeof(QNode) == 128, "QNode must be 128 B");

struct ObjectRec {     // one per object (0 = top level)
  uint32_t node_base;  // into nodes
  uint32_t rec_base;   // into tri records (in units of records)
  uint32_t n_nodes;
  uint32_t n_prims;
  float bmin[3], bmax[3];  // QBVHAccel.bounds (root build-node bounds)
  int32_t single;          // 1: object with a single primitive, no accelerator, no bounds pre-test
  int32_t pad;
};

struct InstanceRec {
  float m[16];
  float minv[16];
  int32_t object;
  int32_t pad[3];
};

struct alignas(16) PrimeRec {  // PRIMES[dim], PRIME_SUMS[dim] (lowdiscrepancy/primes.rs) + ceil(2^64 / prime) for exact division by multiply-high
  uint32_t prime;
  uint32_t perm_offset;
  uint64_t magic;
};

struct SamplerDev {     // samplers/{halton,sobol}.rs state that is constant over the render
  int32_t type, spp;    // spp = effective (sobol: rounded up to a power of two)
  int32_t sb[4];
  int32_t at_center;
  int32_t base_scales[2], base_exponents[2], sample_stride, mult_inverse[2];
  uint32_t resolution, log2_resolution;
  // Sampler "stratified" (a PixelSampler, core/sampler/pixel_sampler.rs): SampleStream::dim holds the 1D requests so far in its
  // low half and the 2D requests in its high half; the first `nd` of each read the pixel's tables, later ones the tile's RNG.
  // inc2d / x2d / y2d: what get2d adds to dim (2, 0, 1 for the index-based samplers: kernel-parameter constants either way)
  int32_t nd, strata, jitter;
  uint32_t inc2d, x2d, y2d;
};

struct LightDistDev {   // dense restatement of SpatialLightDistribution's hash table (a pure function of the voxel)
  int32_t strategy;     // pt_light_strategy actually used
  uint32_t voxels[3];
  float wb_min[3], wb_max[3];
  const float* func;    // [n_vox][n_lights]
  const float* cdf;     // [n_vox][n_lights + 1]
  const float* func_int;  // [n_vox]
};

// The lobes of a material whose parameters are all constants, listed by the host (uber, substrate, translucent): what
// Material::compute_scattering_functions would add, in its order (dbsdf.cuh Lobe fields, structure of arrays)
struct MatLobes {
  int32_t n;              // -1: no BSDF at all (the ray passes through)
  float eta;
  int32_t kind[5], sp0[5], sp1[5], fres[5], scale[5];  // scale: MixMaterial's factor on the lobe (spectrum index, -1 = none)
  float a[5], b[5], ei[5], et[5];
};

struct EnvDev {           // one per LightSource "infinite" (lights/infinite.rs): level 0 of the map + the Distribution2D tables
  float l2w[9], w2l[9];   // upper 3x3 of light_to_world / world_to_light
  int32_t w, h;           // level 0 (powers of two); every lookup on the path has width 0 = the bilinear filter of level 0
  int32_t nu, nv;         // Distribution2D resolution (2w x 2h)
  const f4* texels;       // [h][w] rgb-
  const float* func;      // [nv][nu]      conditional_v[v].func
  const float* cdf;       // [nv][nu + 1]  conditional_v[v].cdf
  const float* row_int;   // [nv]          conditional_v[v].func_int = marginal.func
  const float* mcdf;      // [nv + 1]      marginal.cdf
  float m_int;            // marginal.func_int
  float hitcher;          // projection light: near plane
  int32_t pad[2];
  float xf[16];           // projection light: light_projection | goniometric light: world_to_light (4x4, row-major)
  float screen[4];        // projection light: screen bounds min.x, min.y, max.x, max.y
};

struct DScene {
  uint32_t n_channels;
  uint32_t n_tris, n_lights, n_materials, n_objects, n_instances;
  const QNode* nodes;        // 4-wide layout (accel_type PT_ACCEL_QBVH)
  const f4* lnodes;          // binary LinearBVHNode layout, 2 x f4 per node (accel_type PT_ACCEL_LBVH, dlbvh.cuh)
  int accel_type;
  const f4* recs;            // 3 per record
  // PT_TRI_PERM: the records once more with p.xyz stored as (y,z,x) [copy 0] and (z,x,y) [copy 1]: copy kz holds the components
  // in the order (kx, ky, kz) the watertight test permutes them to for a ray whose dominant axis is kz (triangle.rs:258-268);
  // copy 2 is `recs` itself.  recs_perm + kz * rec_stride = copy kz (rec_stride in f4 units); null when the switch is off.
  const f4* recs_perm;
  uint32_t rec_stride;
  const ObjectRec* objects;
  const InstanceRec* instances;
  // original-order triangle attributes (indexed by global triangle id)
  const f4* tri_v;           // 3 per triangle: p.xyz, w unused (shading / light sampling re-reads by id)
  const float* tri_n;        // 9 per triangle or null
  const float* tri_uv;       // 6 per triangle or null
  const float* tri_s;        // 9 per triangle or null: mesh tangents "S"
  const uint32_t* tri_flags;
  const int32_t* tri_material;
  const int32_t* tri_light;
  const pt_material* materials;
  const MatLobes* mat_lobes;  // per material (meaningful for uber / substrate / translucent)
  const pt_light* lights;
  const uint32_t* light_sample_prefix;  // [n_lights + 1] prefix sums of pt_light.n_samples (directlighting "all": round -> light, sample)
  const pt_spot* spots;      // SpotLight extras, indexed by pt_light.tri
  const EnvDev* envs;        // InfiniteAreaLight extras, indexed by pt_light.tri
  int32_t has_map_lights;    // projection / goniometric lights: their code lives in the ENV variants of the kernels
  const int32_t* inf_lights; // Scene.infinite_lights: indices into lights
  uint32_t n_inf;
  const float* spectra;
  const float* alpha;        // per material: roughness_to_alpha applied on the HOST with libm logf: [n_materials][2]
  float xyz_w[180];
  float xyz_scale;
  float world_radius;
  pt_camera camera;
  pt_film film;
  pt_integrator integrator;
  SamplerDev sampler;
  LightDistDev ld;
  // sampler tables
  const PrimeRec* prime_recs;  // 1024: one 16-byte record per Halton dimension
  const uint16_t* perms;       // sum(primes)
  const uint32_t* sobol32;     // 1024*52
  const uint64_t* vdc;         // (25+26)*52
  int has_instances;
  // Sampler "random": PCG32 state of every 16x16 render tile of the current shard (samplers/random.rs; one sequential
  // stream per tile, sampler.rs:285); rewritten by every draw
  uint64_t* tile_rng;
  // its 2D sample arrays (ambientocclusion, directlighting "all"): RandomSampler::start_pixel draws every element of every
  // requested array before the pixel's first sample (random.rs:46-60); tile_arr_stride floats per tile, refilled per pixel
  float* tile_arrays;
  uint32_t tile_arr_stride;
  // Sampler "stratified": the tile buffer starts with the pixel's tables ([nd][spp] 1D values, [nd][spp] 2D pairs; tile_tab
  // floats), the sample arrays follow; tile_sample = sample number of the tile's path in flight
  uint32_t tile_tab;
  uint32_t* tile_sample;
};

}  // namespace ptd
