/*
 * pbrt_r3_b200.h — C ABI of the B200-native PathIntegrator core.
 *
 * Drop-in boundary for ONE hot path of ototoi/pbrt-r3: the PathIntegrator
 * radiance bounce loop.  The reference has no FFI of its own;
 * the path sits behind
 *     trait Integrator { fn render(&mut self, scene: &Scene); ... }
 *         (reference src/core/integrator/integrator.rs:6-9)
 * selected by   create_integrator("path", ...)
 *         (reference src/integrators/create_integrator.rs:15-36)
 * and fed by the ParseContext callbacks of SceneContext
 *         (reference src/core/api/parse_context.rs:5-66).
 * A Rust `GpuPathContext: ParseContext` shim records the same
 * callbacks into the POD arrays below and calls these entry points.
 *
 * Conventions: plain pointers + sizes, no C++/torch types.  All input arrays
 * are owned by the caller and copied during pt_scene_create().  The callee
 * owns device memory.  Return 0 = ok, non-zero = error (pt_last_error()).
 * Float is f32 ("float-as-double" is off, reference src/core/base/types.rs:3-6).
 */
#ifndef PBRT_R3_B200_H
#define PBRT_R3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_ABI_VERSION 3

/* ---- plugin names carried as enums (string names are resolved by the host) ---- */
enum pt_material_type {       /* reference src/materials/create_material.rs:20-72 */
  PT_MAT_NONE = -1,           /* Material "" / "none": no BSDF, ray passes through */
  PT_MAT_MATTE = 0,           /* materials/matte.rs:26-52   */
  PT_MAT_PLASTIC = 1,         /* materials/plastic.rs:32-70 */
  PT_MAT_GLASS = 2,           /* materials/glass.rs:47-110  */
  PT_MAT_METAL = 3,           /* materials/metal.rs:52-84   */
  PT_MAT_MIRROR = 4,          /* materials/mirror.rs:20-41  */
  PT_MAT_UBER = 5,            /* materials/uber.rs:73-128: opacity pass-through + diffuse + glossy + specular reflection / transmission */
  PT_MAT_SUBSTRATE = 6,       /* materials/substrate.rs:33-60: FresnelBlend */
  PT_MAT_TRANSLUCENT = 7,     /* materials/translucent.rs:37-100: diffuse + glossy reflection and transmission, eta 1.5 */
  PT_MAT_MIX = 8              /* materials/mix.rs:60-93: the lobes of two other materials as ScaledBxDFs (kd, ks = their indices in
                                 materials, neither a mix itself; kr = spectrum index of "amount") */
};

enum pt_light_type {          /* reference src/lights/create_light.rs:13-54 */
  PT_LIGHT_AREA = 0,          /* lights/diffuse.rs  (one per emissive triangle) */
  PT_LIGHT_POINT = 1,         /* lights/point.rs    */
  PT_LIGHT_DISTANT = 2,       /* lights/distant.rs  */
  PT_LIGHT_SPOT = 3,          /* lights/spot.rs     */
  PT_LIGHT_INFINITE = 4,      /* lights/infinite.rs: environment light (constant L, or a lat-long map) */
  PT_LIGHT_PROJECTION = 5,    /* lights/projection.rs: a point light seen through a slide (perspective projection of an image) */
  PT_LIGHT_GONIOMETRIC = 6    /* lights/goniometric.rs: a point light scaled by a lat-long image of directions */
};

enum pt_sampler_type {        /* reference src/samplers/create_sampler.rs:12-33 */
  PT_SAMPLER_HALTON = 0,      /* samplers/halton.rs */
  PT_SAMPLER_SOBOL = 1,       /* samplers/sobol.rs  */
  PT_SAMPLER_RANDOM = 2,      /* samplers/random.rs */
  PT_SAMPLER_STRATIFIED = 3,  /* samplers/stratified.rs (xsamples = ysamples): per-pixel tables + the tile's RNG, like random */
  PT_SAMPLER_ZEROTWO = 4      /* samplers/zerotwosequence.rs ("02sequence" / "lowdiscrepancy"): scrambled (0,2)-sequence tables per pixel;
                                 samples_per_pixel as requested (rounded up to a power of two by the callee), n_sampled_dimensions */
};

enum pt_accel_type {          /* reference src/accelerators/bvh/accel/mod.rs:4-10 */
  PT_ACCEL_QBVH = 0,          /* "bvh" on x86_64/aarch64: accel/qbvh/qbvh_x86.rs */
  PT_ACCEL_LBVH = 1           /* "bvh" elsewhere:          accel/lbvh.rs          */
};

enum pt_split_method {        /* accelerators/bvh/create_bvh_accelerator.rs:13-20 */
  PT_SPLIT_SAH = 0, PT_SPLIT_HLBVH = 1 /* build/hlbvh.rs: Morton-code treelets + SAH over the treelets */, PT_SPLIT_MIDDLE = 2, PT_SPLIT_EQUAL = 3
};

enum pt_light_strategy {      /* core/lightdistrib/create_light_sample_distribution.rs:11-45 */
  PT_LIGHTS_SPATIAL = 0, PT_LIGHTS_UNIFORM = 1, PT_LIGHTS_POWER = 2
};

enum pt_spectrum_kind {       /* core/spectrum/mod.rs:16-20 */
  PT_SPECTRUM_RGB = 0,        /* RGBSpectrum, 3 channels       */
  PT_SPECTRUM_SAMPLED = 1     /* SampledSpectrum, 60 channels  */
};

/* triangle flag bits (per triangle; a mesh's flags are replicated on its triangles) */
#define PT_TRI_REVERSE_ORIENTATION 1u  /* TriangleMesh.reverse_orientation (shapes/triangle.rs:13) */
#define PT_TRI_SWAPS_HANDEDNESS    2u  /* TriangleMesh.swaps_handedness    (shapes/triangle.rs:14) */
#define PT_TRI_TWO_SIDED           4u  /* TriangleMesh.two_sided, pbrt-r3 only (shapes/triangle.rs:15) */
#define PT_TRI_HAS_N               8u  /* mesh has per-vertex normals  */
#define PT_TRI_HAS_UV             16u  /* mesh has per-vertex uv       */
#define PT_TRI_HAS_S              32u  /* mesh has per-vertex tangents "S" (shapes/triangle.rs:18,386-391) */

/* A material after (constant) texture evaluation.  Spectra are indices into
 * pt_scene_desc.spectra (n_channels floats each). */
typedef struct pt_material {
  int32_t type;               /* pt_material_type */
  int32_t kd;                 /* matte Kd | plastic Kd | uber / substrate / translucent Kd       (spectrum index) */
  int32_t ks;                 /* plastic Ks | glass Kt | metal k | uber / substrate / translucent Ks (spectrum index) */
  int32_t kr;                 /* mirror Kr | glass Kr | metal eta | uber Kr | translucent reflect  (spectrum index) */
  float sigma;                /* matte sigma (degrees)                                      */
  float uroughness;           /* plastic/metal/uber/translucent roughness (also u), glass/substrate uroughness */
  float vroughness;           /* glass/metal/uber/substrate vroughness                       */
  float eta;                  /* glass / uber index                                          */
  int32_t remap_roughness;    /* bool                                                        */
  int32_t kt;                 /* uber Kt | translucent transmit            (spectrum index, else -1) */
  int32_t opacity;            /* uber opacity                              (spectrum index, else -1) */
} pt_material;

/* A light.  Area lights reference a top-level triangle (its index in tri_*). */
typedef struct pt_light {
  int32_t type;               /* pt_light_type */
  int32_t spectrum;           /* area: Lemit = L*scale | point: I*scale | distant: L*scale (spectrum index) */
  int32_t tri;                /* area: triangle index; spot: index into pt_scene_desc.spots; infinite, projection, goniometric:
                                 index into pt_scene_desc.envmaps; else -1 */
  int32_t two_sided;          /* area: DiffuseAreaLight.two_sided (lights/diffuse.rs:178) */
  float v[3];                 /* point, spot, projection, goniometric: p_light (world) | distant: w_light (world, normalised) */
  int32_t n_samples;          /* Light::get_sample_count: area "nsamples" (default 1, n / 4 under --quick, diffuse.rs:177-185), else 1;
                                 read by directlighting's "all" strategy only */
} pt_light;

/* InfiniteAreaLight extras (lights/infinite.rs:12-95): level 0 of the MIPMap<RGBSpectrum> the light looks up.  The
 * callee builds the pyramid and the Distribution2D from it (make_distribution, infinite.rs:67-93). */
typedef struct pt_envmap {
  float l2w[9];               /* upper 3x3 of light_to_world, row-major */
  float w2l[9];               /* upper 3x3 of world_to_light */
  int32_t width, height;      /* as read from the file (the MIPMap resamples to powers of two itself); 1 x 1 for a light without "mapname" */
  const float* texels;        /* width * height RGB triples, row 0 = top: read_image(mapname) clamped at 0, times L * scale
                                 (make_mipmap, infinite.rs:47-65); projection / goniometric: clamped at 0 only, the
                                 intensity I * scale (L * scale) is pt_light.spectrum */
  float xf[16];               /* projection: light_projection = Transform::perspective(fov, 1e-3, 1e30) (projection.rs:44-46) |
                                 goniometric: the full world_to_light (its scale() transforms the direction as a POINT,
                                 goniometric.rs:33-34), row-major */
  float screen_bounds[4];     /* projection: min.x, min.y, max.x, max.y (projection.rs:33-41) */
  float hitcher;              /* projection: 1e-3 */
  float cos_total_width;      /* projection: z of the normalised corner direction (projection.rs:47-50), for power() */
} pt_envmap;

/* SpotLight extras (lights/spot.rs:5-56): falloff(w) looks at the z component of normalize(world_to_light * w). */
typedef struct pt_spot {
  float w2l[9];               /* upper 3x3 of world_to_light, row-major */
  float cos_total_width;      /* cos(radians(coneangle)), host libm like the reference */
  float cos_falloff_start;    /* cos(radians(coneangle - conedeltaangle)) */
  float pad;
} pt_spot;

/* An ObjectInstance: a TransformedPrimitive over object `object`
 * (core/primitive/transformed_primitive.rs:26-46).  m = instance-to-world, row-major m[4*r+c]. */
typedef struct pt_instance {
  int32_t object;             /* index into objects (>= 1; object 0 is the top level) */
  float m[16];
  float minv[16];
} pt_instance;

enum pt_camera_type {
  PT_CAMERA_PERSPECTIVE = 0,  /* cameras/perspective.rs */
  PT_CAMERA_ORTHOGRAPHIC = 1, /* cameras/orthographic.rs:60-90 */
  PT_CAMERA_ENVIRONMENT = 2   /* cameras/environment.rs:36-58: lat-long directions from film.full_resolution; no lens, raster_to_camera unused */
};

typedef struct pt_camera {    /* cameras/{perspective,orthographic,environment}.rs, core/camera/projective.rs */
  float raster_to_camera[16]; /* ProjectiveCamera.raster_to_camera.m   (row-major) */
  float camera_to_world[16];  /* BaseCamera.camera_to_world (static)   (row-major) */
  float lens_radius;
  float focal_distance;
  float shutter_open;
  float shutter_close;
  int32_t type;               /* pt_camera_type */
  int32_t pad;
} pt_camera;

typedef struct pt_film {      /* core/film/film.rs:64-164 */
  int32_t full_resolution[2];
  int32_t cropped_pixel_bounds[4]; /* x0,y0,x1,y1 */
  int32_t sample_bounds[4];        /* Film::get_sample_bounds (film.rs:166-179): x0,y0,x1,y1 */
  float filter_radius[2];
  float filter_table[256];         /* 16x16, film.rs:102-120 */
  float scale;
  float max_sample_luminance;
} pt_film;

typedef struct pt_sampler {
  int32_t type;               /* pt_sampler_type */
  int32_t samples_per_pixel;  /* as requested (sobol rounds up to a power of two itself); stratified: strata * strata */
  int32_t sample_at_pixel_center; /* halton only */
  int32_t strata;             /* stratified: xsamples = ysamples (samplers/stratified.rs:130-150) */
  int32_t jitter;             /* stratified: "jitter" */
  int32_t n_sampled_dimensions; /* stratified: "dimensions" = 1D and 2D tables per pixel; later requests draw from the RNG */
} pt_sampler;

enum pt_integrator_type {
  PT_INTEGRATOR_PATH = 0,     /* integrators/path.rs */
  PT_INTEGRATOR_AOV = 1,      /* integrators/aov.rs: one closest-hit query per camera sample, a geometric quantity as colour */
  PT_INTEGRATOR_AO = 2,       /* integrators/ao.rs ("ambientocclusion"): nsamples any-hit queries per camera sample, directions
                                 from the sampler's 2D sample ARRAY (request_2d_array / get_2d_array) */
  PT_INTEGRATOR_DIRECT = 3,   /* integrators/directlighting.rs: Le + direct lighting (all lights, or one picked uniformly) at every
                                 vertex of the specular reflection / transmission TREE (SamplerIntegrator::specular_reflect /
                                 specular_transmit, core/integrator/sampler.rs:37-160), depth < max_depth */
  PT_INTEGRATOR_WHITTED = 4   /* integrators/whitted.rs: one light sample per light (no MIS, no Le) on the same tree */
};
enum pt_direct_strategy {     /* directlighting.rs:7-10, "strategy" parameter (default "all") */
  PT_DIRECT_ALL = 0,          /* uniform_sample_all_lights: n_samples per light from the sampler's 2D sample arrays */
  PT_DIRECT_ONE = 1           /* uniform_sample_one_light without a light distribution */
};
enum pt_aov_target {          /* integrators/aov.rs:10-58; the targets that need ray differentials are not supported */
  PT_AOV_DISTANCE = 0,        /* "distance": t_max / |d|                       */
  PT_AOV_DEPTH = 1,           /* "depth":    |p - o|                           */
  PT_AOV_N = 2,               /* "n", "ng":  0.5 n + 0.5                       */
  PT_AOV_NS = 3,              /* "ns", "shading.n"                             */
  PT_AOV_UV = 4,              /* "uv": (u, v, 0)                               */
  PT_AOV_DPDUS = 5            /* "dpdus", "shading.dpdu"                       */
};
typedef struct pt_integrator { /* integrators/path.rs:252-271, integrators/aov.rs:192-213 */
  int32_t max_depth;
  float rr_threshold;
  int32_t light_strategy;     /* pt_light_strategy */
  int32_t type;               /* pt_integrator_type */
  int32_t aov_target;         /* pt_aov_target */
  float aov_scale;
  int32_t ao_samples;         /* ambientocclusion "nsamples" (default 64; 1 under --quick), ao.rs:124-131 */
  int32_t ao_cos_sample;      /* ambientocclusion "cossample" (default true) */
  int32_t direct_strategy;    /* enum pt_direct_strategy; directlighting only */
} pt_integrator;

typedef struct pt_accel {     /* accelerators/bvh/create_bvh_accelerator.rs:9-28 */
  int32_t type;               /* pt_accel_type */
  int32_t split_method;       /* pt_split_method */
  int32_t max_prims_in_node;  /* default 4 */
  int32_t pad;
} pt_accel;

/* Full flattened scene description (what SceneContext::make_scene/make_integrator hold). */
typedef struct pt_scene_desc {
  uint32_t abi_version;       /* PT_ABI_VERSION */
  uint32_t spectrum_kind;     /* pt_spectrum_kind */
  uint32_t n_channels;        /* 3 or 60 */
  uint32_t n_spectra;
  const float* spectra;       /* n_spectra * n_channels */
  float xyz_weights[3 * 60];  /* rows X,Y,Z over channels: Spectrum::to_xyz weights (rgb.rs:58-66 / sampled.rs:53-67) */
  float xyz_scale;            /* 1 for RGB; (700-400)/(CIE_Y_INTEGRAL*60) for sampled */

  /* Triangles of ALL objects, world space for object 0, object space of the
   * instanced object otherwise.  Degenerate triangles (area <= 1e-16) are
   * already dropped by the host (shapes/triangle.rs:726). */
  uint32_t n_tris;
  const float* tri_p;         /* 9 floats per triangle: p0 p1 p2                        */
  const float* tri_n;         /* 9 floats per triangle or NULL (PT_TRI_HAS_N)           */
  const float* tri_uv;        /* 6 floats per triangle or NULL (PT_TRI_HAS_UV)          */
  const float* tri_s;         /* 9 floats per triangle or NULL (PT_TRI_HAS_S): tangents, object_to_world applied */
  const uint32_t* tri_flags;  /* PT_TRI_* bits                                           */
  const int32_t* tri_material;/* index into materials, or -1 (no material)              */
  const int32_t* tri_light;   /* index into lights (area light of this triangle) or -1 */

  /* Objects: object k owns primitives prims[obj_first[k] .. obj_first[k+1]) in
   * definition order.  A primitive is a triangle index (>= 0) or ~instance
   * index (< 0).  Object 0 is the top-level aggregate (RenderOptions.primitives). */
  uint32_t n_objects;         /* >= 1 */
  const uint32_t* obj_first;  /* n_objects + 1 entries */
  const int32_t* prims;       /* obj_first[n_objects] entries */
  uint32_t n_instances;
  const pt_instance* instances;

  uint32_t n_materials;
  const pt_material* materials;
  uint32_t n_lights;          /* sampled lights, in Scene.lights order (RenderOptions.lights push order) */
  uint32_t n_emitters_only;   /* area lights of INSTANCED shapes: they emit when hit but are never sampled
                                 (scene_context.rs:1296-1299); stored at lights[n_lights ..] */
  const pt_light* lights;     /* n_lights + n_emitters_only records */
  uint32_t n_spots;
  const pt_spot* spots;
  uint32_t n_envmaps;
  const pt_envmap* envmaps;

  pt_camera camera;
  pt_film film;
  pt_sampler sampler;
  pt_integrator integrator;
  pt_accel accel;
} pt_scene_desc;

/* Render options that are not part of the scene (how the work is executed). */
typedef struct pt_render_opts {
  int32_t device;             /* CUDA device ordinal */
  int32_t shard_index;        /* this process's shard (0 .. shard_count-1) */
  int32_t shard_count;        /* tile-interleaved sharding over processes/GPUs (1 = whole image) */
  int32_t tile_size;          /* sharding granule in pixels (default 16 when 0) */
  int32_t spp_begin;          /* render samples [spp_begin, spp_end) of every pixel; 0,0 = all */
  int32_t spp_end;
  int32_t wave_paths;         /* path slots per wave (0 = default) */
  int32_t collect_counters;   /* 1: count nodes/triangles visited per ray class (slower) */
  int32_t wave_segments;      /* waves in flight: the bounce loop is software-pipelined over this many waves, each in its own
                                 segment of the path state, so that the thin late bounces of one wave share their kernel
                                 launches with the early bounces of the next (0 = default, 1 = one wave at a time, max 8).
                                 Results do not depend on it: the film is gathered wave by wave in order. */
} pt_render_opts;

typedef struct pt_stats {
  uint64_t camera_rays;       /* "Integrator/Camera rays traced"                  (sampler.rs:24) */
  uint64_t regular_rays;      /* "Intersections/Regular ray intersection tests"   (scene.rs:11)   */
  uint64_t shadow_rays;       /* "Intersections/Shadow ray intersection tests"    (scene.rs:12)   */
  uint64_t nodes_visited;     /* QBVH/LBVH nodes fetched (collect_counters)       */
  uint64_t tris_tested;       /* "Intersections/Ray-triangle intersection tests"  (triangle.rs:7) */
  uint64_t path_vertices;     /* shaded vertices */
  uint64_t zero_radiance_samples;
  uint64_t kernel_launches;   /* CUDA kernels launched by this call */
  double   device_ms;         /* CUDA-event time of the whole call on the device */
  double   extend_ms;         /* summed CUDA-event time of the closest-hit kernel (when timed) */
  double   shadow_ms;
  double   shade_ms;
  uint64_t extend_launches;
  uint64_t extend_rays;       /* rays processed by the closest-hit kernel */
  uint64_t extend_bytes;      /* algorithmic bytes of those rays: 128*nodes + 48*tris + 48 (collect_counters) */
} pt_stats;

typedef struct pt_scene pt_scene;

/* ---- life cycle ------------------------------------------------------------------ */
/* replaces SceneContext::make_scene + make_integrator (scene_context.rs:675-731):
 * builds the SAH BVH on the host, flattens it to the GPU node layout, uploads. */
int pt_scene_create(const pt_scene_desc* desc, int device, pt_scene** out);
void pt_scene_destroy(pt_scene* scene);
/* Re-sends every flattened scene array (nodes, triangle records, attributes, materials, lights, sampler tables)
 * from the pinned host staging copies kept since pt_scene_create to the device: what a host that rebuilds or
 * edits the scene between frames pays per render.  bytes (may be NULL) receives the host->device byte count. */
int pt_scene_reupload(pt_scene* scene, uint64_t* bytes);
const char* pt_last_error(void);
int pt_abi_version(void);

/* replaces Integrator::render (core/integrator/sampler.rs:259-325 + integrators/path.rs:61-241
 * + Film::write_image film.rs:440-484 up to the RGB float buffer).
 * out_rgb: host buffer, 3 floats per pixel of film.cropped_pixel_bounds, row-major.
 * With shard_count > 1 the result is this shard's PARTIAL film; use pt_render_partial +
 * pt_film_resolve (or the NCCL path in the Python host) to combine. */
int pt_render(pt_scene* scene, const pt_render_opts* opts, float* out_rgb, pt_stats* stats);

/* Same, but leaves the partial film {X,Y,Z,sum_w} (4 floats per pixel) on the device and
 * returns the device pointer (mirrors Film.pixels, film.rs:29-32, before write_image). */
int pt_render_partial(pt_scene* scene, const pt_render_opts* opts, float** dev_film_xyzw, pt_stats* stats);
/* Film::write_image's arithmetic on a device film buffer (film.rs:447-472): xyz->rgb / sum_w, clamp, scale. */
int pt_film_resolve(pt_scene* scene, const float* dev_film_xyzw, float* out_rgb_host);
/* zero the device film (Film::clear, film.rs:404-412) */
int pt_film_clear(pt_scene* scene);

/* ---- stage entry points (each replaces one function of the path; used by parity tests
 *      and by the kernel-level bench) ------------------------------------------------- */
/* Scene::intersect (core/scene/scene.rs:44-48 -> QBVHAccel::intersect qbvh_x86.rs:378-388).
 * rays: n x {ox,oy,oz,tmax, dx,dy,dz,time} (8 floats, HOST).  Outputs (HOST, may be NULL):
 * prim[n] = ORIGINAL top-level primitive index hit (triangle index >= 0, instances report the
 * instanced triangle's global index) or -1; t[n]; bary[3n] = b0,b1,b2. */
int pt_intersect(pt_scene* scene, const float* rays, uint32_t n, int32_t* prim, float* t, float* bary,
                 uint64_t* nodes_visited, uint64_t* tris_tested);
/* Scene::intersect_p (scene.rs:50-54 -> qbvh_x86.rs:390-399): occluded[n] = 0/1 */
int pt_intersect_p(pt_scene* scene, const float* rays, uint32_t n, uint8_t* occluded);
/* Device-resident variant for benchmarking the dominant kernel: rays already in HBM
 * (two float4 arrays o_tmax[n], d_time[n]); hits written to dev_hits (float4 b0,b1,b2,prim bits). */
int pt_intersect_device(pt_scene* scene, const void* dev_o_tmax, const void* dev_d_time, uint32_t n,
                        void* dev_hits, int repeat, float* ms_per_launch);

/* Sampler values (samplers/halton.rs:149-162, sobol.rs:168-185, random.rs): for each of n
 * (pixel x, pixel y, sample number) triples write dims [dim0, dim0+ndims) of the stream. */
int pt_sampler_values(pt_scene* scene, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n,
                      uint32_t dim0, uint32_t ndims, float* out);
/* PerspectiveCamera::generate_ray_differential (cameras/perspective.rs:121-183) on the sample
 * stream of the scene's sampler: out rays n x 8 floats (o,tmax,d,time), out p_film n x 2. */
int pt_camera_rays(pt_scene* scene, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n,
                   float* rays, float* p_film);
/* PathIntegrator::li for explicit samples (integrators/path.rs:61-241): out n x n_channels radiance. */
int pt_li(pt_scene* scene, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* out_l);

/* BVH export for build parity (accelerators/bvh/build/ *.rs, accel/qbvh/qbvh_x86.rs:93-176):
 * object k's ordered primitive list and node count.  ordered may be NULL to query sizes. */
int pt_bvh_info(pt_scene* scene, uint32_t object, uint32_t* n_nodes, uint32_t* n_prims, int32_t* ordered);
/* copy the reference-layout view of node i of object k: boxes[2][3][4], children[4] (-1 = empty),
 * axes[3], is_leaf */
int pt_bvh_node(pt_scene* scene, uint32_t object, uint32_t node, float* boxes24, int64_t* children4,
                int32_t* axes3, int32_t* is_leaf);

/* Host-only variant of the two calls above (no GPU needed): builds every object's accelerator exactly as
 * pt_scene_create does.  Used by the build-parity tests and by tools that only need the tree. */
typedef struct pt_host_bvh pt_host_bvh;
int pt_host_bvh_build(const pt_scene_desc* desc, pt_host_bvh** out);
/* n_nodes = node count of the reference's flattening; gpu_nodes = 128-byte GPU nodes; qdepth = depth of the 4-wide tree */
int pt_host_bvh_info(const pt_host_bvh* h, uint32_t object, uint32_t* n_nodes, uint32_t* n_prims, int32_t* ordered, float* bounds6,
                     uint32_t* gpu_nodes, uint32_t* qdepth);
int pt_host_bvh_node(const pt_host_bvh* h, uint32_t object, uint32_t node, float* boxes24, int64_t* children4, int32_t* axes3,
                     int32_t* is_leaf);
void pt_host_bvh_destroy(pt_host_bvh* h);

/* SpatialLightDistribution::lookup (core/lightdistrib/spatial.rs:200-261): for n points write the
 * voxel's cdf-ready per-light func values (n x n_lights). */
int pt_light_distribution(pt_scene* scene, const float* points, uint32_t n, float* out_func);

/* ---- host front end (the reference's own L5/L6 layers, C++ mirror because rustc is absent) ---- */
typedef struct pt_host_scene pt_host_scene;
/* PbrtOptions (core/options/pbrt_options.rs:6-12), the process-wide options the reference's main sets from its command
 * line (src/bin/pbrt.rs:363-369) and its factories read: --quick renders at 1/4 resolution (film.rs:548-554) with one
 * sample per pixel (halton.rs:288-290, sobol.rs:199-201), --quick-full-resolution keeps the resolution. */
typedef struct pt_options {
  int32_t quick_render;
  int32_t quick_render_full_resolution;
} pt_options;
void pt_set_options(const pt_options* options);   /* PbrtOptions::set; NULL restores the defaults */
void pt_get_options(pt_options* out);
/* pbrt_parse_file + SceneContext (core/parser/parser.rs:60, scene_context.rs): parse a .pbrt file
 * (subset: see DESIGN.md) into a scene description.  Overrides < 0 / NULL mean "keep". */
int pt_parse_file(const char* path, int pixelsamples_override, int spectrum_kind, pt_host_scene** out);
int pt_parse_string(const char* text, int pixelsamples_override, int spectrum_kind, pt_host_scene** out);
const pt_scene_desc* pt_host_scene_desc(const pt_host_scene* hs);
const char* pt_host_scene_filename(const pt_host_scene* hs);   /* Film "filename" */
/* The warnings the parse produced (the reference logs them with warn!: unknown materials falling back to matte,
 * create_material.rs:65-70; invalid meshes; unreadable SPD files, ...): message i < count, else NULL. */
uint32_t pt_host_scene_warning_count(const pt_host_scene* hs);
const char* pt_host_scene_warning(const pt_host_scene* hs, uint32_t i);
void pt_host_scene_destroy(pt_host_scene* hs);

#ifdef __cplusplus
}
#endif
#endif /* PBRT_R3_B200_H */
