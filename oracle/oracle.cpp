// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of pbrt-r3's PathIntegrator hot path (see o_math.h).
// oracle.cpp — C entry points (ctypes) mirroring the stage entry points of include/pbrt_r3_b200.h so that
// the parity tests can feed the SAME pt_scene_desc and the SAME inputs to both sides.
// PARITY UNPINNED at the image level (the Rust reference cannot be built here); leaf arithmetic is pinned by
// tests/test_oracle_known_answers.py against the reference's own known-answer tests.
#include "o_render.h"
#include <atomic>
#include <thread>

#include <cstdio>
#include <string>

using namespace orc;

namespace {
std::vector<uint32_t> g_sobol32;
std::vector<uint64_t> g_vdc;
std::string g_err;

template <typename T>
bool load_blob(const std::string& path, size_t n, std::vector<T>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  out->resize(n);
  size_t got = fread(out->data(), sizeof(T), n, f);
  fclose(f);
  return got == n;
}
}  // namespace

struct orc_scene {
  int nc;
  Scene<3>* s3 = nullptr;
  Scene<60>* s60 = nullptr;
  const pt_scene_desc* d;
  ~orc_scene() { delete s3; delete s60; }
};

#define DISPATCH(sc, expr3, expr60) ((sc)->nc == 3 ? (expr3) : (expr60))

template <int NC>
static void do_li(Scene<NC>& S, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* out_l, int32_t* first_tri, float* first_t,
                  uint64_t* counters) {
  Counters cnt;
  for (uint32_t i = 0; i < n; i++) {
    Sampler sampler(&S.scfg, 0);
    sampler.start_pixel(pixel_xy[2 * i], pixel_xy[2 * i + 1]);
    sampler.set_sample_number(sample_num[i]);
    CameraSample cs;
    Vec2 u = sampler.get_2d();
    cs.p_film = Vec2((Float)pixel_xy[2 * i], (Float)pixel_xy[2 * i + 1]) + u;
    cs.p_lens = sampler.get_2d();
    cs.time = sampler.get_1d();
    Ray ray = camera_ray(S.d->camera, S.d->film, cs);
    PathTrace tr;
    Spec<NC> l = integrator_li(S, ray, sampler, &cnt, &tr);
    for (int k = 0; k < NC; k++) out_l[(size_t)i * NC + k] = l.c[k];
    if (first_tri) first_tri[i] = tr.first_tri;
    if (first_t) first_t[i] = tr.first_t;
  }
  if (counters) {
    counters[0] = cnt.camera_rays; counters[1] = cnt.regular_rays; counters[2] = cnt.shadow_rays;
    counters[3] = cnt.nodes; counters[4] = cnt.tris; counters[5] = cnt.vertices;
  }
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_set_data_dir(const char* dir) {
  std::string d(dir);
  if (!load_blob(d + "/sobol32.bin", 1024 * 52, &g_sobol32) || !load_blob(d + "/vdc.bin", (25 + 26) * 52, &g_vdc)) {
    g_err = "cannot load sobol tables from " + d;
    return 1;
  }
  tables().sobol32 = g_sobol32.data();
  tables().vdc = g_vdc.data();
  return 0;
}

orc_scene* orc_scene_create(const pt_scene_desc* d) {
  orc_scene* s = new orc_scene();
  s->d = d;
  s->nc = (int)d->n_channels;
  if (s->nc == 3) { s->s3 = new Scene<3>(); s->s3->init(d); }
  else if (s->nc == 60) { s->s60 = new Scene<60>(); s->s60->init(d); }
  else { delete s; g_err = "n_channels must be 3 or 60"; return nullptr; }
  return s;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

// counters[6] = camera, regular, shadow, nodes, tris, vertices
int orc_render(orc_scene* s, int n_threads, int s0, int s1, float* out_rgb, float* out_xyzw, uint64_t* counters) {
  Counters c;
  if (s->nc == 3) render(*s->s3, n_threads, s0, s1, out_rgb, out_xyzw, &c);
  else render(*s->s60, n_threads, s0, s1, out_rgb, out_xyzw, &c);
  if (counters) { counters[0] = c.camera_rays; counters[1] = c.regular_rays; counters[2] = c.shadow_rays; counters[3] = c.nodes; counters[4] = c.tris; counters[5] = c.vertices; }
  return 0;
}

// orc_render restricted to the render tiles ti with ti % shard_count == shard_index
int orc_render_shard(orc_scene* s, int n_threads, int s0, int s1, int shard_index, int shard_count, float* out_rgb, float* out_xyzw, uint64_t* counters) {
  Counters c;
  if (s->nc == 3) render(*s->s3, n_threads, s0, s1, out_rgb, out_xyzw, &c, shard_index, shard_count);
  else render(*s->s60, n_threads, s0, s1, out_rgb, out_xyzw, &c, shard_index, shard_count);
  if (counters) { counters[0] = c.camera_rays; counters[1] = c.regular_rays; counters[2] = c.shadow_rays; counters[3] = c.nodes; counters[4] = c.tris; counters[5] = c.vertices; }
  return 0;
}

int orc_intersect(orc_scene* s, const float* rays, uint32_t n, int32_t* prim, float* t, float* bary, uint64_t* nodes, uint64_t* tris) {
  const SceneGeom& g = s->nc == 3 ? s->s3->geom : s->s60->geom;
  Counters c;
  for (uint32_t i = 0; i < n; i++) {
    const float* q = rays + (size_t)i * 8;
    Ray r(Vec3(q[0], q[1], q[2]), Vec3(q[4], q[5], q[6]), q[3], q[7]);
    Hit h;
    bool hit = g.objects[0]->intersect(g, r, &h, &c);
    if (prim) prim[i] = hit ? h.tri : -1;
    if (t) t[i] = hit ? h.t : 0.0f;
    if (bary) { bary[3 * i] = hit ? h.b0 : 0.0f; bary[3 * i + 1] = hit ? h.b1 : 0.0f; bary[3 * i + 2] = hit ? h.b2 : 0.0f; }
  }
  if (nodes) *nodes = c.nodes;
  if (tris) *tris = c.tris;
  return 0;
}
// ExhaustiveAccel::intersect (accelerators/exhaustive/accel.rs): no acceleration structure at all - every primitive of
// the scene in list order against the shrinking t_max, instances through TransformedPrimitive with an exhaustive loop
// over the object's own primitives.  Independent of the BVH build, the 4-wide flattening and the traversal order, so it
// catches errors that the oracle's QBVH and the device's share.  (The accelerator's world-bound pre-test is left out.)
int orc_intersect_exhaustive(orc_scene* s, const float* rays, uint32_t n, int n_threads, int32_t* prim, float* t, int32_t* instance) {
  const pt_scene_desc* d = s->d;
  auto obj_prims = [&](uint32_t k, const int32_t** p, size_t* np) { *p = d->prims + d->obj_first[k]; *np = d->obj_first[k + 1] - d->obj_first[k]; };
  auto one = [&](uint32_t i) {
    const float* q = rays + (size_t)i * 8;
    Ray r(Vec3(q[0], q[1], q[2]), Vec3(q[4], q[5], q[6]), q[3], q[7]);
    int best = -1, best_inst = -1;
    float best_t = 0.0f;
    const int32_t* top; size_t ntop;
    obj_prims(0, &top, &ntop);
    for (size_t a = 0; a < ntop; a++) {
      if (top[a] >= 0) {
        TriHit th;
        if (tri_intersect(get_tri(d, top[a]), r, &th)) { r.t_max = th.t; best = top[a]; best_inst = -1; best_t = th.t; }
      } else {
        const pt_instance& I = d->instances[~top[a]];
        Ray ro = transform_ray(mat_of(I.minv), r);
        const int32_t* op; size_t nop;
        obj_prims((uint32_t)I.object, &op, &nop);
        for (size_t b = 0; b < nop; b++) {
          TriHit th;
          if (op[b] >= 0 && tri_intersect(get_tri(d, op[b]), ro, &th)) { ro.t_max = th.t; r.t_max = th.t; best = op[b]; best_inst = ~top[a]; best_t = th.t; }
        }
      }
    }
    if (prim) prim[i] = best;
    if (t) t[i] = best_t;
    if (instance) instance[i] = best_inst;
  };
  std::atomic<uint32_t> next(0);
  auto worker = [&]() {
    for (;;) {
      uint32_t b = next.fetch_add(64);
      if (b >= n) break;
      for (uint32_t i = b; i < std::min(n, b + 64); i++) one(i);
    }
  };
  std::vector<std::thread> th;
  for (int k = 0; k < std::max(n_threads, 1); k++) th.emplace_back(worker);
  for (auto& x : th) x.join();
  return 0;
}

int orc_intersect_p(orc_scene* s, const float* rays, uint32_t n, uint8_t* occluded) {
  const SceneGeom& g = s->nc == 3 ? s->s3->geom : s->s60->geom;
  for (uint32_t i = 0; i < n; i++) {
    const float* q = rays + (size_t)i * 8;
    Ray r(Vec3(q[0], q[1], q[2]), Vec3(q[4], q[5], q[6]), q[3], q[7]);
    occluded[i] = g.objects[0]->intersect_p(g, r, nullptr) ? 1 : 0;
  }
  return 0;
}

int orc_sampler_values(orc_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, uint32_t dim0, uint32_t ndims, float* out) {
  const SamplerConfig& cfg = s->nc == 3 ? s->s3->scfg : s->s60->scfg;
  if (cfg.type >= PT_SAMPLER_RANDOM) { g_err = "random / stratified samplers have no index-based stream"; return 1; }
  for (uint32_t i = 0; i < n; i++) {
    Sampler sm(&cfg, 0);
    sm.start_pixel(pixel_xy[2 * i], pixel_xy[2 * i + 1]);
    sm.set_sample_number(sample_num[i]);
    for (uint32_t k = 0; k < ndims; k++) out[(size_t)i * ndims + k] = sm.sample_dimension(sm.interval_index, dim0 + k);
  }
  return 0;
}

int orc_camera_rays(orc_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* rays, float* p_film) {
  const SamplerConfig& cfg = s->nc == 3 ? s->s3->scfg : s->s60->scfg;
  for (uint32_t i = 0; i < n; i++) {
    Sampler sm(&cfg, 0);
    sm.start_pixel(pixel_xy[2 * i], pixel_xy[2 * i + 1]);
    sm.set_sample_number(sample_num[i]);
    CameraSample cs;
    Vec2 u = sm.get_2d();
    cs.p_film = Vec2((Float)pixel_xy[2 * i], (Float)pixel_xy[2 * i + 1]) + u;
    cs.p_lens = sm.get_2d();
    cs.time = sm.get_1d();
    Ray r = camera_ray(s->d->camera, s->d->film, cs);
    float* q = rays + (size_t)i * 8;
    q[0] = r.o.x; q[1] = r.o.y; q[2] = r.o.z; q[3] = r.t_max; q[4] = r.d.x; q[5] = r.d.y; q[6] = r.d.z; q[7] = r.time;
    if (p_film) { p_film[2 * i] = cs.p_film.x; p_film[2 * i + 1] = cs.p_film.y; }
  }
  return 0;
}

int orc_li(orc_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* out_l, int32_t* first_tri, float* first_t, uint64_t* counters) {
  if (s->nc == 3) do_li(*s->s3, pixel_xy, sample_num, n, out_l, first_tri, first_t, counters);
  else do_li(*s->s60, pixel_xy, sample_num, n, out_l, first_tri, first_t, counters);
  return 0;
}

int orc_bvh_info(orc_scene* s, uint32_t object, uint32_t* n_nodes, uint32_t* n_prims, int32_t* ordered, uint32_t* max_depth) {
  const SceneGeom& g = s->nc == 3 ? s->s3->geom : s->s60->geom;
  if (object >= g.objects.size()) return 1;
  const Accel& A = *g.objects[object];
  if (n_nodes) *n_nodes = (uint32_t)A.nodes.size();
  if (n_prims) *n_prims = (uint32_t)A.prims.size();
  if (max_depth) *max_depth = (uint32_t)A.max_depth;
  if (ordered) for (size_t i = 0; i < A.prims.size(); i++) ordered[i] = A.prims[i];
  return 0;
}
int orc_bvh_node(orc_scene* s, uint32_t object, uint32_t node, float* boxes24, int64_t* children4, int32_t* axes3, int32_t* is_leaf) {
  const SceneGeom& g = s->nc == 3 ? s->s3->geom : s->s60->geom;
  if (object >= g.objects.size() || node >= g.objects[object]->nodes.size()) return 1;
  const QNode& q = g.objects[object]->nodes[node];
  memcpy(boxes24, q.bboxes, sizeof(q.bboxes));
  for (int k = 0; k < 4; k++) children4[k] = q.children[k];
  axes3[0] = q.axis_top; axes3[1] = q.axis_left; axes3[2] = q.axis_right;
  *is_leaf = q.is_leaf;
  return 0;
}
int orc_world_bound(orc_scene* s, float* b6) {
  const Bounds3& b = s->nc == 3 ? s->s3->world_bound : s->s60->world_bound;
  b6[0] = b.mn.x; b6[1] = b.mn.y; b6[2] = b.mn.z; b6[3] = b.mx.x; b6[4] = b.mx.y; b6[5] = b.mx.z;
  return 0;
}

int orc_light_distribution(orc_scene* s, const float* points, uint32_t n, float* out_func) {
  uint32_t nl = s->d->n_lights;
  for (uint32_t i = 0; i < n; i++) {
    Vec3 p(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    const Distribution1D* D = s->nc == 3 ? s->s3->lookup(p) : s->s60->lookup(p);
    for (uint32_t j = 0; j < nl; j++) out_func[(size_t)i * nl + j] = D->func[j];
  }
  return 0;
}

// ---- leaf functions for the known-answer tests ----------------------------------------------------------
float orc_radical_inverse(uint32_t base_index, uint64_t a) { return radical_inverse(base_index, a); }
float orc_scrambled_radical_inverse(uint32_t base_index, uint64_t a, const uint16_t* perm) { return scrambled_radical_inverse(base_index, a, perm); }
float orc_scrambled_radical_inverse_default(uint32_t base_index, uint64_t a) {
  return scrambled_radical_inverse(base_index, a, tables().perms.data() + tables().prime_sums[base_index]);
}
uint64_t orc_inverse_radical_inverse(uint64_t base, uint64_t inverse, uint32_t ndigits) { return inverse_radical_inverse(base, inverse, ndigits); }
uint64_t orc_prime(uint32_t i) { return tables().primes[i]; }
uint64_t orc_prime_sum(uint32_t i) { return tables().prime_sums[i]; }
uint32_t orc_perm(uint64_t i) { return tables().perms[i]; }
float orc_sobol_sample(int64_t index, uint32_t dim, uint32_t scramble) { return sobol_sample_float(index, dim, scramble); }
uint64_t orc_sobol_interval_to_index(uint32_t m, uint64_t frame, int32_t px, int32_t py) { return sobol_interval_to_index(m, frame, px, py); }
void orc_rng_sequence(uint64_t seq, int use_seq, uint32_t n, uint32_t* out_u32, float* out_f) {
  RNG r = use_seq ? RNG(seq) : RNG();
  for (uint32_t i = 0; i < n; i++) {
    if (out_u32) out_u32[i] = r.uniform_uint32();
    else out_f[i] = r.uniform_float();
  }
}
void orc_distribution1d(const float* func, uint32_t n, float* cdf_out, float* func_int) {
  Distribution1D D(std::vector<Float>(func, func + n));
  for (uint32_t i = 0; i <= n; i++) cdf_out[i] = D.cdf[i];
  *func_int = D.func_int;
}
uint32_t orc_distribution1d_sample_discrete(const float* func, uint32_t n, float u, float* pdf) {
  Distribution1D D(std::vector<Float>(func, func + n));
  return (uint32_t)D.sample_discrete(u, pdf);
}
float orc_distribution1d_sample_continuous(const float* func, uint32_t n, float u, float* pdf, uint32_t* off) {
  Distribution1D D(std::vector<Float>(func, func + n));
  size_t o;
  float v = D.sample_continuous(u, pdf, &o);
  *off = (uint32_t)o;
  return v;
}
// one triangle given by 9 floats; ray = o(3) tmax d(3) time
int orc_tri_intersect_raw(const float* p9, uint32_t flags, const float* ray8, float* t, float* bary3) {
  TriRef T;
  T.p0 = Vec3(p9[0], p9[1], p9[2]); T.p1 = Vec3(p9[3], p9[4], p9[5]); T.p2 = Vec3(p9[6], p9[7], p9[8]);
  T.flags = flags;
  Ray r(Vec3(ray8[0], ray8[1], ray8[2]), Vec3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7]);
  TriHit h;
  if (!tri_intersect(T, r, &h)) return 0;
  if (t) *t = h.t;
  if (bary3) { bary3[0] = h.b0; bary3[1] = h.b1; bary3[2] = h.b2; }
  return 1;
}
// spawn a ray from a hit on a raw triangle towards direction w (tests/shapes.rs:156-198 reintersect)
void orc_spawn_ray_from_hit(const float* p9, uint32_t flags, const float* ray8, const float* bary3, const float* w3, float* out_ray8) {
  pt_scene_desc d;
  memset(&d, 0, sizeof(d));
  d.tri_p = p9; d.tri_flags = &flags; d.n_tris = 1;
  Ray r(Vec3(ray8[0], ray8[1], ray8[2]), Vec3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7]);
  SurfInt si = tri_interaction(&d, 0, r, bary3[0], bary3[1], bary3[2]);
  RefPoint rp; rp.p = si.p; rp.p_error = si.p_error; rp.n = si.n; rp.time = 0;
  Ray o = spawn_ray(rp, Vec3(w3[0], w3[1], w3[2]));
  out_ray8[0] = o.o.x; out_ray8[1] = o.o.y; out_ray8[2] = o.o.z; out_ray8[3] = o.t_max;
  out_ray8[4] = o.d.x; out_ray8[5] = o.d.y; out_ray8[6] = o.d.z; out_ray8[7] = 0;
}
// TrowbridgeReitz / Lambertian sampling for chi^2 tests (tests/bsdfs.rs:482-605)
// kind: 0 lambert, 1 microfacet reflection with TR(alpha, alpha, samplevis) and FresnelNoOp-like dielectric(1,1.5)
int orc_bxdf_sample(int kind, float alpha, int samplevis, const float* wo3, const float* u2, float* wi3, float* pdf) {
  BxDF<3> b;
  b.r = Spec<3>(1.0f);
  if (kind == 0) b.kind = BX_LAMBERT;
  else { b.kind = BX_MF_REFL; b.fresnel.kind = FR_NOOP; b.dist = TRDist(alpha, alpha, samplevis != 0); }
  Spec<3> f; Vec3 wi; Float p; int st;
  if (!b.sample_f(Vec3(wo3[0], wo3[1], wo3[2]), Vec2(u2[0], u2[1]), &f, &wi, &p, &st)) return 0;
  wi3[0] = wi.x; wi3[1] = wi.y; wi3[2] = wi.z; *pdf = p;
  return 1;
}
float orc_bxdf_pdf(int kind, float alpha, int samplevis, const float* wo3, const float* wi3) {
  BxDF<3> b;
  b.r = Spec<3>(1.0f);
  if (kind == 0) b.kind = BX_LAMBERT;
  else { b.kind = BX_MF_REFL; b.fresnel.kind = FR_NOOP; b.dist = TRDist(alpha, alpha, samplevis != 0); }
  return b.pdf(Vec3(wo3[0], wo3[1], wo3[2]), Vec3(wi3[0], wi3[1], wi3[2]));
}

}  // extern "C"
