// dshade.cuh — wavefront stages of the PathIntegrator bounce loop: camera-ray generation, vertex shading
// (emission, one-light NEE with MIS, BSDF continuation, Russian roulette), shadow / MIS resolution.
// Replaces reference src/core/integrator/sampler.rs:201-257 (render_tile's per-sample body),
// src/cameras/perspective.rs:121-183, src/integrators/path.rs:61-241, src/core/integrator/sample_lights.rs:129-176,
// 330-453, src/lights/{diffuse,point,distant}.rs, src/shapes/triangle.rs:590-651, src/core/shape/shape.rs:40-54,
// src/core/lightdistrib/spatial.rs:86-111,200-261 (lookup), src/core/sampling/distribution.rs:13-32,86-101.
//
// One path = one slot of the SoA state below.  li()'s sequential loop becomes: trace(A rays) -> shade -> trace(B shadow,
// C MIS, A continuation rays) -> shade -> ...; the three scene queries per vertex keep their own random numbers, and
// every sample dimension is consumed in the reference's order inside shade, so the Halton/Sobol streams are reproduced.
#pragma once
#include "dbsdf.cuh"
#include "dsampler.cuh"
#include "dtrav.cuh"

namespace ptd {

struct WaveState {            // all arrays have `cap` path slots; spectra are path-major (st_load / st_store): RGB = one
                              // 16-byte record {r, g, b, -} per path, 60 bins = 240 contiguous bytes per path (15 x 16-byte
                              // accesses, whole sectors: the path ids of a shading bin are scattered, so a channel-major
                              // layout costs one 32-byte sector per FLOAT)
  uint32_t cap;
  uint32_t seg_cap;           // waves in flight own disjoint segments of the slots: segment = path id / seg_cap (render_waves)
  uint32_t nseg;
  f4* ray_o;                  // o.xyz, tmax   (continuation / camera ray; tmax shrinks to the hit distance)
  f4* ray_d;                  // d.xyz, time
  f4* hit;                    // b0, b1, b2, bits(tri) (tri = -1: miss)
  int* hit_inst;              // instance index or -1 (only when the scene has instances)
  float* beta;                // NC
  float* L;                   // NC
  float* eta_scale;
  uint32_t* flags;            // bit0 specular_bounce, bits 8.. bounces
  uint32_t* dim;              // next sampler dimension
  long long* sindex;          // sampler index of this sample (halton / sobol)
  int* pix;                   // packed pixel: (y << 16 | x) relative to sample_bounds.min
  f2* p_film;
  f4* sh_o; f4* sh_d;         // shadow ray (d unnormalised, tmax = 1 - eps)
  float* pendA;               // NC: light-sample term, cleared when occluded
  f4* mis_o; f4* mis_d;       // BSDF-sample MIS ray
  float* pendB;               // NC: BSDF-sample term, cleared unless the ray hits the sampled light's emitting side
  int* mis_light;
  // integrators that read sample arrays or walk the specular tree (ambientocclusion, directlighting, whitted); null otherwise
  int* sample_num;            // the camera sample's number within its pixel (selects the slice of a sample array)
  uint32_t* arr_off;          // get_2d_array calls so far in this sample (base_sampler.rs:59-70)
  // the pending specular_transmit branches of the tree walk, level-major [level * cap + path]: the vertex (ray, hit) to
  // come back to, its throughput, its depth (fr_ro.w)
  uint32_t n_frames;
  f4* fr_ro; f4* fr_rd; f4* fr_hit; f4* fr_beta;
  int* fr_inst;
};

struct Queues {
  uint32_t* ext;       // path ids with a continuation ray to trace
  uint32_t* ext_next;
  uint32_t* shadow;
  uint32_t* mis;
  uint32_t* bins;      // traced continuation rays binned by shading class: bins[cls * cap + k]
  uint32_t* counts;    // [0] ext, [1] ext_next, [2] shadow, [3] mis, [4] trace fetch counter, [5] wait_next, [8 + cls] bin sizes,
                       // [16 + segment] continuation rays pushed by the last shade stage for each wave in flight,
                       // [24..29) copy of [1..6) taken after the trace stage (k_ao / k_ri read their input list sizes there)
  uint32_t* wait;      // directlighting / whitted: paths between two lighting rounds of one vertex (no continuation ray)
  uint32_t* wait_next;
};
#define PT_MAX_SEGMENTS 8

// Shading classes: one specialised shade kernel per material plugin ("sort by material"); class 0
// takes the paths that need no BSDF: misses, paths at max depth, surfaces without a material.
// CL_EXTRA: uber, substrate and translucent share one generic kernel with room for five lobes (they are not on the
// BASELINE configs; the five classes above keep their two-lobe BSDF in registers).  At most 8 classes (bin counters).
enum { CL_NONE = 0, CL_MATTE = 1, CL_PLASTIC = 2, CL_GLASS = 3, CL_METAL = 4, CL_MIRROR = 5, CL_EXTRA = 6, CL_COUNT = 7 };
PT_HD constexpr int cls_lobes(int c) { return c == CL_EXTRA ? PT_LOBES_EXTRA : PT_MAX_LOBES; }
PT_HD constexpr int cls_cm(int c) {
  return c == CL_MATTE ? PT_CM_MATTE : c == CL_PLASTIC ? PT_CM_PLASTIC : c == CL_GLASS ? PT_CM_GLASS : c == CL_METAL ? PT_CM_METAL : c == CL_MIRROR ? PT_CM_MIRROR : PT_CM_ANY;
}
PT_HD constexpr int cls_mt(int c) { return (c >= CL_MATTE && c <= CL_MIRROR) ? c - 1 : -1; }
// class of a traced continuation ray (pt_material_type + 1)
PT_HD int shade_class(const DScene& S, int tri, int bounces) {
  if (tri < 0 || bounces >= S.integrator.max_depth) return CL_NONE;
  int mi = ldg(S.tri_material + tri);
  if (mi < 0) return CL_NONE;
  int t = S.materials[mi].type;
  return (t >= 0 && t <= 4) ? t + 1 : ((t >= PT_MAT_UBER && t <= PT_MAT_MIX) ? CL_EXTRA : CL_NONE);
}

// {Perspective,Orthographic,Environment}Camera::generate_ray_differential without the differentials (constant textures:
// they cannot reach L)
PT_HD DRay camera_ray(const pt_camera& C, const int32_t* full_res, f2 p_film, f2 p_lens, float time) {
  DRay r;
  r.tmax = pt_inf();
  r.time = lerpf(time, C.shutter_open, C.shutter_close);
  if (C.type == PT_CAMERA_ENVIRONMENT) {  // environment.rs:36-58
    const float theta = PT_PI * p_film.y / (float)full_res[1], phi = 2.0f * PT_PI * p_film.x / (float)full_res[0];
    r.o = mk3(0.0f, 0.0f, 0.0f);
    r.d = mk3(pt_sinf(theta) * pt_cosf(phi), pt_cosf(theta), pt_sinf(theta) * pt_sinf(phi));
    return xform_ray(C.camera_to_world, r);
  }
  f3 pc = mat_point(C.raster_to_camera, mk3(p_film.x, p_film.y, 0.0f));
  if (C.type == PT_CAMERA_ORTHOGRAPHIC) {  // orthographic.rs:60-90: parallel rays from the film plane
    r.o = pc;
    r.d = mk3(0.0f, 0.0f, 1.0f);
  } else {
    r.o = mk3(0.0f, 0.0f, 0.0f);
    r.d = norm3(pc);
  }
  if (C.lens_radius > 0.0f) {
    f2 dl = concentric_disk(p_lens);
    f2 pl = mk2(dl.x * C.lens_radius, dl.y * C.lens_radius);
    float ft = C.focal_distance / r.d.z;
    f3 pf = r.o + r.d * ft;
    r.o = mk3(pl.x, pl.y, 0.0f);
    r.d = norm3(pf - r.o);
  }
  return xform_ray(C.camera_to_world, r);
}

// ---- lights ------------------------------------------------------------------------------------------------
struct RefPt { f3 p, perr, n; };

PT_HD f3 tri_geo_normal_for_sample(const DScene& S, int tri, f3 p0, f3 p1, f3 p2, float b0, float b1) {
  uint32_t flags = ldg(S.tri_flags + tri);
  f3 n = norm3(cross3(p1 - p0, p2 - p0));
  if ((flags & PT_TRI_HAS_N) && S.tri_n) {
    const float* q = S.tri_n + (size_t)tri * 9;
    f3 ns = b0 * mk3(ldg(q), ldg(q + 1), ldg(q + 2)) + b1 * mk3(ldg(q + 3), ldg(q + 4), ldg(q + 5)) + (1.0f - b0 - b1) * mk3(ldg(q + 6), ldg(q + 7), ldg(q + 8));
    n = faceforward(n, ns);
  } else if (((flags & PT_TRI_REVERSE_ORIENTATION) != 0) != ((flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) {
    n = n * -1.0f;
  }
  return n;
}

// ---- InfiniteAreaLight (lights/infinite.rs) --------------------------------------------------------------------------
// MIPMap::lookup(st, 0.0) = triangle(0, st) (mipmap.rs:620-637, 711-767), s repeats, t clamps (mipmap.rs:1013-1026)
PT_HD f3 env_lookup0(const EnvDev& E, float s_, float t_) {
  float s = s_ * (float)E.w - 0.5f, t = t_ * (float)E.h - 0.5f;
  int s0 = (int)floorf(s), t0 = (int)floorf(t);
  float ds = s - (float)s0, dt = t - (float)t0;
  int s1 = s0 + 1, t1 = t0 + 1;
  s0 &= E.w - 1; s1 &= E.w - 1;
  t0 = t0 < 0 ? 0 : (t0 > E.h - 1 ? E.h - 1 : t0);
  t1 = t1 < 0 ? 0 : (t1 > E.h - 1 ? E.h - 1 : t1);
  f3 a = xyz(ldg4(E.texels + (size_t)t0 * E.w + s0)) * ((1.0f - ds) * (1.0f - dt));
  f3 b = xyz(ldg4(E.texels + (size_t)t1 * E.w + s0)) * ((1.0f - ds) * dt);
  f3 c = xyz(ldg4(E.texels + (size_t)t0 * E.w + s1)) * (ds * (1.0f - dt));
  f3 d = xyz(ldg4(E.texels + (size_t)t1 * E.w + s1)) * (ds * dt);
  return a + b + c + d;
}
template <int NC>
PT_HD Sp<NC> sp_from_rgb3(f3 c) {  // Spectrum::from_rgb for RGBSpectrum (the front end admits infinite lights with RGB spectra only)
  Sp<NC> r = sp_const<NC>(0.0f);
  r.c[0] = c.x; r.c[NC > 1 ? 1 : 0] = c.y; r.c[NC > 2 ? 2 : 0] = c.z;
  return r;
}
PT_HD f3 mat3_vec(const float* m, f3 v) { return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }
PT_HD float spherical_theta(f3 v) { return acosf(clampf(v.z, -1.0f, 1.0f)); }                             // geometry/misc.rs:96-104
PT_HD float spherical_phi(f3 v) { float p = atan2f(v.y, v.x); return p < 0.0f ? p + 2.0f * PT_PI : p; }
// InfiniteAreaLight::le (infinite.rs:110-121)
template <int NC>
PT_HD1 Sp<NC> env_le(const DScene& S, int li, f3 d) {
  const EnvDev& E = S.envs[S.lights[li].tri];
  f3 w = norm3(mat3_vec(E.w2l, d));
  float s = spherical_phi(w) * (PT_INV_PI * 0.5f), t = spherical_theta(w) * PT_INV_PI;
  return sp_from_rgb3<NC>(env_lookup0(E, s, t));
}
// Distribution1D::sample_continuous (distribution.rs:66-86) on cdf[0..n], func[0..n)
PT_HD float dist1d_sample(const float* cdf, const float* func, float func_int, int n, float u, float* pdf, int* off) {
  uint32_t first = 0, len = (uint32_t)n + 1u;
  while (len > 0) {
    uint32_t half = len >> 1, mid = first + half;
    if (ldg(cdf + mid) <= u) { first = mid + 1; len -= half + 1; }
    else len = half;
  }
  int o = first == 0 ? 0 : (int)first - 1;
  if (o > n - 1) o = n - 1;
  float c0 = ldg(cdf + o), c1 = ldg(cdf + o + 1);
  float du = u - c0;
  if (c1 - c0 > 0.0f) du /= c1 - c0;
  *pdf = func_int > 0.0f ? ldg(func + o) / func_int : 0.0f;
  *off = o;
  return ((float)o + du) * (1.0f / (float)n);
}
// Distribution2D::pdf (distribution.rs:139-145) at (phi / 2 pi, theta / pi)
PT_HD1 float env_pdf(const DScene& S, int li, f3 w_world) {  // InfiniteAreaLight::pdf_li (infinite.rs:168-185)
  const EnvDev& E = S.envs[S.lights[li].tri];
  f3 wi = mat3_vec(E.w2l, w_world);
  float theta = spherical_theta(wi), phi = spherical_phi(wi);
  float sin_theta = clampf(pt_sinf(theta), 0.0f, 1.0f);
  if (sin_theta == 0.0f) return 0.0f;
  float pu = phi * (PT_INV_PI * 0.5f) * (float)E.nu, pv = theta * PT_INV_PI * (float)E.nv;
  int iu = pu > 0.0f ? (pu >= (float)E.nu ? E.nu - 1 : (int)pu) : 0, iv = pv > 0.0f ? (pv >= (float)E.nv ? E.nv - 1 : (int)pv) : 0;
  return (ldg(E.func + (size_t)iv * E.nu + iu) / E.m_int) / (2.0f * PT_PI * PT_PI * sin_theta);
}

template <int NC>
struct LightSampleD {
  Sp<NC> li;
  f3 wi;
  float pdf;
  f3 sh_o, sh_d;  // VisibilityTester: p0.spawn_ray_to(p1)
};

// Light::sample_li for area (one triangle), point, spot, distant and - ENV, compiled in only where the scene has one -
// infinite lights
template <int NC, bool ENV = false>
PT_HD1 bool light_sample(const DScene& S, int li, const RefPt& it, f2 u, LightSampleD<NC>* out) {
  const pt_light& L = S.lights[li];
  f3 lp, lperr = mk3(0.0f, 0.0f, 0.0f), ln = mk3(0.0f, 0.0f, 0.0f);
  if (ENV && L.type == PT_LIGHT_INFINITE) {  // infinite.rs:123-166
    const EnvDev& E = S.envs[L.tri];
    float pdf1, pdf0;
    int v, uo;
    float d1 = dist1d_sample(E.mcdf, E.row_int, E.m_int, E.nv, u.y, &pdf1, &v);
    float d0 = dist1d_sample(E.cdf + (size_t)v * (E.nu + 1), E.func + (size_t)v * E.nu, ldg(E.row_int + v), E.nu, u.x, &pdf0, &uo);
    float map_pdf = pdf0 * pdf1;
    if (map_pdf <= 0.0f) return false;
    float theta = d1 * PT_PI, phi = d0 * 2.0f * PT_PI;
    float cos_theta = pt_cosf(theta), sin_theta = clampf(pt_sinf(theta), 0.0f, 1.0f);
    float cos_phi = pt_cosf(phi), sin_phi = pt_sinf(phi);
    f3 wi = mat3_vec(E.l2w, mk3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
    out->wi = wi;
    out->pdf = sin_theta == 0.0f ? 0.0f : map_pdf / (2.0f * PT_PI * PT_PI * sin_theta);
    out->li = sp_from_rgb3<NC>(env_lookup0(E, d0, d1));
    lp = it.p + wi * (2.0f * S.world_radius);
  } else if (L.type == PT_LIGHT_AREA) {
    f3 p0, p1, p2;
    tri_verts(S, L.tri, &p0, &p1, &p2);
    float su0 = sqrtf(u.x);
    float b0 = 1.0f - su0, b1 = u.y * su0;  // uniform_sample_triangle
    f3 p = b0 * p0 + b1 * p1 + (1.0f - b0 - b1) * p2;
    f3 n = tri_geo_normal_for_sample(S, L.tri, p0, p1, p2, b0, b1);
    f3 pas = abs3(b0 * p0) + abs3(b1 * p1) + abs3((1.0f - b0 - b1) * p2);
    lperr = PT_GAMMA6_REF * mk3(pas.x, pas.y, pas.z);
    float area = 0.5f * len3(cross3(p1 - p0, p2 - p0));
    float pdf_area = 1.0f / area;
    f3 wi = p - it.p;
    if (len2(wi) <= 0.0f) return false;
    wi = norm3(wi);
    if (!(ldg(S.tri_flags + L.tri) & PT_TRI_TWO_SIDED)) {
      if (dot3(n, -wi) <= 0.0f) return false;
    }
    float pdf = pdf_area * dist2(it.p, p) / absdot3(n, -wi);
    if (pdf <= 0.0f || isinf_(pdf)) return false;
    if (len2(p - it.p) <= 0.0f) return false;
    out->wi = norm3(p - it.p);
    out->pdf = pdf;
    if (L.two_sided || dot3(n, -out->wi) > 0.0f) out->li = sp_load<NC>(S.spectra + (size_t)L.spectrum * NC);
    else out->li = sp_const<NC>(0.0f);
    lp = p; ln = n;
  } else if (L.type == PT_LIGHT_POINT) {
    f3 v = mk3(L.v[0], L.v[1], L.v[2]);
    out->li = sp_load<NC>(S.spectra + (size_t)L.spectrum * NC) / dist2(v, it.p);
    out->wi = norm3(v - it.p);
    out->pdf = 1.0f;
    lp = v;
  } else if (ENV && (L.type == PT_LIGHT_PROJECTION || L.type == PT_LIGHT_GONIOMETRIC)) {
    // projection.rs:64-110, goniometric.rs:33-62: a point light times an image looked up in the direction -wi
    f3 v = mk3(L.v[0], L.v[1], L.v[2]);
    const EnvDev& E = S.envs[L.tri];
    f3 wi = norm3(v - it.p);
    f3 w = -wi;
    f3 rgb = mk3(0.0f, 0.0f, 0.0f);
    if (L.type == PT_LIGHT_PROJECTION) {
      f3 wl = mat3_vec(E.w2l, w);
      if (!(wl.z < E.hitcher)) {
        f3 p = mat_point(E.xf, wl);
        if (p.x >= E.screen[0] && p.x <= E.screen[2] && p.y >= E.screen[1] && p.y <= E.screen[3]) {
          float sx = p.x - E.screen[0], sy = p.y - E.screen[1];  // Bounds2f::offset (bounds2.rs:44-53)
          if (E.screen[2] > E.screen[0]) sx = sx / (E.screen[2] - E.screen[0]);
          if (E.screen[3] > E.screen[1]) sy = sy / (E.screen[3] - E.screen[1]);
          rgb = env_lookup0(E, sx, sy);
        }
      }
    } else {
      f3 wp = mat_point(E.xf, w);  // the direction as a POINT: the light's translation is applied (goniometric.rs:34)
      rgb = env_lookup0(E, spherical_phi(wp) * (PT_INV_PI * 0.5f), spherical_theta(wp) * PT_INV_PI);
    }
    out->li = (sp_load<NC>(S.spectra + (size_t)L.spectrum * NC) * sp_from_rgb3<NC>(rgb)) / dist2(v, it.p);
    out->wi = wi;
    out->pdf = 1.0f;
    lp = v;
  } else if (L.type == PT_LIGHT_SPOT) {  // lights/spot.rs:44-57 (falloff), 62-80 (sample_li)
    f3 v = mk3(L.v[0], L.v[1], L.v[2]);
    const pt_spot& P = S.spots[L.tri];
    f3 wi = norm3(v - it.p);
    f3 w = -wi;
    f3 wl = norm3(mk3(P.w2l[0] * w.x + P.w2l[1] * w.y + P.w2l[2] * w.z, P.w2l[3] * w.x + P.w2l[4] * w.y + P.w2l[5] * w.z,
                      P.w2l[6] * w.x + P.w2l[7] * w.y + P.w2l[8] * w.z));
    float ct = wl.z, fall;
    if (ct < P.cos_total_width) fall = 0.0f;
    else if (ct >= P.cos_falloff_start) fall = 1.0f;
    else {
      float delta = (ct - P.cos_total_width) / (P.cos_falloff_start - P.cos_total_width);
      fall = (delta * delta) * (delta * delta);
    }
    out->li = sp_load<NC>(S.spectra + (size_t)L.spectrum * NC) * (fall / dist2(v, it.p));
    out->wi = wi;
    out->pdf = 1.0f;
    lp = v;
  } else {
    f3 v = mk3(L.v[0], L.v[1], L.v[2]);
    out->li = sp_load<NC>(S.spectra + (size_t)L.spectrum * NC);
    out->wi = v;
    out->pdf = 1.0f;
    lp = it.p + v * (2.0f * S.world_radius);
  }
  // Interaction::spawn_ray_to (interaction.rs:118-127)
  f3 origin = offset_origin(it.p, it.perr, it.n, lp - it.p);
  f3 target = offset_origin(lp, lperr, ln, origin - lp);
  out->sh_o = origin;
  out->sh_d = target - origin;
  return true;
}

// Light::pdf_li for the area light's triangle = Shape::pdf_from (re-intersect that one triangle)
template <bool ENV = false>
PT_HD1 float light_pdf_li(const DScene& S, int li, const RefPt& it, f3 wi) {
  const pt_light& L = S.lights[li];
  if (ENV && L.type == PT_LIGHT_INFINITE) return env_pdf(S, li, wi);
  if (L.type != PT_LIGHT_AREA) return 0.0f;
  f3 o = offset_origin(it.p, it.perr, it.n, wi);
  f3 p0, p1, p2;
  tri_verts(S, L.tri, &p0, &p1, &p2);
  TriHitD th;
  if (!tri_test(p0, p1, p2, ldg(S.tri_flags + L.tri), o, wi, pt_inf(), &th)) return 0.0f;
  f3 sp, sn;
  tri_point_normal(S, L.tri, wi, th.b0, th.b1, th.b2, &sp, &sn);
  float area = 0.5f * len3(cross3(p1 - p0, p2 - p0));
  float pdf = dist2(it.p, sp) / (absdot3(sn, -wi) * area);
  if (isinf_(pdf)) return 0.0f;
  return pdf;
}

// SurfaceInteraction::le -> DiffuseAreaLight::l
template <int NC>
PT_HD Sp<NC> surface_le(const DScene& S, const SurfD& si, f3 w) {
  int li = ldg(S.tri_light + si.tri);
  if (li < 0) return sp_const<NC>(0.0f);
  const pt_light& L = S.lights[li];
  if (L.two_sided || dot3(si.n, w) > 0.0f) return sp_load<NC>(S.spectra + (size_t)L.spectrum * NC);
  return sp_const<NC>(0.0f);
}

// LightDistribution::lookup + Distribution1D::sample_discrete on the dense voxel table
PT_HD uint32_t ld_voxel(const LightDistDev& D, f3 p) {
  float o[3] = {p.x - D.wb_min[0], p.y - D.wb_min[1], p.z - D.wb_min[2]};
  uint32_t pi[3];
  for (int i = 0; i < 3; i++) {
    if (D.wb_max[i] > D.wb_min[i]) o[i] = div_rn(o[i], D.wb_max[i] - D.wb_min[i]);
    float c = clampf(o[i], 0.0f, 1.0f);
    uint32_t v = (uint32_t)(c * (float)D.voxels[i]);  // `as u32` saturates; c >= 0
    pi[i] = v < D.voxels[i] - 1 ? v : D.voxels[i] - 1;
  }
  return (pi[0] * D.voxels[1] + pi[1]) * D.voxels[2] + pi[2];
}
PT_HD1 int ld_sample(const DScene& S, f3 p, float u, float* pdf) {
  const LightDistDev& D = S.ld;
  uint32_t nl = S.n_lights;
  uint32_t vox = D.strategy == PT_LIGHTS_SPATIAL ? ld_voxel(D, p) : 0u;
  const float* cdf = D.cdf + (size_t)vox * (nl + 1);
  // find_interval(cdf.len(), |i| cdf[i] <= u) (distribution.rs:13-32)
  uint32_t first = 0, len = nl + 1;
  while (len > 0) {
    uint32_t half = len >> 1, mid = first + half;
    if (ldg(cdf + mid) <= u) { first = mid + 1; len -= half + 1; }
    else len = half;
  }
  uint32_t off = first == 0 ? 0 : first - 1;
  if (off > nl - 1) off = nl - 1;
  float func_int = ldg(D.func_int + vox);
  float f = ldg(D.func + (size_t)vox * nl + off);
  *pdf = func_int > 0.0f ? f * (1.0f / (float)nl) / func_int : 0.0f;
  return (int)off;
}

// ---- path state helpers ------------------------------------------------------------------------------------------
// floats per path of a spectral state array
PT_HD constexpr uint32_t st_stride(uint32_t nc) { return nc == 3 ? 4u : nc; }
template <int NC> PT_HD Sp<NC> st_load(const float* a, uint32_t cap, uint32_t i) {
  Sp<NC> r;
  if (NC == 3) {
    f4 v = reinterpret_cast<const f4*>(a)[i];
    r.c[0] = v.x; r.c[1] = v.y; r.c[NC - 1] = v.z;
    return r;
  }
  static_assert(NC == 3 || NC % 4 == 0, "spectral state records are read as 16-byte vectors");
  const f4* p = reinterpret_cast<const f4*>(a + (size_t)i * NC);
#pragma unroll
  for (int k = 0; k < NC / 4; k++) {
    f4 v = p[k];
    r.c[4 * k] = v.x; r.c[4 * k + 1] = v.y; r.c[4 * k + 2] = v.z; r.c[4 * k + 3] = v.w;
  }
  (void)cap;
  return r;
}
template <int NC> PT_HD void st_store(float* a, uint32_t cap, uint32_t i, const Sp<NC>& v) {
  if (NC == 3) {
    reinterpret_cast<f4*>(a)[i] = mk4(v.c[0], v.c[1], v.c[NC - 1], 0.0f);
    return;
  }
  f4* p = reinterpret_cast<f4*>(a + (size_t)i * NC);
#pragma unroll
  for (int k = 0; k < NC / 4; k++) p[k] = mk4(v.c[4 * k], v.c[4 * k + 1], v.c[4 * k + 2], v.c[4 * k + 3]);
  (void)cap;
}

#if defined(__CUDA_ARCH__)
PT_D uint32_t queue_push(uint32_t* counter) {  // warp-aggregated append (ballot / popc)
  unsigned m = __activemask();
  int leader = __ffs(m) - 1;
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
#else
inline uint32_t queue_push(uint32_t* counter) { return (*counter)++; }
#endif
// the path goes on: append it to the next extend queue and count it for its wave (the host retires a wave when the
// shade stage pushed no continuation ray for its segment)
PT_HD void push_next(const WaveState& W, Queues& Q, uint32_t i) {
  Q.ext_next[queue_push(&Q.counts[1])] = i;
  if (W.nseg > 1) {
    const uint32_t seg = i / W.seg_cap;
#if defined(__CUDA_ARCH__)
    unsigned m = __activemask();
    unsigned peers = __match_any_sync(m, seg);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&Q.counts[16 + seg], (uint32_t)__popc(peers));
#else
    Q.counts[16 + seg]++;
#endif
  }
}

// the appends of one shaded vertex to the shadow, MIS and next extend queues, warp-aggregated, the three atomics in flight
// together (one round trip)
// count_sh: shadow rays keep a wave alive as well (ambientocclusion: a path is nothing but its current occlusion ray)
PT_HD void push_rays(const WaveState& W, Queues& Q, uint32_t i, bool sh, bool mis, bool ext, bool count_sh = false) {
#if defined(__CUDA_ARCH__)
  const unsigned m = __activemask();
  const unsigned bs = __ballot_sync(m, sh), bm = __ballot_sync(m, mis), be = __ballot_sync(m, ext);
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t b0 = 0, b1 = 0, b2 = 0;
  if (lane == leader) {
    if (bs) b0 = atomicAdd(&Q.counts[2], (uint32_t)__popc(bs));
    if (bm) b1 = atomicAdd(&Q.counts[3], (uint32_t)__popc(bm));
    if (be) b2 = atomicAdd(&Q.counts[1], (uint32_t)__popc(be));
  }
  b0 = __shfl_sync(m, b0, leader);
  b1 = __shfl_sync(m, b1, leader);
  b2 = __shfl_sync(m, b2, leader);
  const unsigned lt = (1u << lane) - 1u;
  if (sh) Q.shadow[b0 + (uint32_t)__popc(bs & lt)] = i;
  if (mis) Q.mis[b1 + (uint32_t)__popc(bm & lt)] = i;
  if (ext) Q.ext_next[b2 + (uint32_t)__popc(be & lt)] = i;
  const bool alive = ext || (count_sh && sh);
  if (alive && W.nseg > 1) {  // rays per wave in flight that keep it alive (the host retires a wave that pushed none)
    const uint32_t seg = i / W.seg_cap;
    const unsigned peers = __match_any_sync(count_sh ? (be | bs) : be, seg);
    if (lane == __ffs(peers) - 1) atomicAdd(&Q.counts[16 + seg], (uint32_t)__popc(peers));
  }
#else
  if (sh) Q.shadow[Q.counts[2]++] = i;
  if (mis) Q.mis[Q.counts[3]++] = i;
  if (ext) push_next(W, Q, i);
  else if (count_sh && sh && W.nseg > 1) Q.counts[16 + i / W.seg_cap]++;
#endif
}

// camera-ray generation for path slot i of the current wave: pixel (px,py) absolute, sample number n
// Sampler "random": tile_slot / tile_seq select the PCG32 stream (packed into st.index: slot << 32 | sequence number = tile
// index of the whole film); the path is appended to the queue at position qpos instead of i.
template <int NC>
PT_HD void raygen_body(const DScene& S, WaveState& W, Queues& Q, uint32_t i, int px, int py, int n, int tile_slot = -1, uint32_t tile_seq = 0u,
                       uint32_t qpos = 0xffffffffu) {
  SampleStream st;
  st.px = px; st.py = py; st.dim = 0;
  if (tile_slot >= 0) st.index = (int64_t)(((uint64_t)(uint32_t)tile_slot << 32) | (uint64_t)tile_seq);
  else st.index = sampler_index(S, px, py, n);
  f2 u = get2d(S, st);
  f2 pf = mk2((float)px + u.x, (float)py + u.y);
  f2 pl = get2d(S, st);
  float tm = get1d(S, st);
  DRay r = camera_ray(S.camera, S.film.full_resolution, pf, pl, tm);
  W.ray_o[i] = mk4(r.o.x, r.o.y, r.o.z, r.tmax);
  W.ray_d[i] = mk4(r.d.x, r.d.y, r.d.z, r.time);
  W.p_film[i] = pf;
  W.sindex[i] = st.index;
  W.dim[i] = st.dim;
  W.pix[i] = ((py - S.sampler.sb[1]) << 16) | (px - S.sampler.sb[0]);
  W.flags[i] = 0u;
  W.eta_scale[i] = 1.0f;
  st_store<NC>(W.beta, W.cap, i, sp_const<NC>(1.0f));
  st_store<NC>(W.L, W.cap, i, sp_const<NC>(0.0f));
  st_store<NC>(W.pendA, W.cap, i, sp_const<NC>(0.0f));
  st_store<NC>(W.pendB, W.cap, i, sp_const<NC>(0.0f));
  if (W.sample_num) { W.sample_num[i] = n; W.arr_off[i] = 0u; }
  Q.ext[qpos == 0xffffffffu ? i : qpos] = i;
}

// One iteration of PathIntegrator::li's loop body after scene.intersect (path.rs:87-233) for path slot i.
// ENV: the scene has infinite lights (their code is compiled into a second set of shade kernels, so that scenes without
// one run the same kernels as before): escaped rays reach this stage (class CL_NONE) and add Le of the environment.
template <int NC, int CLS, bool ENV = false>
PT_HD void shade_body(const DScene& S, WaveState& W, Queues& Q, uint32_t i, unsigned long long* vertex_counter) {
  constexpr int CM = cls_cm(CLS), MT = cls_mt(CLS);
  const uint32_t cap = W.cap;
  // contributions of the previous vertex (shadow / MIS rays were resolved by the trace stage)
  Sp<NC> L = st_load<NC>(W.L, cap, i);
  {
    Sp<NC> a = st_load<NC>(W.pendA, cap, i), b = st_load<NC>(W.pendB, cap, i);
    L = L + a;
    L = L + b;
    st_store<NC>(W.pendA, cap, i, sp_const<NC>(0.0f));
    st_store<NC>(W.pendB, cap, i, sp_const<NC>(0.0f));
  }
  f4 ro4 = W.ray_o[i], rd4 = W.ray_d[i], h4 = W.hit[i];
  f3 ro = xyz(ro4), rd = xyz(rd4);
  float rtime = rd4.w;
  int tri = (int)fbits(h4.w);
  bool found = tri >= 0;
  uint32_t fl = W.flags[i];
  int bounces = (int)(fl >> 8);
  bool specular_bounce = (fl & 1u) != 0;
  Sp<NC> beta = st_load<NC>(W.beta, cap, i);
  SurfD si;
  if (found) {
    int inst = S.has_instances ? W.hit_inst[i] : -1;
    si = make_surface(S, tri, inst, ro, rd, rtime, h4.x, h4.y, h4.z);
  }
  if (bounces == 0 || specular_bounce) {
    if (found) L = L + beta * surface_le<NC>(S, si, -rd);
    else if (ENV) {  // path.rs:92-95: every infinite light's Le along the escaped ray
      for (uint32_t k = 0; k < S.n_inf; k++) L = L + beta * env_le<NC>(S, ldg(S.inf_lights + k), rd);
    }
  }
  if (!found || bounces >= S.integrator.max_depth) { st_store<NC>(W.L, cap, i, L); return; }
  BsdfT<cls_lobes(CLS)> B;
  B.n = -1;
  if (CLS != CL_NONE) make_bsdf<NC, MT>(S, si, &B);
  RefPt ref;
  ref.p = si.p; ref.perr = si.perr; ref.n = si.n;
  if (CLS == CL_NONE || B.n < 0) {  // no BSDF: continue through the surface, bounces unchanged (path.rs:108-111)
    f3 o = offset_origin(si.p, si.perr, si.n, rd);
    W.ray_o[i] = mk4(o.x, o.y, o.z, pt_inf());
    st_store<NC>(W.L, cap, i, L);
    push_next(W, Q, i);
    return;
  }
  if (vertex_counter) {
#if defined(__CUDA_ARCH__)
    atomicAdd(vertex_counter, 1ull);
#else
    (*vertex_counter)++;
#endif
  }
  // The three queue appends of a vertex (shadow ray, MIS ray, continuation ray) are made together at the end: a warp-aggregated
  // append waits for its atomic's return value, and three of them spread over the body cost three L2 round trips per warp
  // (ncu, matte kernel of config 2: 11 % of the stall samples on the shuffles that hand the base out).
  bool push_sh = false, push_mis = false, push_ext = false;
  SampleStream st;
  st.index = W.sindex[i];
  st.dim = W.dim[i];
  int pk = W.pix[i];
  st.px = (pk & 0xffff) + S.sampler.sb[0];
  st.py = (pk >> 16) + S.sampler.sb[1];
  const int NONSPEC = F_ALL & ~F_SPEC;

  // --- uniform_sample_one_light_surface + estimate_direct_surface (sample_lights.rs:129-176, 330-453)
  if (CLS != CL_MIRROR && bsdf_count<CM>(B, NONSPEC) > 0 && S.n_lights > 0) {
    // path.rs:122-129 adds beta * uniform_sample_one_light(..) even when the estimate is zero: a non-finite beta makes
    // l invalid at this vertex whatever the light / MIS rays find (inf * 0 = NaN).  The terms parked in pendA / pendB
    // are dropped when occluded, so the same effect is folded in here; for a finite beta this adds +0.
    L = L + beta * 0.0f;
    float light_pdf;
    int ln = ld_sample(S, si.p, get1d(S, st), &light_pdf);
#ifdef PT_DEBUG_PRINT
    fprintf(stderr, "[hc] p=(%g %g %g) n=(%g %g %g) perr=(%g %g %g) light %d pdf %g dim %u\n", si.p.x, si.p.y, si.p.z, si.n.x, si.n.y, si.n.z, si.perr.x, si.perr.y, si.perr.z, ln, light_pdf, st.dim);
#endif
    if (light_pdf > 0.0f) {
      f2 u_light = get2d(S, st);
      f2 u_scat = get2d(S, st);
      const pt_light& LT = S.lights[ln];
      bool delta = LT.type != PT_LIGHT_AREA && !(ENV && LT.type == PT_LIGHT_INFINITE);
      LightSampleD<NC> ls;
      if (light_sample<NC, ENV>(S, ln, ref, u_light, &ls)) {
#ifdef PT_DEBUG_PRINT
        fprintf(stderr, "[hc] ls pdf %g li %g wi=(%g %g %g) sh_o=(%.9g %.9g %.9g) sh_d=(%.9g %.9g %.9g)\n", ls.pdf, ls.li.c[0], ls.wi.x, ls.wi.y, ls.wi.z, ls.sh_o.x, ls.sh_o.y, ls.sh_o.z, ls.sh_d.x, ls.sh_d.y, ls.sh_d.z);
#endif
        if (ls.pdf > 0.0f && !sp_black(ls.li)) {
          Sp<NC> f;
          float spdf;
          if (NC == 3) {  // RGB: one fused pass (shares the microfacet terms); 60 bins: the separate functions move fewer spectra
            bsdf_f_pdf<NC, CM>(S, B, si.wo, ls.wi, NONSPEC, &f, &spdf);
            f = f * absdot3(ls.wi, si.sn);
          } else {
            f = bsdf_f<NC, CM>(S, B, si.wo, ls.wi, NONSPEC) * absdot3(ls.wi, si.sn);
            spdf = bsdf_pdf<CM>(B, si.wo, ls.wi, NONSPEC);
          }
          if (!sp_black(f)) {
            Sp<NC> ld;
            if (delta) ld = f * ls.li / ls.pdf;
            else ld = f * ls.li * (power_heuristic1(ls.pdf, spdf) / ls.pdf);
            st_store<NC>(W.pendA, cap, i, beta * (ld / light_pdf));
            W.sh_o[i] = mk4(ls.sh_o.x, ls.sh_o.y, ls.sh_o.z, 0.0f);
            W.sh_d[i] = mk4(ls.sh_d.x, ls.sh_d.y, ls.sh_d.z, 1.0f - PT_SHADOW_EPS);
            push_sh = true;
          }
        }
      }
      if (!delta) {
        Sp<NC> f2v;
        f3 wi2;
        float pdf2;
        int stype;
        if (bsdf_sample<NC, CM>(S, B, si.wo, u_scat, NONSPEC, &f2v, &wi2, &pdf2, &stype)) {
          Sp<NC> f = f2v * absdot3(wi2, si.sn);
          if (!sp_black(f) && pdf2 > 0.0f) {
            float weight = 1.0f;
            bool go = true;
            if (!(stype & F_SPEC)) {
              float lp = light_pdf_li<ENV>(S, ln, ref, wi2);
              if (lp == 0.0f) go = false;
              else weight = power_heuristic1(pdf2, lp);
            }
            if (go) {
              // an area light's Le is its constant; an infinite light's is looked up along wi (the ray must then ESCAPE, mis_resolve)
              Sp<NC> li = (ENV && LT.type == PT_LIGHT_INFINITE) ? env_le<NC>(S, ln, wi2) : sp_load<NC>(S.spectra + (size_t)LT.spectrum * NC);
              Sp<NC> ld = f * li * sp_const<NC>(1.0f) * (weight / pdf2);
              st_store<NC>(W.pendB, cap, i, beta * (ld / light_pdf));
              f3 o = offset_origin(si.p, si.perr, si.n, wi2);
              W.mis_o[i] = mk4(o.x, o.y, o.z, pt_inf());
              W.mis_d[i] = mk4(wi2.x, wi2.y, wi2.z, rtime);
              W.mis_light[i] = ln;
              push_mis = true;
            }
          }
        }
      }
    }
  }

  // --- sample the BSDF for the next direction (path.rs:139-167)
  st_store<NC>(W.L, cap, i, L);
  do {
    f3 wo = -rd;
    Sp<NC> f;
    f3 wi;
    float pdf;
    int sflags;
    f2 ub = get2d(S, st);
    if (!bsdf_sample<NC, CM>(S, B, wo, ub, F_ALL, &f, &wi, &pdf, &sflags)) break;
    if (sp_black(f) || pdf == 0.0f) break;
    beta = beta * (f * (absdot3(wi, si.sn) / pdf));
    specular_bounce = (sflags & F_SPEC) != 0;
    float eta_scale = W.eta_scale[i];
    if ((sflags & F_SPEC) && (sflags & F_TRANS)) {
      float eta = B.eta;
      eta_scale *= (dot3(wo, si.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta);
    }
    f3 o = offset_origin(si.p, si.perr, si.n, wi);
    // Russian roulette (path.rs:219-229)
    Sp<NC> rr = beta * eta_scale;
    float rrmax = sp_max(rr);
    if (rrmax < S.integrator.rr_threshold && bounces > 3) {
      float q = fmaxf(0.05f, 1.0f - rrmax);
      if (get1d(S, st) < q) break;
      beta = beta / (1.0f - q);
    }
    bounces += 1;
    st_store<NC>(W.beta, cap, i, beta);
    W.eta_scale[i] = eta_scale;
    W.flags[i] = ((uint32_t)bounces << 8) | (specular_bounce ? 1u : 0u);
    W.dim[i] = st.dim;
    W.ray_o[i] = mk4(o.x, o.y, o.z, pt_inf());
    W.ray_d[i] = mk4(wi.x, wi.y, wi.z, rtime);
    push_ext = true;
  } while (0);
  push_rays(W, Q, i, push_sh, push_mis, push_ext);
}

// AOVIntegrator::li (integrators/aov.rs:101-181) for path slot i after its camera ray was traced: a geometric quantity
// of the first hit as colour, clamped to [0, 1] and scaled; black on a miss.  RGB only.
template <int NC>
PT_HD void aov_body(const DScene& S, WaveState& W, uint32_t i) {
  f4 ro4 = W.ray_o[i], rd4 = W.ray_d[i], h4 = W.hit[i];
  const int tri = (int)fbits(h4.w);
  Sp<NC> out = sp_const<NC>(0.0f);
  if (tri >= 0) {
    f3 ro = xyz(ro4), rd = xyz(rd4);
    const int inst = S.has_instances ? W.hit_inst[i] : -1;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    const int target = S.integrator.aov_target;
    if (target == PT_AOV_DISTANCE) { r = g = b = ro4.w / len3(rd); }
    else if (target == PT_AOV_UV) {
      f2 uv0 = mk2(0.0f, 0.0f), uv1 = mk2(1.0f, 0.0f), uv2 = mk2(1.0f, 1.0f);
      if ((ldg(S.tri_flags + tri) & PT_TRI_HAS_UV) && S.tri_uv) {
        const float* q = S.tri_uv + (size_t)tri * 6;
        uv0 = mk2(ldg(q), ldg(q + 1)); uv1 = mk2(ldg(q + 2), ldg(q + 3)); uv2 = mk2(ldg(q + 4), ldg(q + 5));
      }
      r = h4.x * uv0.x + h4.y * uv1.x + h4.z * uv2.x;
      g = h4.x * uv0.y + h4.y * uv1.y + h4.z * uv2.y;
    } else {
      SurfD si = make_surface(S, tri, inst, ro, rd, rd4.w, h4.x, h4.y, h4.z);
      if (target == PT_AOV_DEPTH) { r = g = b = len3(si.p - ro); }
      else {
        f3 v = target == PT_AOV_N ? si.n : (target == PT_AOV_NS ? si.sn : si.sdpdu);
        f3 c = 0.5f * v + mk3(0.5f, 0.5f, 0.5f);
        r = c.x; g = c.y; b = c.z;
      }
    }
    const float sc = S.integrator.aov_scale;
    out.c[0] = clampf(r, 0.0f, 1.0f) * sc;
    out.c[1] = clampf(g, 0.0f, 1.0f) * sc;
    out.c[NC > 2 ? 2 : 0] = clampf(b, 0.0f, 1.0f) * sc;
  }
  st_store<NC>(W.L, W.cap, i, out);
}

// AOIntegrator::li (integrators/ao.rs:50-105) as a per-path state machine: flags >> 8 = number of occlusion rays made so
// far.  Round 0 (after the camera ray's closest-hit query): a miss ends the path, a surface without a material lets the
// ray through (ao.rs:68-71), anything else makes occlusion ray 0.  Round r > 0 (after occlusion ray r - 1 was traced: the
// trace stage has cleared pendA if it was occluded): l += pendA, then occlusion ray r; the last ray's term is folded in by
// k_finish, in the same order.  Element r of get_2d_array(n) for sample number s of the pixel is dimensions (5, 6) of the
// index of sample NUMBER n * s + r (GlobalSampler::start_pixel, halton.rs:193-207, base_sampler.rs:59-70).
// `Q.shadow` is the queue being filled (the one the last trace stage read is the caller's input list).  RGB only.
template <int NC>
PT_HD void ao_body(const DScene& S, WaveState& W, Queues& Q, uint32_t i) {
  const int ns = S.integrator.ao_samples;
  const int round = (int)(W.flags[i] >> 8);
  f4 ro4 = W.ray_o[i], rd4 = W.ray_d[i], h4 = W.hit[i];
  const int tri = (int)fbits(h4.w);
  bool push_sh = false, push_ext = false;
  if (tri >= 0 && round < ns) {
    f3 ro = xyz(ro4), rd = xyz(rd4);
    const int inst = S.has_instances ? W.hit_inst[i] : -1;
    SurfD si = make_surface(S, tri, inst, ro, rd, rd4.w, h4.x, h4.y, h4.z);
    if (round == 0 && ldg(S.tri_material + tri) < 0) {
      f3 o = offset_origin(si.p, si.perr, si.n, rd);
      W.ray_o[i] = mk4(o.x, o.y, o.z, pt_inf());
      push_ext = true;
    } else {
      if (round > 0) {
        Sp<NC> L = st_load<NC>(W.L, W.cap, i);
        L = L + st_load<NC>(W.pendA, W.cap, i);
        st_store<NC>(W.L, W.cap, i, L);
      }
      f3 n = faceforward(si.n, -rd);
      f3 s = norm3(si.dpdu);
      f3 t = cross3(si.n, s);
      const int pk = W.pix[i];
      const int px = (pk & 0xffff) + S.sampler.sb[0], py = (pk >> 16) + S.sampler.sb[1];
      f2 u;
      if (S.sampler.type >= PT_SAMPLER_RANDOM) {  // element round of this sample's slice of the pixel's array
        const float* a = S.tile_arrays + (size_t)((uint64_t)W.sindex[i] >> 32) * S.tile_arr_stride + S.tile_tab + 2u * (uint32_t)(ns * W.sample_num[i] + round);
        u = mk2(a[0], a[1]);
      } else {
        const int64_t idx = sampler_index(S, px, py, (int64_t)ns * (int64_t)W.sample_num[i] + (int64_t)round);
        u = mk2(sampler_dim(S, idx, 5u, px, py), sampler_dim(S, idx, 6u, px, py));
      }
      f3 wl;
      float pdf;
      if (S.integrator.ao_cos_sample) { wl = cosine_hemisphere(u); pdf = fabsf(wl.z) * PT_INV_PI; }
      else { wl = uniform_hemisphere(u); pdf = PT_INV_PI * 0.5f; }
      f3 wi = mk3(wl.x * s.x + wl.y * t.x + wl.z * n.x, wl.x * s.y + wl.y * t.y + wl.z * n.y, wl.x * s.z + wl.y * t.z + wl.z * n.z);
      st_store<NC>(W.pendA, W.cap, i, sp_const<NC>(dot3(wi, n) / (pdf * (float)ns)));
      f3 o = offset_origin(si.p, si.perr, si.n, wi);
      W.sh_o[i] = mk4(o.x, o.y, o.z, 0.0f);
      W.sh_d[i] = mk4(wi.x, wi.y, wi.z, pt_inf());
      W.flags[i] = (uint32_t)(round + 1) << 8;
      push_sh = true;
    }
  }
  push_rays(W, Q, i, push_sh, false, push_ext, true);
}

// ---- directlighting / whitted (integrators/directlighting.rs:66-125, whitted.rs:37-98) ------------------------------
// li() is a recursion over the TREE of specular reflections and transmissions (SamplerIntegrator::specular_reflect /
// specular_transmit, core/integrator/sampler.rs:37-160); the sampler's dimensions are consumed depth first.  One path slot
// walks its tree depth first, one ray at a time: a vertex whose reflection branch is taken leaves a frame (ray, hit,
// throughput, depth) for its transmission branch, which is sampled - at the dimension the sampler has reached by then -
// when the reflection subtree is finished.  Radiance is gathered as throughput x term into L (the recursion's nested
// f * li * w products in a different association: a few ulp).  The lighting of a vertex is a sequence of ROUNDS, one
// light sample each (shadow ray [+ MIS ray] per round, through the same pendA / pendB mechanism as the path integrator):
// whitted: one per light; directlighting "one": a single round on a uniformly picked light; "all": n_samples per light
// from the 2D sample arrays (dimensions (5, 6) of sample NUMBER n * s + k for both arrays - the reference fills every 2D
// array from the same two dimensions), or one get_2d pair per light once the max_depth * n_lights * 2 arrays are used up.
// flags: bits 0-7 frames on the stack, 8-15 depth, 16-30 lighting round of the vertex, 31 the vertex reads sample arrays.
#define PT_RI_MAX_DEPTH 32
// ML: lobes per BSDF (2, or 5 when the scene has uber / translucent materials: the wider BSDF costs 10-19 % of this stage)
// ENV: the scene has infinite lights (as for the shade kernels, their code is compiled into separate variants)
// WHITTED: the integrator, as a compile-time switch (each kernel then holds one lighting loop, not both)
// The body has ONE call site each for make_surface / make_bsdf and two for bsdf_sample (light sampling, specular branches):
// the popped transmission branch re-enters the loop at the branch stage.  (ncu, profiles/r2z: the first, fully divergent
// launch of a step stalled 19 cycles per issue on instruction fetch with three inlined copies of the sampling code.)
template <int NC, int ML, bool ENV, bool WHITTED>
PT_HD void ri_body(const DScene& S, WaveState& W, Queues& Q, uint32_t i, bool fresh, unsigned long long* vertex_counter) {
  const uint32_t cap = W.cap;
  constexpr int CM = ML == PT_MAX_LOBES ? (PT_CM_ANY & ~((1 << LB_LAMBERT_T) | (1 << LB_FRESNEL_BLEND))) : PT_CM_ANY;  // the extra materials' lobes only with them
  constexpr bool whitted = WHITTED;
  const int NONSPEC = F_ALL & ~F_SPEC;
  Sp<NC> L = st_load<NC>(W.L, cap, i);
  L = L + st_load<NC>(W.pendA, cap, i);
  L = L + st_load<NC>(W.pendB, cap, i);
  st_store<NC>(W.pendA, cap, i, sp_const<NC>(0.0f));
  st_store<NC>(W.pendB, cap, i, sp_const<NC>(0.0f));
  uint32_t fl = W.flags[i];
  uint32_t sp = fl & 0xffu, round = fresh ? 0u : ((fl >> 16) & 0x7fffu);
  int depth = (int)((fl >> 8) & 0xffu);
  bool arrays = !fresh && (fl >> 31) != 0u;
  f4 ro4 = W.ray_o[i], rd4 = W.ray_d[i], h4 = W.hit[i];
  int inst = S.has_instances ? W.hit_inst[i] : -1;
  Sp<NC> beta = st_load<NC>(W.beta, cap, i);
  SampleStream st;
  st.index = W.sindex[i];
  st.dim = W.dim[i];
  const int pk = W.pix[i];
  st.px = (pk & 0xffff) + S.sampler.sb[0];
  st.py = (pk >> 16) + S.sampler.sb[1];
  uint32_t arr_off = W.arr_off[i];
  {  // GlobalSampler: get_1d / get_2d skip the dimensions of the sample arrays, [5, 5 + 2 x arrays) (halton.rs:227-231, 241-245)
    const uint32_t n_arrays = (!whitted && S.integrator.direct_strategy == PT_DIRECT_ALL) ? 2u * S.n_lights * (uint32_t)S.integrator.max_depth : 0u;
    if (S.sampler.type < PT_SAMPLER_RANDOM && st.dim < 5u + 2u * n_arrays) st.dim = 5u + 2u * n_arrays;
  }
  const float rtime = rd4.w;
  bool push_sh = false, push_mis = false, push_ext = false, push_wait = false;
  bool need_pop = false, have_vertex = false;
  SurfD si;
  BsdfT<ML> B;
  B.n = -1;
  bool need_make = false, lit = false;  // lit: the vertex's lighting rounds are over (a popped transmission branch)
  int branch0 = 0;
  if ((int)fbits(h4.w) < 0) {  // the ray left the scene: li = sum of Light::le (directlighting.rs:112-122, whitted.rs:86-96)
    need_pop = true;
    if (ENV && fresh)
      for (uint32_t k = 0; k < S.n_inf; k++) L = L + beta * env_le<NC>(S, ldg(S.inf_lights + k), xyz(rd4));
  } else need_make = true;
  for (;;) {
    if (need_make) {  // the interaction and BSDF of the vertex in (ro4, rd4, h4): just traced, between two rounds, or popped
      need_make = false;
      si = make_surface(S, (int)fbits(h4.w), inst, xyz(ro4), xyz(rd4), rtime, h4.x, h4.y, h4.z);
      make_bsdf<NC, -1, false>(S, si, &B);
      if (B.n < 0) {  // no BSDF: whitted returns zero (whitted.rs:84), directlighting follows the ray at the same depth (:108-110)
        if (whitted) need_pop = true;
        else {
          f3 o = offset_origin(si.p, si.perr, si.n, xyz(rd4));
          W.ray_o[i] = mk4(o.x, o.y, o.z, pt_inf());
          push_ext = true;
        }
      } else {
        have_vertex = true;
        if (fresh && !lit) {
          if (vertex_counter) {
#if defined(__CUDA_ARCH__)
            atomicAdd(vertex_counter, 1ull);
#else
            (*vertex_counter)++;
#endif
          }
          if (!whitted) L = L + beta * surface_le<NC>(S, si, si.wo);
        }
      }
    }
    if (have_vertex) {
      RefPt ref;
      ref.p = si.p; ref.perr = si.perr; ref.n = si.n;
      const uint32_t nl = lit ? 0u : S.n_lights;
      const bool all = !whitted && S.integrator.direct_strategy == PT_DIRECT_ALL;
      if (all && round == 0u && nl > 0u) {  // uniform_sample_all_lights asks for two arrays per light (sample_lights.rs:38-39)
        arrays = arr_off < 2u * nl * (uint32_t)S.integrator.max_depth;
        if (arrays) arr_off += 2u * nl;
      }
      const uint32_t rounds = nl == 0u ? 0u : (whitted ? nl : (all ? (arrays ? ldg(S.light_sample_prefix + nl) : nl) : 1u));
      bool emitted = false;
      while (round < rounds && !emitted) {
        int ln;
        f2 u_light, u_scat;
        float divisor = 1.0f;  // the estimate is divided by light_pdf ("one") or by n_samples ("all" with arrays)
        if (whitted) { ln = (int)round; u_light = get2d(S, st); u_scat = u_light; }
        else if (!all) {  // uniform_sample_one_light without a distribution (sample_lights.rs:104-110)
          uint32_t k = (uint32_t)(get1d(S, st) * (float)nl);
          ln = (int)(k < nl - 1u ? k : nl - 1u);
          divisor = 1.0f / (float)nl;
          u_light = get2d(S, st);
          u_scat = get2d(S, st);
        } else if (!arrays) { ln = (int)round; u_light = get2d(S, st); u_scat = get2d(S, st); }
        else {
          uint32_t lo = 0u, hi = nl;  // the light whose samples hold round: prefix[lo] <= round < prefix[lo + 1]
          while (hi - lo > 1u) {
            uint32_t mid = (lo + hi) >> 1;
            if (ldg(S.light_sample_prefix + mid) <= round) lo = mid; else hi = mid;
          }
          ln = (int)lo;
          const uint32_t n = (uint32_t)S.lights[ln].n_samples, k = round - ldg(S.light_sample_prefix + lo);
          if (S.sampler.type >= PT_SAMPLER_RANDOM) {
            // arrays are laid out in request order: per depth, per light, (u_light, u_scattering), n x spp elements each;
            // this vertex holds arrays arr_off - 2 nl .. arr_off
            const uint32_t spp = (uint32_t)S.sampler.spp, d = (arr_off - 2u * nl) / (2u * nl);
            const uint32_t e0 = spp * (d * 2u * ldg(S.light_sample_prefix + nl) + 2u * ldg(S.light_sample_prefix + lo)) + n * (uint32_t)W.sample_num[i] + k;
            const float* a = S.tile_arrays + (size_t)((uint64_t)st.index >> 32) * S.tile_arr_stride + S.tile_tab;
            u_light = mk2(a[2u * e0], a[2u * e0 + 1u]);
            u_scat = mk2(a[2u * (e0 + spp * n)], a[2u * (e0 + spp * n) + 1u]);
          } else {
            const int64_t idx = sampler_index(S, st.px, st.py, (int64_t)n * (int64_t)W.sample_num[i] + (int64_t)k);
            u_light = mk2(sampler_dim(S, idx, 5u, st.px, st.py), sampler_dim(S, idx, 6u, st.px, st.py));
            u_scat = u_light;
          }
          divisor = (float)n;
        }
        round++;
        const pt_light& LT = S.lights[ln];
        LightSampleD<NC> ls;
        const bool ok = light_sample<NC, ENV>(S, ln, ref, u_light, &ls);
        if (whitted) {  // whitted.rs:61-79: BSDF over all lobes, shading normal, no MIS
          if (ok && !(ls.pdf <= 0.0f || sp_black(ls.li))) {
            Sp<NC> f = bsdf_f<NC, CM>(S, B, si.wo, ls.wi, F_ALL);
            if (!sp_black(f)) {
              st_store<NC>(W.pendA, cap, i, beta * (f * ls.li * (absdot3(ls.wi, si.sn) / ls.pdf)));
              W.sh_o[i] = mk4(ls.sh_o.x, ls.sh_o.y, ls.sh_o.z, 0.0f);
              W.sh_d[i] = mk4(ls.sh_d.x, ls.sh_d.y, ls.sh_d.z, 1.0f - PT_SHADOW_EPS);
              push_sh = true;
              emitted = true;
            }
          }
          continue;
        }
        // estimate_direct_surface (sample_lights.rs:330-453), as in shade_body
        const bool delta = LT.type != PT_LIGHT_AREA && !(ENV && LT.type == PT_LIGHT_INFINITE);
        if (ok && ls.pdf > 0.0f && !sp_black(ls.li)) {
          Sp<NC> f = bsdf_f<NC, CM>(S, B, si.wo, ls.wi, NONSPEC) * absdot3(ls.wi, si.sn);
          float spdf = bsdf_pdf<CM>(B, si.wo, ls.wi, NONSPEC);
          if (!sp_black(f)) {
            Sp<NC> ld;
            if (delta) ld = f * ls.li / ls.pdf;
            else ld = f * ls.li * (power_heuristic1(ls.pdf, spdf) / ls.pdf);
            st_store<NC>(W.pendA, cap, i, beta * (ld / divisor));
            W.sh_o[i] = mk4(ls.sh_o.x, ls.sh_o.y, ls.sh_o.z, 0.0f);
            W.sh_d[i] = mk4(ls.sh_d.x, ls.sh_d.y, ls.sh_d.z, 1.0f - PT_SHADOW_EPS);
            push_sh = true;
            emitted = true;
          }
        }
        if (!delta) {
          Sp<NC> f2v;
          f3 wi2;
          float pdf2;
          int stype;
          if (bsdf_sample<NC, CM>(S, B, si.wo, u_scat, NONSPEC, &f2v, &wi2, &pdf2, &stype)) {
            Sp<NC> f = f2v * absdot3(wi2, si.sn);
            if (!sp_black(f) && pdf2 > 0.0f) {
              float weight = 1.0f;
              bool go = true;
              if (!(stype & F_SPEC)) {
                float lp = light_pdf_li<ENV>(S, ln, ref, wi2);
                if (lp == 0.0f) go = false;
                else weight = power_heuristic1(pdf2, lp);
              }
              if (go) {
                Sp<NC> li = (ENV && LT.type == PT_LIGHT_INFINITE) ? env_le<NC>(S, ln, wi2) : sp_load<NC>(S.spectra + (size_t)LT.spectrum * NC);
                Sp<NC> ld = f * li * sp_const<NC>(1.0f) * (weight / pdf2);
                st_store<NC>(W.pendB, cap, i, beta * (ld / divisor));
                f3 o = offset_origin(si.p, si.perr, si.n, wi2);
                W.mis_o[i] = mk4(o.x, o.y, o.z, pt_inf());
                W.mis_d[i] = mk4(wi2.x, wi2.y, wi2.z, rtime);
                W.mis_light[i] = ln;
                push_mis = true;
                emitted = true;
              }
            }
          }
        }
      }
      if (round < rounds) { push_wait = true; break; }  // the next round waits for this one's rays
      // specular_reflect, then specular_transmit (directlighting.rs:101-104, whitted.rs:80-83)
      need_pop = true;
      if (depth + 1 < S.integrator.max_depth) {
        for (int branch = branch0; branch < 2 && need_pop; branch++) {
          f2 u = get2d(S, st);
          Sp<NC> f;
          f3 wi;
          float pdf;
          int stype;
          if (!bsdf_sample<NC, CM>(S, B, si.wo, u, (branch == 0 ? F_REFL : F_TRANS) | F_SPEC, &f, &wi, &pdf, &stype)) continue;
          const float wi_ns = absdot3(wi, si.sn);
          if (!(pdf > 0.0f && !sp_black(f) && wi_ns != 0.0f)) continue;
          if (branch == 0) {  // come back for the transmission branch
            const size_t fi = (size_t)sp * cap + i;
            W.fr_ro[fi] = mk4(ro4.x, ro4.y, ro4.z, bitsf((uint32_t)depth));
            W.fr_rd[fi] = rd4;
            W.fr_hit[fi] = h4;
            reinterpret_cast<f4*>(W.fr_beta)[fi] = mk4(beta.c[0], beta.c[1], beta.c[NC > 2 ? 2 : 0], 0.0f);
            if (W.fr_inst) W.fr_inst[fi] = inst;
            sp++;
          }
          beta = beta * (f * (wi_ns / pdf));
          f3 o = offset_origin(si.p, si.perr, si.n, wi);
          W.ray_o[i] = mk4(o.x, o.y, o.z, pt_inf());
          W.ray_d[i] = mk4(wi.x, wi.y, wi.z, rtime);
          depth += 1;
          push_ext = true;
          need_pop = false;
        }
      }
    }
    if (!need_pop) break;
    // this subtree is finished: the innermost pending transmission branch, or the end of the path
    if (sp == 0u) break;
    sp--;
    {
      const size_t fi = (size_t)sp * cap + i;
      ro4 = W.fr_ro[fi]; rd4 = W.fr_rd[fi]; h4 = W.fr_hit[fi];
      const f4 b4 = reinterpret_cast<const f4*>(W.fr_beta)[fi];
      beta.c[0] = b4.x; beta.c[1] = b4.y; beta.c[NC > 2 ? 2 : 0] = b4.z;
      inst = W.fr_inst ? W.fr_inst[fi] : -1;
      depth = (int)fbits(ro4.w);
    }
    // specular_transmit of the frame's vertex: back to the branch stage with the vertex rebuilt
    need_make = true; have_vertex = false; lit = true; branch0 = 1; need_pop = false;
  }
  st_store<NC>(W.L, cap, i, L);
  st_store<NC>(W.beta, cap, i, beta);
  W.flags[i] = (arrays ? 0x80000000u : 0u) | ((round & 0x7fffu) << 16) | ((uint32_t)depth << 8) | sp;
  W.dim[i] = st.dim;
  W.arr_off[i] = arr_off;
  push_rays(W, Q, i, push_sh, push_mis, push_ext);
  if (push_wait) {
    Q.wait_next[queue_push(&Q.counts[5])] = i;
    if (W.nseg > 1) {
      const uint32_t seg = i / W.seg_cap;
#if defined(__CUDA_ARCH__)
      unsigned m = __activemask();
      unsigned peers = __match_any_sync(m, seg);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&Q.counts[16 + seg], (uint32_t)__popc(peers));
#else
      Q.counts[16 + seg]++;
#endif
    }
  }
}

// MIS ray resolution: the BSDF-sampled ray contributes only if it hits the sampled light (pointer equality in the
// reference, sample_lights.rs:431-441) on its emitting side.
template <int NC>
PT_HD void mis_resolve(const DScene& S, WaveState& W, uint32_t i, const HitD& h, bool found, f3 rd) {
  bool keep = false;
  if (!found) {  // light.le(ray) (sample_lights.rs:443-445): non-zero for an infinite light only
    if (S.n_inf != 0u) keep = S.lights[W.mis_light[i]].type == PT_LIGHT_INFINITE;
  } else if (h.instance < 0) {
    int ln = W.mis_light[i];
    if (ldg(S.tri_light + h.tri) == ln) {
      const pt_light& LT = S.lights[ln];
      keep = LT.two_sided != 0;
      if (!keep) {
        f3 sp, sn;
        tri_point_normal(S, h.tri, rd, h.b0, h.b1, h.b2, &sp, &sn);
        keep = dot3(sn, -rd) > 0.0f;
      }
    }
  }
  if (!keep) st_store<NC>(W.pendB, W.cap, i, sp_const<NC>(0.0f));
}

// end of a wave: fold the last pending terms into L and apply validate_radiance_result (sampler.rs:151-176)
template <int NC>
PT_HD void finish_body(const DScene& S, WaveState& W, uint32_t i) {
  Sp<NC> L = st_load<NC>(W.L, W.cap, i);
  L = L + st_load<NC>(W.pendA, W.cap, i);
  L = L + st_load<NC>(W.pendB, W.cap, i);
  if (!sp_valid(L)) L = sp_const<NC>(0.0f);
  else {
    float y = sp_y(S, L);
    if (y < -1e-5f || isinf_(y)) L = sp_const<NC>(0.0f);
  }
  st_store<NC>(W.L, W.cap, i, L);
}

}  // namespace ptd
