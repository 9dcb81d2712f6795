// ORACLE — TEST INFRASTRUCTURE ONLY (see o_math.h header).
// o_render.h — perspective camera, PathIntegrator::li, one-light direct estimate, FilmTile/Film, tile driver.
// Follows reference src/cameras/perspective.rs:121-183, src/integrators/path.rs:61-241,
// src/core/integrator/sample_lights.rs:129-176,330-453, src/core/integrator/sampler.rs:151-325,
// src/core/film/{film,film_tile}.rs, src/core/spectrum/convert.rs.
#pragma once
#include "o_shading.h"

#include <atomic>
#include <cstdio>
#include <mutex>
#include <thread>

namespace orc {

struct CameraSample { Vec2 p_film, p_lens; Float time; };

// PerspectiveCamera::generate_ray_differential (perspective.rs:121-183), OrthographicCamera (orthographic.rs:60-90),
// EnvironmentCamera::generate_ray (environment.rs:36-58); differentials dropped: constant textures only
inline Ray camera_ray(const pt_camera& C, const pt_film& F, const CameraSample& s) {
  Mat4 r2c = mat_of(C.raster_to_camera), c2w = mat_of(C.camera_to_world);
  if (C.type == PT_CAMERA_ENVIRONMENT) {
    Float theta = PI * s.p_film.y / (Float)F.full_resolution[1];
    Float phi = 2.0f * PI * s.p_film.x / (Float)F.full_resolution[0];
    Vec3 dir(std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi));
    Ray ray(Vec3(0, 0, 0), dir, INF, lerp(s.time, C.shutter_open, C.shutter_close));
    return transform_ray(c2w, ray);
  }
  Vec3 p_camera = r2c.point(Vec3(s.p_film.x, s.p_film.y, 0.0f));
  Ray ray(Vec3(0, 0, 0), Vec3(0, 0, 0), INF, s.time);
  if (C.type == PT_CAMERA_ORTHOGRAPHIC) { ray.o = p_camera; ray.d = Vec3(0.0f, 0.0f, 1.0f); }
  else ray.d = normalize(p_camera);
  if (C.lens_radius > 0.0f) {
    Vec2 pl = concentric_sample_disk(s.p_lens) * C.lens_radius;
    Float ft = C.focal_distance / ray.d.z;
    Vec3 p_focus = ray.o + ray.d * ft;
    ray.o = Vec3(pl.x, pl.y, 0.0f);
    ray.d = normalize(p_focus - ray.o);
  }
  ray.time = lerp(s.time, C.shutter_open, C.shutter_close);
  return transform_ray(c2w, ray);
}

template <int NC>
struct Scene {
  const pt_scene_desc* d = nullptr;
  SceneGeom geom;
  Bounds3 world_bound;
  Float world_radius = 0;
  LightDistrib<NC> distrib;
  SamplerConfig scfg;
  std::mutex distrib_mutex;
  std::vector<EnvMap> envs;          // InfiniteAreaLight maps, indexed by pt_light.tri
  std::vector<int> infinite_lights;  // Scene.infinite_lights (scene.rs:24-25)

  void init(const pt_scene_desc* desc) {
    d = desc;
    geom.d = desc;
    geom.objects.resize(desc->n_objects);
    // instanced objects first (object k >= 1), then the top level (object 0)
    for (uint32_t k = desc->n_objects; k-- > 0;) {
      const int32_t* pr = desc->prims + desc->obj_first[k];
      size_t n = desc->obj_first[k + 1] - desc->obj_first[k];
      geom.objects[k] = build_accel(geom, pr, n, desc->accel, k == 0);
    }
    world_bound = geom.objects[0]->bounds;
    Vec3 dg = world_bound.mx - world_bound.mn;
    world_radius = length(dg) * 0.5f;  // Bounds3f::bounding_sphere (bounds3.rs:184-188)
    envs.resize(desc->n_envmaps);
    for (uint32_t k = 0; k < desc->n_envmaps; k++) envs[k].init(desc->envmaps[k]);
    for (uint32_t i = 0; i < desc->n_lights; i++)
      if (desc->lights[i].type == PT_LIGHT_INFINITE) infinite_lights.push_back((int)i);
    if (desc->n_lights > 0) distrib.init(desc, world_bound, world_radius, envs.data());
    scfg = make_sampler_config(desc->sampler.type, desc->sampler.samples_per_pixel, desc->film.sample_bounds, desc->sampler.sample_at_pixel_center != 0);
    if (desc->sampler.type == PT_SAMPLER_ZEROTWO) {  // ZeroTwoSequenceSampler::new (zerotwosequence.rs:13-26)
      uint32_t p2 = 1;
      while (p2 < (uint32_t)std::max(desc->sampler.samples_per_pixel, 1)) p2 *= 2;
      scfg.spp = (int)p2;
      scfg.strat_dims = desc->sampler.n_sampled_dimensions;
    }
    if (desc->sampler.type == PT_SAMPLER_STRATIFIED) {  // create_stratified_sampler (stratified.rs:130-150)
      scfg.strat_n = desc->sampler.strata;
      scfg.strat_dims = desc->sampler.n_sampled_dimensions;
      scfg.jitter = desc->sampler.jitter != 0;
    }
    if (desc->integrator.type == PT_INTEGRATOR_AO) {  // AOIntegrator::new: request_2d_array(n_samples) (ao.rs:29)
      scfg.n_arrays_2d = 1;
      scfg.array_sizes_2d.assign(1, (uint32_t)desc->integrator.ao_samples);
    }
    if (desc->integrator.type == PT_INTEGRATOR_DIRECT && desc->integrator.direct_strategy == PT_DIRECT_ALL) {  // directlighting.rs:57-63
      scfg.n_arrays_2d = 2u * n_sampled_lights() * (uint32_t)std::max(desc->integrator.max_depth, 0);
      for (int dd = 0; dd < desc->integrator.max_depth; dd++)
        for (uint32_t j = 0; j < n_sampled_lights(); j++) { scfg.array_sizes_2d.push_back((uint32_t)desc->lights[j].n_samples); scfg.array_sizes_2d.push_back((uint32_t)desc->lights[j].n_samples); }
    }
  }
  // scene.lights: the lights that can be sampled (area lights inside object instances only emit when hit; front end)
  uint32_t n_sampled_lights() const { return d->n_lights; }
  bool intersect(const Ray& r, Hit* h, Counters* c) const {
    if (c) c->regular_rays++;
    return geom.objects[0]->intersect(geom, r, h, c);
  }
  bool intersect_p(const Ray& r, Counters* c) const {
    if (c) c->shadow_rays++;
    return geom.objects[0]->intersect_p(geom, r, c);
  }
  const Distribution1D* lookup(Vec3 p) {
    std::lock_guard<std::mutex> g(distrib_mutex);
    return distrib.lookup(p);
  }
};

template <int NC>
inline Spec<NC> surface_le(const pt_scene_desc* d, const SurfInt& si, Vec3 w) {  // SurfaceInteraction::le
  int li = d->tri_light[si.tri];
  if (li < 0) return Spec<NC>();
  return area_l<NC>(d, d->lights[li], si.n, w);
}

// estimate_direct_surface (sample_lights.rs:330-453), handle_media = false, specular = false
template <int NC>
inline Spec<NC> estimate_direct(Scene<NC>& sc, const SurfInt& it, const BSDF<NC>& bsdf, Vec2 u_scattering, int light_num, Vec2 u_light, Counters* cnt) {
  const pt_scene_desc* d = sc.d;
  const int flags = BSDF_ALL & ~BSDF_SPECULAR;
  Spec<NC> ld;
  RefPoint ref;
  ref.p = it.p; ref.p_error = it.p_error; ref.n = it.n; ref.time = it.time;
  const pt_light& L = d->lights[light_num];
  LightSample<NC> ls;
  if (light_sample_li<NC>(d, sc.world_radius, light_num, ref, u_light, &ls, sc.envs.data())) {
#ifdef ORC_DEBUG_PRINT
    fprintf(stderr, "[or] p=(%g %g %g) n=(%g %g %g) perr=(%g %g %g) light %d\n", it.p.x, it.p.y, it.p.z, it.n.x, it.n.y, it.n.z, it.p_error.x, it.p_error.y, it.p_error.z, light_num);
    fprintf(stderr, "[or] ls pdf %g li %g wi=(%g %g %g) sh_o=(%.9g %.9g %.9g) sh_d=(%.9g %.9g %.9g) occluded %d\n", ls.pdf, ls.li.c[0], ls.wi.x, ls.wi.y, ls.wi.z, ls.shadow.o.x, ls.shadow.o.y, ls.shadow.o.z, ls.shadow.d.x, ls.shadow.d.y, ls.shadow.d.z, (int)sc.intersect_p(ls.shadow, nullptr));
#endif
    if (ls.pdf > 0.0f && !ls.li.is_black()) {
      Spec<NC> f = bsdf.f(it.wo, ls.wi, flags) * abs_dot(ls.wi, it.sn);
      Float scattering_pdf = bsdf.pdf(it.wo, ls.wi, flags);
      if (!f.is_black()) {
        Spec<NC> li = ls.li;
        if (sc.intersect_p(ls.shadow, cnt)) li = Spec<NC>();
        if (!li.is_black()) {
          if (light_is_delta(L)) ld = ld + f * li / ls.pdf;
          else {
            Float weight = power_heuristic(1, ls.pdf, 1, scattering_pdf);
            ld = ld + f * li * (weight / ls.pdf);
          }
        }
      }
    }
  }
  if (!light_is_delta(L)) {
    Spec<NC> f;
    Vec3 wi;
    bool sampled_specular = false;
    Float scattering_pdf = 0.0f;
    Spec<NC> f2;
    Vec3 wi2;
    Float pdf2;
    int stype;
    if (bsdf.sample_f(it.wo, u_scattering, flags, &f2, &wi2, &pdf2, &stype)) {
      f = f2 * abs_dot(wi2, it.sn);
      wi = wi2;
      scattering_pdf = pdf2;
      sampled_specular = (stype & BSDF_SPECULAR) != 0;
    }
    if (!f.is_black() && scattering_pdf > 0.0f) {
      Float weight = 1.0f;
      if (!sampled_specular) {
        Float light_pdf = light_pdf_li(d, light_num, ref, wi, sc.envs.data());
        if (light_pdf == 0.0f) return ld;
        weight = power_heuristic(1, scattering_pdf, 1, light_pdf);
      }
      Ray ray = spawn_ray(ref, wi);
      Hit h;
      Spec<NC> li;
      if (sc.intersect(ray, &h, cnt)) {
        if (d->tri_light[h.tri] == light_num) {
          SurfInt ls2 = make_si(d, h);
          li = surface_le<NC>(d, ls2, -wi);
        }
      } else if (L.type == PT_LIGHT_INFINITE) li = env_le<NC>(sc.envs[L.tri], ray.d);  // light.le(ray): zero for the other lights
      if (!li.is_black()) ld = ld + f * li * Spec<NC>(1.0f) * (weight / scattering_pdf);
    }
  }
  return ld;
}

// uniform_sample_one_light_surface (sample_lights.rs:129-176)
template <int NC>
inline Spec<NC> sample_one_light(Scene<NC>& sc, const SurfInt& it, const BSDF<NC>& bsdf, Sampler& sampler, const Distribution1D* distrib, Counters* cnt) {
  size_t n_lights = sc.d->n_lights;
  if (n_lights == 0) return Spec<NC>();
  Float light_pdf;
  size_t light_num = distrib->sample_discrete(sampler.get_1d(), &light_pdf);
  if (light_pdf <= 0.0f) return Spec<NC>();
  Vec2 u_light = sampler.get_2d();
  Vec2 u_scattering = sampler.get_2d();
  return estimate_direct(sc, it, bsdf, u_scattering, (int)light_num, u_light, cnt) / light_pdf;
}

struct PathTrace {  // per-sample debug record for parity tests
  int first_tri = -1;
  Float first_t = 0;
  int bounces = 0;
};

// AOVIntegrator::li (integrators/aov.rs:101-181) for the supported targets; RGB only
template <int NC>
inline Spec<NC> aov_li(Scene<NC>& sc, const Ray& r, Counters* cnt, PathTrace* trace) {
  const pt_scene_desc* d = sc.d;
  Spec<NC> out;
  Ray ray = r;
  Hit h;
  if (!sc.intersect(ray, &h, cnt)) return out;
  SurfInt si = make_si(d, h);
  if (trace) { trace->first_tri = h.tri; trace->first_t = h.t; }
  const Float scale = d->integrator.aov_scale;
  Float rgb[3] = {0, 0, 0};
  auto v2c = [&](Vec3 v) { Vec3 c = 0.5f * v + Vec3(0.5f, 0.5f, 0.5f); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z; };
  switch (d->integrator.aov_target) {
    case PT_AOV_DISTANCE: { Float dist = ray.t_max / length(ray.d); rgb[0] = rgb[1] = rgb[2] = dist; break; }
    case PT_AOV_DEPTH: { Float dist = length(si.p - ray.o); rgb[0] = rgb[1] = rgb[2] = dist; break; }
    case PT_AOV_N: v2c(si.n); break;
    case PT_AOV_NS: v2c(si.sn); break;
    case PT_AOV_UV: {
      Vec2 uv[3];
      tri_uvs(d, h.tri, uv);
      rgb[0] = h.b0 * uv[0].x + h.b1 * uv[1].x + h.b2 * uv[2].x;  // triangle.rs:353-354: b0*uv[0] + b1*uv[1] + b2*uv[2]
      rgb[1] = h.b0 * uv[0].y + h.b1 * uv[1].y + h.b2 * uv[2].y;
      rgb[2] = 0.0f;
      break;
    }
    default: v2c(si.sdpdu); break;
  }
  for (int c = 0; c < 3 && c < NC; c++) {
    Float v = rgb[c];
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);  // Spectrum::clamp(0, 1)
    out.c[c] = v * scale;
  }
  return out;
}

// AOIntegrator::li (integrators/ao.rs:50-105): n_samples any-hit queries from the first surface that has a BSDF, directions
// from the sampler's 2D sample array in the frame of the TRUE geometry
template <int NC>
inline Spec<NC> ao_li(Scene<NC>& sc, const Ray& r, Sampler& sampler, Counters* cnt, PathTrace* trace) {
  const pt_scene_desc* d = sc.d;
  Spec<NC> l;
  Ray ray = r;
  const uint32_t n_samples = (uint32_t)d->integrator.ao_samples;  // round_count is the identity for halton / sobol (sampler.rs:41-43)
  const bool cos_sample = d->integrator.ao_cos_sample != 0;
  for (;;) {
    Hit h;
    if (!sc.intersect(ray, &h, cnt)) break;
    SurfInt isect = make_si(d, h);
    if (trace && trace->first_tri == -1) { trace->first_tri = h.tri; trace->first_t = h.t; }
    RefPoint rp; rp.p = isect.p; rp.p_error = isect.p_error; rp.n = isect.n; rp.time = isect.time;
    if (d->tri_material[isect.tri] < 0) {  // isect.bsdf.is_none(): continue through the surface (ao.rs:68-71)
      ray = spawn_ray(rp, ray.d);
      continue;
    }
    Vec3 n = face_forward(isect.n, ray.d * -1.0f);
    Vec3 s = normalize(isect.dpdu);
    Vec3 t = cross(isect.n, s);
    for (uint32_t i = 0; i < n_samples; i++) {
      Vec2 u = sampler.array_2d_elem_of(0, n_samples, i);
      Vec3 wl;
      Float pdf;
      if (cos_sample) { wl = cosine_sample_hemisphere(u); pdf = fabsf(wl.z) * INV_PI; }
      else { wl = uniform_sample_hemisphere(u); pdf = INV_2_PI; }
      Vec3 wi(wl.x * s.x + wl.y * t.x + wl.z * n.x, wl.x * s.y + wl.y * t.y + wl.z * n.y, wl.x * s.z + wl.y * t.z + wl.z * n.z);
      if (!sc.intersect_p(spawn_ray(rp, wi), cnt)) l = l + Spec<NC>(dot(wi, n) / (pdf * (Float)n_samples));
    }
    break;
  }
  return l;
}

// uniform_sample_all_lights (sample_lights.rs:24-78): n_samples per light from two 2D sample arrays (which hold the SAME
// values: every 2D array is filled from dimensions (5, 6), halton.rs:195-206), single get_2d pairs once the requested
// arrays are used up
template <int NC>
inline Spec<NC> sample_all_lights(Scene<NC>& sc, const SurfInt& it, const BSDF<NC>& bsdf, Sampler& sampler, Counters* cnt) {
  Spec<NC> l;
  for (uint32_t j = 0; j < sc.n_sampled_lights(); j++) {
    const uint32_t n = (uint32_t)sc.d->lights[j].n_samples;
    const bool a = sampler.next_2d_array(), b = sampler.next_2d_array();
    if (a && b) {
      Spec<NC> ld;
      const uint32_t ka = sampler.array2d_offset - 2, kb = sampler.array2d_offset - 1;  // u_light_array, u_scattering_array
      for (uint32_t k = 0; k < n; k++) {
        Vec2 u_light = sampler.array_2d_elem_of(ka, n, k), u_scattering = sampler.array_2d_elem_of(kb, n, k);
        ld = ld + estimate_direct(sc, it, bsdf, u_scattering, (int)j, u_light, cnt);
      }
      l = l + ld / (Float)n;
    } else {
      Vec2 u_light = sampler.get_2d();
      Vec2 u_scattering = sampler.get_2d();
      l = l + estimate_direct(sc, it, bsdf, u_scattering, (int)j, u_light, cnt);
    }
  }
  return l;
}
// uniform_sample_one_light without a distribution (sample_lights.rs:80-128)
template <int NC>
inline Spec<NC> sample_one_light_uniform(Scene<NC>& sc, const SurfInt& it, const BSDF<NC>& bsdf, Sampler& sampler, Counters* cnt) {
  const size_t n_lights = sc.n_sampled_lights();
  if (n_lights == 0) return Spec<NC>();
  size_t light_num = std::min((size_t)(sampler.get_1d() * (Float)n_lights), n_lights - 1);
  Float light_pdf = 1.0f / (Float)n_lights;
  Vec2 u_light = sampler.get_2d();
  Vec2 u_scattering = sampler.get_2d();
  return estimate_direct(sc, it, bsdf, u_scattering, (int)light_num, u_light, cnt) / light_pdf;
}

template <int NC>
inline Spec<NC> recursive_li(Scene<NC>& sc, const Ray& ray, Sampler& sampler, Counters* cnt, PathTrace* trace, int depth);
// SamplerIntegrator::specular_reflect / specular_transmit (core/integrator/sampler.rs:37-160) without the differentials
template <int NC>
inline Spec<NC> specular_bounce(Scene<NC>& sc, const Ray& ray, const SurfInt& isect, const BSDF<NC>& bsdf, Sampler& sampler, Counters* cnt, int depth, int type) {
  Vec3 wo = isect.wo;
  Vec2 u = sampler.get_2d();
  Spec<NC> f;
  Vec3 wi;
  Float pdf;
  int stype;
  if (!bsdf.sample_f(wo, u, type | BSDF_SPECULAR, &f, &wi, &pdf, &stype)) return Spec<NC>();
  Float wi_ns = abs_dot(wi, isect.sn);
  if (pdf > 0.0f && !f.is_black() && wi_ns != 0.0f) {
    RefPoint rp; rp.p = isect.p; rp.p_error = isect.p_error; rp.n = isect.n; rp.time = isect.time;
    return f * recursive_li(sc, spawn_ray(rp, wi), sampler, cnt, nullptr, depth + 1) * (wi_ns / pdf);
  }
  return Spec<NC>();
}
// DirectLightingIntegrator::li (directlighting.rs:66-125) and WhittedIntegrator::li (whitted.rs:37-98)
template <int NC>
inline Spec<NC> recursive_li(Scene<NC>& sc, const Ray& ray, Sampler& sampler, Counters* cnt, PathTrace* trace, int depth) {
  const pt_scene_desc* d = sc.d;
  const bool whitted = d->integrator.type == PT_INTEGRATOR_WHITTED;
  Hit h;
  if (!sc.intersect(ray, &h, cnt)) {  // sum of light.le(ray) over ALL lights (directlighting.rs:112-122): non-zero for infinite lights
    Spec<NC> l;
    for (int li : sc.infinite_lights) l = l + env_le<NC>(sc.envs[sc.d->lights[li].tri], ray.d);
    return l;
  }
  SurfInt isect = make_si(d, h);
  if (trace && trace->first_tri == -1) { trace->first_tri = h.tri; trace->first_t = h.t; }
  BSDF<NC> bsdf;
  compute_scattering(d, isect, &bsdf, false);
  if (!bsdf.valid) {
    if (whitted) return Spec<NC>();  // whitted.rs:84
    RefPoint rp; rp.p = isect.p; rp.p_error = isect.p_error; rp.n = isect.n; rp.time = isect.time;
    return recursive_li(sc, spawn_ray(rp, ray.d), sampler, cnt, trace, depth);  // directlighting.rs:108-110
  }
  if (cnt) cnt->vertices++;
  Spec<NC> l;
  if (whitted) {  // whitted.rs:61-79: one sample of every light, shading normal, no MIS, no emission
    RefPoint ref; ref.p = isect.p; ref.p_error = isect.p_error; ref.n = isect.n; ref.time = isect.time;
    for (uint32_t j = 0; j < sc.n_sampled_lights(); j++) {
      Vec2 u = sampler.get_2d();
      LightSample<NC> ls;
      if (!light_sample_li<NC>(d, sc.world_radius, (int)j, ref, u, &ls, sc.envs.data())) { l = l + Spec<NC>(); continue; }
      if (ls.pdf <= 0.0f || ls.li.is_black()) { l = l + Spec<NC>(); continue; }
      Spec<NC> f = bsdf.f(isect.wo, ls.wi, BSDF_ALL);
      if (!f.is_black() && !sc.intersect_p(ls.shadow, cnt)) l = l + f * ls.li * (abs_dot(ls.wi, isect.sn) / ls.pdf);
      else l = l + Spec<NC>();
    }
  } else {
    l = surface_le<NC>(d, isect, isect.wo);
    if (sc.n_sampled_lights() > 0) {
      if (d->integrator.direct_strategy == PT_DIRECT_ALL) l = l + sample_all_lights(sc, isect, bsdf, sampler, cnt);
      else l = l + sample_one_light_uniform(sc, isect, bsdf, sampler, cnt);
    }
  }
  if (depth + 1 < d->integrator.max_depth) {
    l = l + specular_bounce(sc, ray, isect, bsdf, sampler, cnt, depth, BSDF_REFLECTION);
    l = l + specular_bounce(sc, ray, isect, bsdf, sampler, cnt, depth, BSDF_TRANSMISSION);
  }
  return l;
}

template <int NC>
inline Spec<NC> path_li(Scene<NC>& sc, const Ray& r, Sampler& sampler, Counters* cnt, PathTrace* trace);
// SamplerIntegrator::li of the scene's integrator
template <int NC>
inline Spec<NC> integrator_li(Scene<NC>& sc, const Ray& r, Sampler& sampler, Counters* cnt, PathTrace* trace) {
  if (sc.d->integrator.type == PT_INTEGRATOR_AOV) return aov_li(sc, r, cnt, trace);
  if (sc.d->integrator.type == PT_INTEGRATOR_AO) return ao_li(sc, r, sampler, cnt, trace);
  if (sc.d->integrator.type == PT_INTEGRATOR_DIRECT || sc.d->integrator.type == PT_INTEGRATOR_WHITTED) return recursive_li(sc, r, sampler, cnt, trace, 0);
  return path_li(sc, r, sampler, cnt, trace);
}

// PathIntegrator::li (path.rs:61-241)
template <int NC>
inline Spec<NC> path_li(Scene<NC>& sc, const Ray& r, Sampler& sampler, Counters* cnt, PathTrace* trace) {
  const pt_scene_desc* d = sc.d;
  if (d->n_lights == 0) return Spec<NC>();  // light_distribution is None
  int max_depth = d->integrator.max_depth;
  Ray ray = r;
  Float eta_scale = 1.0f;
  Spec<NC> l, beta(1.0f);
  bool specular_bounce = false;
  int bounces = 0;
  for (;;) {
    Hit h;
    bool found = sc.intersect(ray, &h, cnt);
    SurfInt isect;
    if (found) isect = make_si(d, h);
    if (trace && bounces == 0 && trace->first_tri == -1 && found) { trace->first_tri = h.tri; trace->first_t = h.t; }
    if (bounces == 0 || specular_bounce) {
      if (found) l = l + beta * surface_le<NC>(d, isect, -ray.d);
      else for (int li : sc.infinite_lights) l = l + beta * env_le<NC>(sc.envs[d->lights[li].tri], ray.d);  // path.rs:92-95
    }
    if (!found || bounces >= max_depth) break;
    BSDF<NC> bsdf;
    compute_scattering(d, isect, &bsdf);
    if (!bsdf.valid) {
      RefPoint rp; rp.p = isect.p; rp.p_error = isect.p_error; rp.n = isect.n; rp.time = isect.time;
      ray = spawn_ray(rp, ray.d);
      continue;
    }
    if (cnt) cnt->vertices++;
    const Distribution1D* distrib = sc.lookup(isect.p);
    if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
      Spec<NC> ld = beta * sample_one_light(sc, isect, bsdf, sampler, distrib, cnt);
      l = l + ld;
    }
    Vec3 wo = -ray.d;
    Spec<NC> f;
    Vec3 wi;
    Float pdf;
    int flags;
    if (!bsdf.sample_f(wo, sampler.get_2d(), BSDF_ALL, &f, &wi, &pdf, &flags)) break;
    if (f.is_black() || pdf == 0.0f) break;
    beta = beta * (f * (abs_dot(wi, isect.sn) / pdf));
    specular_bounce = (flags & BSDF_SPECULAR) != 0;
    if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
      Float eta = bsdf.eta;
      eta_scale *= (dot(wo, isect.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta);
    }
    RefPoint rp; rp.p = isect.p; rp.p_error = isect.p_error; rp.n = isect.n; rp.time = isect.time;
    ray = spawn_ray(rp, wi);
    Spec<NC> rr_beta = beta * eta_scale;
    if (rr_beta.max_component() < d->integrator.rr_threshold && bounces > 3) {
      Float q = rmax(0.05f, 1.0f - rr_beta.max_component());
      if (sampler.get_1d() < q) break;
      beta = beta / (1.0f - q);
    }
    bounces += 1;
  }
  if (trace) trace->bounces = bounces;
  return l;
}

// validate_radiance_result (sampler.rs:151-176)
template <int NC>
inline Spec<NC> validate_radiance(const pt_scene_desc* d, const Spec<NC>& l) {
  if (!l.is_valid()) return Spec<NC>();
  Float y = spec_y(d, l);
  if (y < -1e-5f) return Spec<NC>();
  if (std::isinf(y)) return Spec<NC>();
  return l;
}

// ---- Film -------------------------------------------------------------------------------------------
template <int NC>
struct TilePixel { Spec<NC> contrib; Float wsum = 0; };

template <int NC>
struct FilmTile {
  int b[4];  // pixel bounds
  std::vector<TilePixel<NC>> px;
  const pt_film* F;
  const pt_scene_desc* d;
  FilmTile(const pt_scene_desc* desc, const int tile[4]) : F(&desc->film), d(desc) {  // Film::get_film_tile (film.rs:203-217)
    Float rx = F->filter_radius[0], ry = F->filter_radius[1];
    int p0x = f2i(floorf((Float)tile[0] - rx)), p0y = f2i(floorf((Float)tile[1] - ry));
    int p1x = f2i(ceilf((Float)tile[2] + rx)), p1y = f2i(ceilf((Float)tile[3] + ry));
    const int* c = F->cropped_pixel_bounds;
    b[0] = std::max(p0x, c[0]); b[1] = std::max(p0y, c[1]); b[2] = std::min(p1x, c[2]); b[3] = std::min(p1y, c[3]);
    int w = std::max(b[2] - b[0], 0), hh = std::max(b[3] - b[1], 0);
    px.resize((size_t)w * hh);
  }
  void add_sample(Vec2 p_film, Spec<NC> l, Float sample_weight) {  // film_tile.rs:84-183
    Float y = spec_y(d, l);
    if (y > F->max_sample_luminance) l = l * (F->max_sample_luminance / y);
    Float rx = F->filter_radius[0], ry = F->filter_radius[1];
    Float irx = 1.0f / rx, iry = 1.0f / ry;
    int p0x = f2i(floorf(p_film.x - rx)), p0y = f2i(floorf(p_film.y - ry));
    int p1x = f2i(ceilf(p_film.x + rx)), p1y = f2i(ceilf(p_film.y + ry));
    p0x = std::max(p0x, b[0]); p0y = std::max(p0y, b[1]);
    p1x = std::min(p1x, b[2]); p1y = std::min(p1y, b[3]);
    int dx = p1x - p0x, dy = p1y - p0y;
    if (dx <= 0 || dy <= 0) return;
    std::vector<int> ifx((size_t)dx), ify((size_t)dy);
    Float lx = irx * 15.0f, ly = iry * 15.0f;
    for (int x = p0x; x < p1x; x++) {
      Float dd = fabsf((Float)x + 0.5f - p_film.x);
      ifx[(size_t)(x - p0x)] = dd <= rx ? std::min(f2i(floorf(dd * lx)), 15) : -1;
    }
    for (int yy = p0y; yy < p1y; yy++) {
      Float dd = fabsf((Float)yy + 0.5f - p_film.y);
      ify[(size_t)(yy - p0y)] = dd <= ry ? std::min(f2i(floorf(dd * ly)), 15) : -1;
    }
    std::vector<Float> w((size_t)dx * dy, 0.0f);
    for (int yy = 0; yy < dy; yy++) {
      if (ify[(size_t)yy] < 0) continue;
      for (int x = 0; x < dx; x++) {
        if (ifx[(size_t)x] < 0) continue;
        w[(size_t)yy * dx + x] = F->filter_table[ify[(size_t)yy] * 16 + ifx[(size_t)x]];
      }
    }
    Float sum = 0.0f;
    for (Float v : w) sum += v;
    if (sum <= 0.0f) return;
    Float isum = 1.0f / sum;
    for (Float& v : w) v *= isum;
    int width = b[2] - b[0];
    for (int yy = 0; yy < dy; yy++)
      for (int x = 0; x < dx; x++) {
        Float fw = w[(size_t)yy * dx + x];
        size_t pi = (size_t)(p0y + yy - b[1]) * width + (size_t)(p0x + x - b[0]);
        px[pi].contrib = px[pi].contrib + l * sample_weight * fw;
        px[pi].wsum += fw;
      }
  }
};

struct FilmPixel { Float xyz[3] = {0, 0, 0}; Float wsum = 0; };

struct Film {
  const pt_film* F;
  std::vector<FilmPixel> px;
  std::mutex m;
  explicit Film(const pt_film* f) : F(f) {
    const int* c = F->cropped_pixel_bounds;
    px.resize((size_t)std::max(c[2] - c[0], 0) * std::max(c[3] - c[1], 0));
  }
  template <int NC>
  void merge(const FilmTile<NC>& t) {  // film.rs:219-241
    std::lock_guard<std::mutex> g(m);
    const int* c = F->cropped_pixel_bounds;
    int w = c[2] - c[0], tw = t.b[2] - t.b[0];
    for (int y = t.b[1]; y < t.b[3]; y++)
      for (int x = t.b[0]; x < t.b[2]; x++) {
        const TilePixel<NC>& s = t.px[(size_t)(y - t.b[1]) * tw + (x - t.b[0])];
        FilmPixel& dpx = px[(size_t)(y - c[1]) * w + (x - c[0])];
        Float xyz[3];
        spec_to_xyz(t.d, s.contrib, xyz);
        for (int i = 0; i < 3; i++) dpx.xyz[i] += xyz[i];
        dpx.wsum += s.wsum;
      }
  }
  void write_rgb(float* out) const {  // film.rs:440-472 (no splats)
    for (size_t i = 0; i < px.size(); i++) {
      const Float* xyz = px[i].xyz;
      Float c[3];
      c[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
      c[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
      c[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
      if (px[i].wsum > 0.0f) {
        Float inv = 1.0f / px[i].wsum;
        for (int k = 0; k < 3; k++) c[k] = rmax(0.0f, c[k] * inv);
      }
      // splat term: xyz_to_rgb(0) = 0 ; c += 0
      for (int k = 0; k < 3; k++) out[3 * i + k] = (c[k] + 0.0f) * F->scale;
    }
  }
};

// SampleIntegratorCore::render / render_tile (sampler.rs:201-325). spp range [s0, s1) of every pixel; 0,0 = all.
template <int NC>
// shard_count > 1: only the 16x16 render tiles whose number is congruent to shard_index (the product's tile sharding,
// used by bench.py to time a bounded sample of a large workload).
inline void render(Scene<NC>& sc, int n_threads, int s0, int s1, float* out_rgb, float* out_xyzw, Counters* total, int shard_index = 0, int shard_count = 1) {
  const pt_scene_desc* d = sc.d;
  const int* sb = d->film.sample_bounds;
  const int TILE = 16;
  int ntx = (sb[2] - sb[0] + TILE - 1) / TILE, nty = (sb[3] - sb[1] + TILE - 1) / TILE;
  Film film(&d->film);
  std::atomic<int> next(0);
  std::mutex cm;
  int spp = sc.scfg.spp;
  if (s0 == 0 && s1 == 0) s1 = spp;
  auto worker = [&]() {
    Counters cnt;
    for (;;) {
      int ti = next.fetch_add(1);
      if (ti >= ntx * nty) break;
      if (shard_count > 1 && ti % shard_count != shard_index) continue;
      int tx = ti % ntx, ty = ti / ntx;
      int tile[4] = {sb[0] + tx * TILE, sb[1] + ty * TILE, std::min(sb[0] + tx * TILE + TILE, sb[2]), std::min(sb[1] + ty * TILE + TILE, sb[3])};
      Sampler sampler(&sc.scfg, (uint32_t)ti);
      FilmTile<NC> ft(d, tile);
      for (int y = tile[1]; y < tile[3]; y++)
        for (int x = tile[0]; x < tile[2]; x++) {
          sampler.start_pixel(x, y);
          const bool indexed = d->sampler.type < PT_SAMPLER_RANDOM;  // halton / sobol: sample s is a pure function of (pixel, s)
          for (int s = indexed ? std::max(s0, 0) : 0; s < (indexed ? std::min(s1, spp) : spp); s++) {
            if (s > 0) sampler.set_sample_number(s);
            CameraSample cs;
            Vec2 u = sampler.get_2d();
            cs.p_film = Vec2((Float)x, (Float)y) + u;
            cs.p_lens = sampler.get_2d();
            cs.time = sampler.get_1d();
            Ray ray = camera_ray(d->camera, d->film, cs);
            cnt.camera_rays++;
            Spec<NC> l = integrator_li(sc, ray, sampler, &cnt, nullptr);
            l = validate_radiance(d, l);
            if (s >= s0 && s < s1) ft.add_sample(cs.p_film, l, 1.0f);
          }
        }
      film.merge(ft);
    }
    std::lock_guard<std::mutex> g(cm);
    if (total) {
      total->nodes += cnt.nodes; total->tris += cnt.tris; total->regular_rays += cnt.regular_rays;
      total->shadow_rays += cnt.shadow_rays; total->camera_rays += cnt.camera_rays; total->vertices += cnt.vertices;
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(n_threads, 1); i++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  if (out_rgb) film.write_rgb(out_rgb);
  if (out_xyzw)
    for (size_t i = 0; i < film.px.size(); i++) {
      for (int k = 0; k < 3; k++) out_xyzw[4 * i + k] = film.px[i].xyz[k];
      out_xyzw[4 * i + 3] = film.px[i].wsum;
    }
}

}  // namespace orc
