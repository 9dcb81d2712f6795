// ORACLE — TEST INFRASTRUCTURE ONLY (see o_math.h header).
// o_shading.h — spectra, surface interaction, BSDF/BxDFs, microfacet distribution, materials, lights,
// light-selection distributions.
// Follows reference src/core/spectrum/{rgb,sampled}.rs, src/core/interaction/*.rs, src/shapes/triangle.rs:132-186,
// 354-450,590-651, src/core/reflection/{bsdf,bxdf,math,lambertian,oren_nayar,specular,fresnel,microfacet}.rs,
// src/core/distribution/{microfacet,trowbridge_reitz}.rs, src/materials/{matte,plastic,glass,metal,mirror}.rs,
// src/lights/{diffuse,point,distant}.rs, src/core/shape/shape.rs:40-54, src/core/lightdistrib/*.rs.
#pragma once
#include "o_accel.h"
#include "o_sampling.h"

namespace orc {

// ---- Spectrum (component-wise f32 ops) --------------------------------------------------------------
template <int NC>
struct Spec {
  Float c[NC];
  Spec() { for (int i = 0; i < NC; i++) c[i] = 0.0f; }
  explicit Spec(Float v) { for (int i = 0; i < NC; i++) c[i] = v; }
  static Spec load(const float* p) { Spec s; for (int i = 0; i < NC; i++) s.c[i] = p[i]; return s; }
  bool is_black() const {
    if (NC == 3) { for (int i = 0; i < NC; i++) if (c[i] != 0.0f) return false; return true; }  // rgb.rs:105-113
    for (int i = 0; i < NC; i++) if (!(c[i] <= 0.0f)) return false;                               // sampled.rs:126-129
    return true;
  }
  bool is_valid() const { for (int i = 0; i < NC; i++) if (!std::isfinite(c[i])) return false; return true; }
  Float max_component() const { Float m = c[0]; for (int i = 1; i < NC; i++) m = rmax(m, c[i]); return m; }
  Spec clamp_zero() const { Spec r; for (int i = 0; i < NC; i++) r.c[i] = rclamp(c[i], 0.0f, INF); return r; }
  Spec sqrt_() const { Spec r; for (int i = 0; i < NC; i++) r.c[i] = sqrtf(c[i]); return r; }
};
template <int NC> inline Spec<NC> operator+(Spec<NC> a, const Spec<NC>& b) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] + b.c[i]; return a; }
template <int NC> inline Spec<NC> operator-(Spec<NC> a, const Spec<NC>& b) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] - b.c[i]; return a; }
template <int NC> inline Spec<NC> operator*(Spec<NC> a, const Spec<NC>& b) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] * b.c[i]; return a; }
template <int NC> inline Spec<NC> operator/(Spec<NC> a, const Spec<NC>& b) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] / b.c[i]; return a; }
template <int NC> inline Spec<NC> operator*(Spec<NC> a, Float s) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] * s; return a; }
template <int NC> inline Spec<NC> operator*(Float s, Spec<NC> a) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] * s; return a; }
template <int NC> inline Spec<NC> operator/(Spec<NC> a, Float s) { for (int i = 0; i < NC; i++) a.c[i] = a.c[i] / s; return a; }

// Spectrum::to_xyz / y through the weights of the scene description (rgb.rs:55-63, sampled.rs:53-79)
template <int NC>
inline void spec_to_xyz(const pt_scene_desc* d, const Spec<NC>& s, Float xyz[3]) {
  for (int r = 0; r < 3; r++) {
    const float* w = d->xyz_weights + r * 60;
    Float acc = w[0] * s.c[0];
    for (int i = 1; i < NC; i++) acc += w[i] * s.c[i];
    xyz[r] = acc * d->xyz_scale;
  }
}
template <int NC>
inline Float spec_y(const pt_scene_desc* d, const Spec<NC>& s) {
  const float* w = d->xyz_weights + 60;
  Float acc = w[0] * s.c[0];
  for (int i = 1; i < NC; i++) acc += w[i] * s.c[i];
  return acc * d->xyz_scale;
}

// ---- BSDF inline functions (reflection/math.rs) --------------------------------------------------------
inline Float cos_theta(Vec3 w) { return w.z; }
inline Float cos2_theta(Vec3 w) { return w.z * w.z; }
inline Float abs_cos_theta(Vec3 w) { return fabsf(w.z); }
inline Float sin2_theta(Vec3 w) { return rmax(0.0f, 1.0f - cos2_theta(w)); }
inline Float sin_theta(Vec3 w) { return sqrtf(sin2_theta(w)); }
inline Float tan_theta(Vec3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan2_theta(Vec3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(Vec3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : rclamp(w.x / s, -1.0f, 1.0f); }
inline Float sin_phi(Vec3 w) { Float s = sin_theta(w); return s == 0.0f ? 0.0f : rclamp(w.y / s, -1.0f, 1.0f); }
inline Float cos2_phi(Vec3 w) { return cos_phi(w) * cos_phi(w); }
inline Float sin2_phi(Vec3 w) { return sin_phi(w) * sin_phi(w); }
inline Vec3 reflect(Vec3 wo, Vec3 n) { Vec3 a = (2.0f * dot(wo, n)) * n; Vec3 b = -wo; return a + b; }
inline bool refract(Vec3 wi, Vec3 n, Float eta, Vec3* wt) {
  Float cos_i = dot(n, wi);
  Float sin2_i = rmax(0.0f, 1.0f - cos_i * cos_i);
  Float sin2_t = eta * eta * sin2_i;
  if (sin2_t >= 1.0f) return false;
  Float cos_t = sqrtf(1.0f - sin2_t);
  Vec3 a = eta * (-wi);
  Vec3 b = (eta * cos_i - cos_t) * n;
  *wt = a + b;
  return true;
}
inline bool same_hemisphere(Vec3 w, Vec3 wp) { return w.z * wp.z > 0.0f; }
inline Vec3 spherical_direction(Float sin_t, Float cos_t, Float phi) { return Vec3(sin_t * cosf(phi), sin_t * sinf(phi), cos_t); }

enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

// ---- Fresnel (reflection/fresnel.rs:16-82) -------------------------------------------------------------
inline Float fr_dielectric(Float cos_i, Float eta_i, Float eta_t) {
  cos_i = rclamp(cos_i, -1.0f, 1.0f);
  bool entering = cos_i > 0.0f;
  if (!entering) { std::swap(eta_i, eta_t); cos_i = fabsf(cos_i); }
  Float sin_i = sqrtf(rmax(0.0f, 1.0f - cos_i * cos_i));
  Float sin_t = eta_i / eta_t * sin_i;
  if (sin_t >= 1.0f) return 1.0f;
  Float cos_t = sqrtf(rmax(0.0f, 1.0f - sin_t * sin_t));
  Float rparl = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
  Float rperp = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
  return (rparl * rparl + rperp * rperp) / 2.0f;
}
template <int NC>
inline Spec<NC> fr_conductor(Float cos_i, const Spec<NC>& eta_i, const Spec<NC>& eta_t, const Spec<NC>& k) {
  cos_i = rclamp(cos_i, -1.0f, 1.0f);
  Spec<NC> eta = eta_t / eta_i;
  Spec<NC> etak = k / eta_i;
  Float c2 = cos_i * cos_i;
  Float s2 = 1.0f - c2;
  Float s22 = s2 * s2;
  Spec<NC> eta2 = eta * eta, etak2 = etak * etak;
  Spec<NC> C2(c2), S2(s2), S22(s22);
  Spec<NC> t0 = eta2 - etak2 - S2;
  Spec<NC> a2plusb2 = (t0 * t0 + eta2 * etak2 * 4.0f).sqrt_();
  Spec<NC> t1 = a2plusb2 + C2;
  Spec<NC> a = ((a2plusb2 + t0) * 0.5f).sqrt_();
  Spec<NC> t2 = Spec<NC>(cos_i) * a * 2.0f;
  Spec<NC> rs = (t1 - t2) / (t1 + t2);
  Spec<NC> t3 = C2 * a2plusb2 + S22;
  Spec<NC> t4 = t2 * S2;
  Spec<NC> rp = rs * (t3 - t4) / (t3 + t4);
  return (rp + rs) * 0.5f;
}
enum FresnelKind { FR_NOOP, FR_DIELECTRIC, FR_CONDUCTOR };
template <int NC>
struct Fresnel {
  int kind = FR_NOOP;
  Float eta_i = 1, eta_t = 1;
  Spec<NC> c_eta_i, c_eta_t, c_k;
  Spec<NC> evaluate(Float cos_i) const {
    if (kind == FR_NOOP) return Spec<NC>(1.0f);
    if (kind == FR_DIELECTRIC) return Spec<NC>(fr_dielectric(cos_i, eta_i, eta_t));
    return fr_conductor(fabsf(cos_i), c_eta_i, c_eta_t, c_k);
  }
};

// ---- TrowbridgeReitz (distribution/trowbridge_reitz.rs) ------------------------------------------------
inline void tr_sample_11(Float cos_t, Float u1, Float u2, Float* sx, Float* sy) {
  if (cos_t > 0.9999f) {
    Float r = sqrtf(u1 / (1.0f - u1));
    Float phi = 2.0f * PI * u2;
    *sx = r * cosf(phi);
    *sy = r * sinf(phi);
    return;
  }
  Float sin_t = sqrtf(rmax(0.0f, 1.0f - cos_t * cos_t));
  Float tan_t = sin_t / cos_t;
  Float a = 1.0f / tan_t;
  Float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
  a = 2.0f * u1 / g1 - 1.0f;
  Float tmp = rmin(1e10f, 1.0f / (a * a - 1.0f));
  Float b = tan_t;
  Float d = sqrtf(rmax(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
  Float s1 = b * tmp - d, s2 = b * tmp + d;
  Float slope_x = (a < 0.0f || s2 > 1.0f / tan_t) ? s1 : s2;
  Float s;
  if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
  else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
  Float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *sx = slope_x;
  *sy = s * z * sqrtf(1.0f + slope_x * slope_x);
}
inline Vec3 tr_sample(Vec3 wi, Float ax, Float ay, Float u1, Float u2) {
  Vec3 ws = normalize(Vec3(ax * wi.x, ay * wi.y, wi.z));
  Float sx, sy;
  tr_sample_11(cos_theta(ws), u1, u2, &sx, &sy);
  Float tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
  sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
  sx = tmp;
  sx *= ax;
  sy *= ay;
  return normalize(Vec3(-sx, -sy, 1.0f));
}
inline Float roughness_to_alpha(Float roughness) {
  roughness = rmax(roughness, 1e-3f);
  Float x = logf(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
struct TRDist {
  Float ax, ay;
  bool samplevis;
  TRDist() : ax(0.001f), ay(0.001f), samplevis(true) {}
  TRDist(Float x, Float y, bool sv) : ax(rmax(0.001f, x)), ay(rmax(0.001f, y)), samplevis(sv) {}
  Float d(Vec3 wh) const {
    Float t2 = tan2_theta(wh);
    if (std::isinf(t2)) return 0.0f;
    Float c2 = cos2_theta(wh);
    Float c4 = c2 * c2;
    Float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * t2;
    Float e2 = (1.0f + e) * (1.0f + e);
    return 1.0f / (PI * ax * ay * c4 * e2);
  }
  Float lambda(Vec3 w) const {
    Float att = fabsf(tan_theta(w));
    if (std::isinf(att)) return 0.0f;
    Float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    Float a2t2 = (alpha * att) * (alpha * att);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
  }
  Float g1(Vec3 w) const { return 1.0f / (1.0f + lambda(w)); }
  Float g(Vec3 wo, Vec3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
  Vec3 sample_wh(Vec3 wo, Vec2 u) const {
    if (!samplevis) {
      Float u1 = rclamp(u.x, 1e-6f, 1.0f - 1e-6f), u2 = rclamp(u.y, 1e-6f, 1.0f - 1e-6f);
      Float sin_t, cos_t, phi;
      if (ax == ay) {
        Float tan2 = ax * ax * u1 / (1.0f - u1);
        phi = 2.0f * PI * u2;
        cos_t = 1.0f / sqrtf(1.0f + tan2);
        sin_t = sqrtf(rmax(0.0f, 1.0f - cos_t * cos_t));
      } else {
        phi = atanf(ay / ax * tanf(2.0f * PI * u2 + 0.5f * PI));
        if (u2 > 0.5f) phi += PI;
        Float sp = sinf(phi), cp = cosf(phi);
        Float alpha2 = 1.0f / (cp * cp / (ax * ax) + sp * sp / (ay * ay));
        Float tan2 = alpha2 * u1 / (1.0f - u1);
        cos_t = 1.0f / sqrtf(1.0f + tan2);
        sin_t = sqrtf(rmax(0.0f, 1.0f - cos_t * cos_t));
      }
      Vec3 wh = normalize(spherical_direction(sin_t, cos_t, phi));
      if (!same_hemisphere(wo, wh)) wh = -wh;
      return wh;
    }
    bool flip = wo.z < 0.0f;
    Vec3 w = flip ? -wo : wo;
    Vec3 wh = tr_sample(w, ax, ay, u.x, u.y);
    if (flip) wh = -wh;
    return wh;
  }
  Float pdf(Vec3 wo, Vec3 wh) const {
    if (samplevis) return d(wh) * g1(wo) * abs_dot(wo, wh) / abs_cos_theta(wo);
    return d(wh) * abs_cos_theta(wh);
  }
};

// ---- BxDFs ------------------------------------------------------------------------------------------
enum BxKind { BX_LAMBERT, BX_OREN_NAYAR, BX_SPEC_REFL, BX_SPEC_TRANS, BX_FRESNEL_SPEC, BX_MF_REFL, BX_MF_TRANS,
              BX_LAMBERT_T /* lambertian.rs:59-99 */, BX_FRESNEL_BLEND /* fresnel_blend.rs: r = Rd, t = Rs */ };
inline Float pow5(Float v) { return powf(v, 5.0f); }  // fresnel_blend.rs:13-15
template <int NC>
struct BxDF {
  int kind;
  Spec<NC> r, t;       // reflectance / transmittance
  Float a = 0, b = 0;  // OrenNayar
  Float eta_a = 1, eta_b = 1;
  Fresnel<NC> fresnel;
  TRDist dist;

  int type() const {
    switch (kind) {
      case BX_LAMBERT: case BX_OREN_NAYAR: return BSDF_REFLECTION | BSDF_DIFFUSE;
      case BX_SPEC_REFL: return BSDF_REFLECTION | BSDF_SPECULAR;
      case BX_SPEC_TRANS: return BSDF_TRANSMISSION | BSDF_SPECULAR;
      case BX_FRESNEL_SPEC: return BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
      case BX_MF_REFL: return BSDF_REFLECTION | BSDF_GLOSSY;
      case BX_LAMBERT_T: return BSDF_TRANSMISSION | BSDF_DIFFUSE;
      case BX_FRESNEL_BLEND: return BSDF_REFLECTION | BSDF_GLOSSY;
      default: return BSDF_TRANSMISSION | BSDF_GLOSSY;
    }
  }
  bool matches(int flags) const { int tp = type(); return (tp & flags) == tp; }

  // ScaledBxDF (reflection/scaled.rs:7-62, used by MixMaterial): f and sample_f times `scale`; pdf and type unchanged
  bool scaled = false;
  Spec<NC> scale;
  Spec<NC> f(Vec3 wo, Vec3 wi) const { Spec<NC> v = f_raw(wo, wi); return scaled ? scale * v : v; }
  bool sample_f(Vec3 wo, Vec2 u, Spec<NC>* fo, Vec3* wio, Float* pdfo, int* stype) const {
    if (!sample_raw(wo, u, fo, wio, pdfo, stype)) return false;
    if (scaled) *fo = scale * *fo;
    return true;
  }

  Spec<NC> f_raw(Vec3 wo, Vec3 wi) const {
    switch (kind) {
      case BX_LAMBERT: return r * INV_PI;
      case BX_OREN_NAYAR: {
        Float sti = sin_theta(wi), sto = sin_theta(wo);
        Float max_cos = 0.0f;
        if (sti > 1e-4f && sto > 1e-4f) {
          Float d_cos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
          max_cos = rmax(d_cos, 0.0f);
        }
        Float sin_alpha, tan_beta;
        if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sto; tan_beta = sti / abs_cos_theta(wi); }
        else { sin_alpha = sti; tan_beta = sto / abs_cos_theta(wo); }
        return r * INV_PI * (a + b * max_cos * sin_alpha * tan_beta);
      }
      case BX_MF_REFL: {
        Float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
        Vec3 wh = wi + wo;
        if (cti == 0.0f || cto == 0.0f) return Spec<NC>();
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spec<NC>();
        wh = normalize(wh);
        Spec<NC> F = fresnel.evaluate(dot(wi, face_forward(wh, Vec3(0, 0, 1))));
        return r * F * (dist.d(wh) * dist.g(wo, wi) / (4.0f * cti * cto));
      }
      case BX_MF_TRANS: {
        if (same_hemisphere(wo, wi)) return Spec<NC>();
        Float cto = cos_theta(wo), cti = cos_theta(wi);
        if (cti == 0.0f || cto == 0.0f) return Spec<NC>();
        Float eta = cto > 0.0f ? eta_b / eta_a : eta_a / eta_b;
        Vec3 wh = normalize(wo + (wi * eta));
        if (wh.z < 0.0f) wh = -wh;
        Float wo_wh = dot(wo, wh), wi_wh = dot(wi, wh);
        if (wo_wh * wi_wh > 0.0f) return Spec<NC>();
        Spec<NC> F(fr_dielectric(dot(wo, wh), eta_a, eta_b));
        Float sqrt_denom = wo_wh + eta * wi_wh;
        Float factor = 1.0f / eta;  // TransportMode::Radiance
        Float dd = fabsf(dist.d(wh) * dist.g(wo, wi) * eta * eta * abs_dot(wi, wh) * abs_dot(wo, wh) * factor * factor /
                         (cti * cto * sqrt_denom * sqrt_denom));
        return ((Spec<NC>(1.0f) - F) * t) * dd;
      }
      case BX_LAMBERT_T: return t * INV_PI;
      case BX_FRESNEL_BLEND: {  // fresnel_blend.rs:37-57
        Spec<NC> diffuse = r * (Spec<NC>(1.0f) - t) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo))) * (28.0f / (23.0f * PI));
        Vec3 wh = wi + wo;
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spec<NC>();
        wh = normalize(wh);
        Spec<NC> schlick = t + (Spec<NC>(1.0f) - t) * pow5(1.0f - dot(wi, wh));
        Spec<NC> specular = schlick * (dist.d(wh) / (4.0f * abs_dot(wi, wh) * rmax(abs_cos_theta(wi), abs_cos_theta(wo))));
        return diffuse + specular;
      }
      default: return Spec<NC>();  // specular lobes
    }
  }
  Float pdf(Vec3 wo, Vec3 wi) const {
    switch (kind) {
      case BX_LAMBERT: case BX_OREN_NAYAR: return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
      case BX_MF_REFL: {
        if (!same_hemisphere(wo, wi)) return 0.0f;
        Vec3 wh = normalize(wo + wi);
        if (dot(wo, wh) < 0.0f) return 0.0f;
        return dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
      }
      case BX_MF_TRANS: {
        if (same_hemisphere(wo, wi)) return 0.0f;
        Float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
        Vec3 wh = normalize(wo + (wi * eta));
        Float wo_wh = dot(wo, wh), wi_wh = dot(wi, wh);
        if (wo_wh * wi_wh > 0.0f) return 0.0f;
        Float sqrt_denom = wo_wh + eta * wi_wh;
        Float dwh_dwi = fabsf((eta * eta * wi_wh) / (sqrt_denom * sqrt_denom));
        return dist.pdf(wo, wh) * dwh_dwi;
      }
      case BX_LAMBERT_T: return !same_hemisphere(wo, wi) ? abs_cos_theta(wi) : 0.0f;  // lambertian.rs:80-86: no 1 / pi
      case BX_FRESNEL_BLEND: {  // fresnel_blend.rs:92-103
        if (!same_hemisphere(wo, wi)) return 0.0f;
        Vec3 wh = normalize(wo + wi);
        if (dot(wo, wh) < 0.0f) return 0.0f;
        return 0.5f * (abs_cos_theta(wi) * INV_PI + dist.pdf(wo, wh) / (4.0f * dot(wo, wh)));
      }
      default: return 0.0f;
    }
  }
  // returns false for None; sampled_type 0 = "use get_type()"
  bool sample_raw(Vec3 wo, Vec2 u, Spec<NC>* fo, Vec3* wio, Float* pdfo, int* stype) const {
    *stype = 0;
    switch (kind) {
      case BX_LAMBERT: case BX_OREN_NAYAR: {  // bxdf.rs:74-85 sample_f_default
        Vec3 wi = cosine_sample_hemisphere(u);
        if (wo.z < 0.0f) wi.z *= -1.0f;
        *pdfo = pdf(wo, wi);
        *fo = f_raw(wo, wi);
        *wio = wi;
        return true;
      }
      case BX_LAMBERT_T: {  // lambertian.rs:63-79
        Vec3 wi = cosine_sample_hemisphere(u);
        if (wo.z > 0.0f) wi.z *= -1.0f;
        *pdfo = pdf(wo, wi);
        if (!(*pdfo > 0.0f)) return false;
        *fo = f_raw(wo, wi);
        *wio = wi;
        return true;
      }
      case BX_FRESNEL_BLEND: {  // fresnel_blend.rs:59-90
        Float ux = u.x;
        Vec3 wi;
        if (ux < 0.5f) {
          ux = rmin(2.0f * ux, ONE_MINUS_EPSILON);
          wi = cosine_sample_hemisphere(Vec2(ux, u.y));
          if (wo.z < 0.0f) wi.z *= -1.0f;
        } else {
          ux = rmin(2.0f * (ux - 0.5f), ONE_MINUS_EPSILON);
          Vec3 wh = dist.sample_wh(wo, Vec2(ux, u.y));
          wi = reflect(wo, wh);
          if (!same_hemisphere(wo, wi)) return false;
        }
        *pdfo = pdf(wo, wi);
        if (!(*pdfo > 0.0f)) return false;
        *fo = f_raw(wo, wi);
        *wio = wi;
        return true;
      }
      case BX_SPEC_REFL: {
        Vec3 wi(-wo.x, -wo.y, wo.z);
        *pdfo = 1.0f;
        Spec<NC> s = fresnel.evaluate(cos_theta(wi));
        *fo = (s * r) / abs_cos_theta(wi);
        *wio = wi;
        return true;
      }
      case BX_SPEC_TRANS: {
        bool entering = cos_theta(wo) > 0.0f;
        Float ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
        Vec3 wi;
        if (!refract(wo, face_forward(Vec3(0, 0, 1), wo), ei / et, &wi)) return false;
        *pdfo = 1.0f;
        Spec<NC> ft = t * (Spec<NC>(1.0f) - Spec<NC>(fr_dielectric(cos_theta(wi), eta_a, eta_b)));
        ft = ft * ((ei * ei) / (et * et));
        *fo = ft / abs_cos_theta(wi);
        *wio = wi;
        return true;
      }
      case BX_FRESNEL_SPEC: {
        Float F = fr_dielectric(cos_theta(wo), eta_a, eta_b);
        if (u.x < F) {
          Vec3 wi(-wo.x, -wo.y, wo.z);
          *pdfo = F;
          *stype = BSDF_SPECULAR | BSDF_REFLECTION;
          Float k = F / abs_cos_theta(wi);
          *fo = r * k;
          *wio = wi;
          return true;
        }
        bool entering = cos_theta(wo) > 0.0f;
        Float ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
        Vec3 wi;
        if (!refract(wo, face_forward(Vec3(0, 0, 1), wo), ei / et, &wi)) return false;
        Spec<NC> ft = t * (1.0f - F);
        ft = ft * ((ei * ei) / (et * et));
        *stype = BSDF_SPECULAR | BSDF_TRANSMISSION;
        *pdfo = 1.0f - F;
        *fo = ft / abs_cos_theta(wi);
        *wio = wi;
        return true;
      }
      case BX_MF_REFL: {
        if (wo.z == 0.0f) return false;
        Vec3 wh = dist.sample_wh(wo, u);
        if (dot(wo, wh) < 0.0f) return false;
        Vec3 wi = reflect(wo, wh);
        if (!same_hemisphere(wo, wi)) return false;
        Float p = dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
        if (p == 0.0f) return false;
        *pdfo = p;
        *fo = f_raw(wo, wi);
        *wio = wi;
        return true;
      }
      default: {  // BX_MF_TRANS
        if (wo.z == 0.0f) return false;
        Vec3 wh = dist.sample_wh(wo, u);
        if (dot(wo, wh) < 0.0f) return false;
        Float eta = cos_theta(wo) > 0.0f ? eta_a / eta_b : eta_b / eta_a;
        Vec3 wi;
        if (!refract(wo, wh, eta, &wi)) return false;
        Float p = pdf(wo, wi);
        if (!(p > 0.0f)) return false;
        *pdfo = p;
        *fo = f_raw(wo, wi);
        *wio = wi;
        return true;
      }
    }
  }
};

// ---- SurfaceInteraction built from a hit (triangle.rs:354-450, surface_interaction.rs) -------------------
struct SurfInt {
  Vec3 p, p_error, n, wo;
  Vec3 dpdu;
  Vec3 sn, sdpdu;  // shading.n, shading.dpdu
  int tri = -1, instance = -1;
  Float time = 0;
};

inline void tri_uvs(const pt_scene_desc* d, int tri, Vec2 uv[3]) {
  if ((d->tri_flags[tri] & PT_TRI_HAS_UV) && d->tri_uv) {
    const float* q = d->tri_uv + (size_t)tri * 6;
    uv[0] = Vec2(q[0], q[1]); uv[1] = Vec2(q[2], q[3]); uv[2] = Vec2(q[4], q[5]);
  } else { uv[0] = Vec2(0, 0); uv[1] = Vec2(1, 0); uv[2] = Vec2(1, 1); }
}

// Triangle::intersect's interaction part, in the triangle's own space; `r` is the ray in that space.
inline SurfInt tri_interaction(const pt_scene_desc* d, int tri, const Ray& r, Float b0, Float b1, Float b2) {
  TriRef T = get_tri(d, tri);
  Vec3 p0 = T.p0, p1 = T.p1, p2 = T.p2;
  bool rev = (T.flags & PT_TRI_REVERSE_ORIENTATION) != 0, swp = (T.flags & PT_TRI_SWAPS_HANDEDNESS) != 0;
  Vec3 n = cross(p0 - p2, p1 - p2);
  if (rev ^ swp) n = n * -1.0f;
  n = normalize(n);
  // get_dpdu_dpdv (triangle.rs:132-186)
  Vec2 uv[3];
  tri_uvs(d, tri, uv);
  Vec2 duv02 = uv[0] - uv[2], duv12 = uv[1] - uv[2];
  Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
  Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
  bool degenerate_uv = fabsf(determinant) < 1e-8f;
  Vec3 dpdu, dpdv;
  bool ok = false;
  if (!degenerate_uv) {
    Float invdet = 1.0f / determinant;
    dpdu = (duv12.y * dp02 - duv02.y * dp12) * invdet;
    dpdv = ((-duv12.x) * dp02 + duv02.x * dp12) * invdet;
    ok = !(length_squared(cross(dpdu, dpdv)) <= 0.0f);
  }
  if (!ok) {
    Vec3 ng = cross(p2 - p0, p1 - p0);
    coordinate_system(normalize(ng), &dpdu, &dpdv);
  }
  SurfInt si;
  Float xs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
  Float ys = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
  Float zs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
  si.p_error = GAMMA7 * Vec3(xs, ys, zs);
  si.p = b0 * p0 + b1 * p1 + b2 * p2;
  si.wo = -r.d;
  si.n = n;
  si.dpdu = dpdu;
  si.sn = n;
  si.sdpdu = dpdu;
  si.tri = tri;
  si.time = r.time;
  const bool has_n = (T.flags & PT_TRI_HAS_N) && d->tri_n, has_s = (T.flags & PT_TRI_HAS_S) && d->tri_s;
  if (has_n || has_s) {  // triangle.rs:376-444
    Vec3 ns = si.n;
    if (has_n) {
      const float* q = d->tri_n + (size_t)tri * 9;
      Vec3 n0(q[0], q[1], q[2]), n1(q[3], q[4], q[5]), n2(q[6], q[7], q[8]);
      Vec3 nns = b0 * n0 + b1 * n1 + b2 * n2;
      if (length_squared(nns) > 0.0f) ns = normalize(nns);
    }
    Vec3 ss = normalize(si.dpdu);
    if (has_s) {
      const float* q = d->tri_s + (size_t)tri * 9;
      Vec3 s0(q[0], q[1], q[2]), s1(q[3], q[4], q[5]), s2(q[6], q[7], q[8]);
      Vec3 nss = b0 * s0 + b1 * s1 + b2 * s2;
      if (length_squared(nss) > 0.0f) ss = normalize(nss);
    }
    Vec3 ts = cross(ns, ss);
    if (length_squared(ts) > 0.0f) {
      ts = normalize(ts);
      ss = normalize(cross(ts, ns));
    } else {
      coordinate_system(ns, &ss, &ts);
    }
    if (rev) ts = ts * -1.0f;
    // set_shading_geometry(ss, ts, .., true) (surface_interaction.rs:140-161)
    si.sn = normalize(cross(ss, ts));
    si.n = face_forward(si.n, si.sn);
    si.sdpdu = ss;
  }
  return si;
}

// Transform::transform_surface_interaction (transform.rs:299-323), instance -> world
inline SurfInt transform_si(const pt_instance& I, const SurfInt& s) {
  Mat4 M = mat_of(I.m), Minv = mat_of(I.minv);
  SurfInt r = s;
  Float x = s.p.x, y = s.p.y, z = s.p.z;
  r.p = M.point(s.p);
  const float* m = M.m;
  Float ex = (GAMMA3 + 1.0f) * (fabsf(m[0]) * s.p_error.x + fabsf(m[1]) * s.p_error.y + fabsf(m[2]) * s.p_error.z) +
             GAMMA3 * (fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z) + fabsf(m[3]));
  Float ey = (GAMMA3 + 1.0f) * (fabsf(m[4]) * s.p_error.x + fabsf(m[5]) * s.p_error.y + fabsf(m[6]) * s.p_error.z) +
             GAMMA3 * (fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z) + fabsf(m[7]));
  Float ez = (GAMMA3 + 1.0f) * (fabsf(m[8]) * s.p_error.x + fabsf(m[9]) * s.p_error.y + fabsf(m[10]) * s.p_error.z) +
             GAMMA3 * (fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z) + fabsf(m[11]));
  r.p_error = Vec3(ex, ey, ez);
  r.n = normalize(Minv.normal_with(s.n));
  r.wo = normalize(M.vector(s.wo));
  r.dpdu = M.vector(s.dpdu);
  r.sn = normalize(Minv.normal_with(s.sn));
  r.sdpdu = M.vector(s.sdpdu);
  r.sn = face_forward(r.sn, r.n);
  return r;
}

inline SurfInt make_si(const pt_scene_desc* d, const Hit& h) {
  SurfInt si = tri_interaction(d, h.tri, h.local_ray, h.b0, h.b1, h.b2);
  if (h.instance >= 0) si = transform_si(d->instances[h.instance], si);
  si.instance = h.instance;
  return si;
}

// ---- BSDF (reflection/bsdf.rs) -----------------------------------------------------------------------
template <int NC>
struct BSDF {
  Float eta = 1;
  Vec3 ns, ng, ss, ts;
  int n = 0;
  BxDF<NC> bx[8];  // MAX_BXDFS = 8 (bsdf.rs)
  bool valid = false;  // si.bsdf.is_some()

  void init(const SurfInt& si, Float eta_) {
    eta = eta_;
    ns = si.sn;
    ng = si.n;
    ss = normalize(si.sdpdu);
    ts = normalize(cross(ns, ss));
    n = 0;
    valid = true;
  }
  void add(const BxDF<NC>& b) { bx[n++] = b; }
  int num_components(int flags) const { int c = 0; for (int i = 0; i < n; i++) if (bx[i].matches(flags)) c++; return c; }
  Vec3 w2l(Vec3 v) const { return Vec3(dot(v, ss), dot(v, ts), dot(v, ns)); }
  Vec3 l2w(Vec3 v) const {
    return Vec3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
  }
  static bool finite3(Vec3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

  Spec<NC> f(Vec3 wo_w, Vec3 wi_w, int flags) const {
    Vec3 wi = w2l(wi_w), wo = w2l(wo_w);
    if (wo.z == 0.0f || !finite3(wo)) return Spec<NC>();
    bool refl = (dot(wi_w, ng) * dot(wo_w, ng)) > 0.0f;
    Spec<NC> r;
    for (int i = 0; i < n; i++) {
      int tp = bx[i].type();
      if (bx[i].matches(flags) && ((refl && (tp & BSDF_REFLECTION)) || (!refl && (tp & BSDF_TRANSMISSION)))) r = r + bx[i].f(wo, wi);
    }
    return r;
  }
  Float pdf(Vec3 wo_w, Vec3 wi_w, int flags) const {
    Vec3 wi = w2l(wi_w), wo = w2l(wo_w);
    if (wo.z == 0.0f || !finite3(wo)) return 0.0f;
    int count = 0;
    Float p = 0.0f;
    for (int i = 0; i < n; i++)
      if (bx[i].matches(flags)) { p = p + bx[i].pdf(wo, wi); count++; }
    return count > 0 ? p / (Float)count : 0.0f;
  }
  bool sample_f(Vec3 wo_w, Vec2 u, int flags, Spec<NC>* fo, Vec3* wi_wo, Float* pdfo, int* stype) const {
    int matching = num_components(flags);
    if (matching == 0) return false;
    int comp = std::min(f2i(floorf(u.x * (Float)matching)), matching - 1);
    int index = -1, count = comp;
    for (int i = 0; i < n; i++)
      if (bx[i].matches(flags)) { if (count == 0) { index = i; break; } count--; }
    const BxDF<NC>& b = bx[index];
    Vec2 remapped(rmin((u.x * (Float)matching) - (Float)comp, ONE_MINUS_EPSILON), u.y);
    Vec3 wo = w2l(wo_w);
    if (wo.z == 0.0f || !finite3(wo)) return false;
    int sampled = b.type();
    Spec<NC> fv;
    Vec3 wi;
    Float pdf_;
    int t;
    if (!b.sample_f(wo, remapped, &fv, &wi, &pdf_, &t)) return false;
    if (pdf_ <= 0.0f) return false;
    if (t != 0) sampled = t;
    int btype = b.type();
    Vec3 wi_w = l2w(wi);
    if ((btype & BSDF_SPECULAR) == 0 && matching > 1)
      for (int i = 0; i < n; i++)
        if (i != index && bx[i].matches(flags)) pdf_ += bx[i].pdf(wo, wi);
    if (matching > 1) pdf_ /= (Float)matching;
    if ((btype & BSDF_SPECULAR) == 0) {
      bool refl = (dot(wi_w, ng) * dot(wo_w, ng)) > 0.0f;
      Spec<NC> r;
      for (int i = 0; i < n; i++) {
        int tp = bx[i].type();
        if (bx[i].matches(flags) && ((refl && (tp & BSDF_REFLECTION)) || (!refl && (tp & BSDF_TRANSMISSION)))) r = r + bx[i].f(wo, wi);
      }
      fv = r;
    }
    *fo = fv; *wi_wo = wi_w; *pdfo = pdf_; *stype = sampled;
    return true;
  }
};

// ---- Materials (materials/*.rs compute_scattering_functions; constant textures) -----------------------------
template <int NC>
inline void material_scattering(const pt_scene_desc* d, int mi, const SurfInt& si, BSDF<NC>* bsdf, bool allow_multiple_lobes);
template <int NC>
inline void compute_scattering(const pt_scene_desc* d, const SurfInt& si, BSDF<NC>* bsdf, bool allow_multiple_lobes = true) {
  bsdf->valid = false;
  int mi = d->tri_material[si.tri];
  if (mi < 0) return;
  material_scattering<NC>(d, mi, si, bsdf, allow_multiple_lobes);
}
template <int NC>
inline void material_scattering(const pt_scene_desc* d, int mi, const SurfInt& si, BSDF<NC>* bsdf, bool allow_multiple_lobes) {
  bsdf->valid = false;
  const pt_material& M = d->materials[mi];
  auto S = [&](int idx) { return Spec<NC>::load(d->spectra + (size_t)idx * NC); };
  switch (M.type) {
    case PT_MAT_MIX: {  // mix.rs:60-93 (kd, ks = the two materials, kr = amount)
      Spec<NC> s1 = S(M.kr).clamp_zero();
      Spec<NC> s2 = (Spec<NC>(1.0f) - s1).clamp_zero();
      BSDF<NC> b1, b2;
      material_scattering<NC>(d, M.kd, si, &b1, allow_multiple_lobes);
      material_scattering<NC>(d, M.ks, si, &b2, allow_multiple_lobes);
      if (!b1.valid || !b2.valid) return;  // the reference asserts both exist
      bsdf->init(si, b1.eta);
      for (int i = 0; i < b1.n; i++) { BxDF<NC> b = b1.bx[i]; b.scaled = true; b.scale = s1; bsdf->add(b); }
      for (int i = 0; i < b2.n; i++) { BxDF<NC> b = b2.bx[i]; b.scaled = true; b.scale = s2; bsdf->add(b); }
      break;
    }
    case PT_MAT_MATTE: {
      bsdf->init(si, 1.0f);
      Spec<NC> r = S(M.kd);  // NOT clamped (matte.rs:39-41)
      Float sig = rclamp(M.sigma, 0.0f, 90.0f);
      if (!r.is_black()) {
        BxDF<NC> b;
        b.r = r;
        if (sig == 0.0f) b.kind = BX_LAMBERT;
        else {
          b.kind = BX_OREN_NAYAR;
          Float s = sig * (PI / 180.0f);
          Float s2 = s * s;
          b.a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
          b.b = 0.45f * s2 / (s2 + 0.09f);
        }
        bsdf->add(b);
      }
      break;
    }
    case PT_MAT_PLASTIC: {
      bsdf->init(si, 1.0f);
      Spec<NC> kd = S(M.kd).clamp_zero();
      if (!kd.is_black()) { BxDF<NC> b; b.kind = BX_LAMBERT; b.r = kd; bsdf->add(b); }
      Spec<NC> ks = S(M.ks).clamp_zero();
      if (!ks.is_black()) {
        Float rough = M.uroughness;
        if (M.remap_roughness) rough = roughness_to_alpha(rough);
        BxDF<NC> b;
        b.kind = BX_MF_REFL;
        b.r = ks;
        b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.5f; b.fresnel.eta_t = 1.0f;
        b.dist = TRDist(rough, rough, true);
        bsdf->add(b);
      }
      break;
    }
    case PT_MAT_MIRROR: {
      bsdf->init(si, 1.0f);
      Spec<NC> r = S(M.kr).clamp_zero();
      if (!r.is_black()) { BxDF<NC> b; b.kind = BX_SPEC_REFL; b.r = r; b.fresnel.kind = FR_NOOP; bsdf->add(b); }
      break;
    }
    case PT_MAT_GLASS: {
      Float eta = M.eta, ur = M.uroughness, vr = M.vroughness;
      Spec<NC> r = S(M.kr), t = S(M.ks);
      if (r.is_black() && t.is_black()) return;
      bsdf->init(si, eta);
      bool is_spec = ur == 0.0f && vr == 0.0f;
      if (is_spec && allow_multiple_lobes) {  // allow_multiple_lobes = true in PathIntegrator (path.rs:106)
        BxDF<NC> b; b.kind = BX_FRESNEL_SPEC; b.r = r; b.t = t; b.eta_a = 1.0f; b.eta_b = eta; bsdf->add(b);
      } else if (is_spec) {  // glass.rs:80-99 (directlighting / whitted): SpecularReflection + SpecularTransmission
        if (!r.is_black()) {
          BxDF<NC> b; b.kind = BX_SPEC_REFL; b.r = r;
          b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.0f; b.fresnel.eta_t = eta;
          bsdf->add(b);
        }
        if (!t.is_black()) { BxDF<NC> b; b.kind = BX_SPEC_TRANS; b.t = t; b.eta_a = 1.0f; b.eta_b = eta; bsdf->add(b); }
      } else {
        if (M.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        if (!r.is_black()) {
          BxDF<NC> b; b.kind = BX_MF_REFL; b.r = r;
          b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.0f; b.fresnel.eta_t = eta;
          b.dist = TRDist(ur, vr, true);
          bsdf->add(b);
        }
        if (!t.is_black()) {
          BxDF<NC> b; b.kind = BX_MF_TRANS; b.t = t; b.eta_a = 1.0f; b.eta_b = eta; b.dist = TRDist(ur, vr, true);
          bsdf->add(b);
        }
      }
      break;
    }
    case PT_MAT_METAL: {
      bsdf->init(si, 1.0f);
      Float ur = M.uroughness, vr = M.vroughness;
      if (M.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
      BxDF<NC> b;
      b.kind = BX_MF_REFL;
      b.r = Spec<NC>(1.0f);
      b.fresnel.kind = FR_CONDUCTOR; b.fresnel.c_eta_i = Spec<NC>(1.0f); b.fresnel.c_eta_t = S(M.kr); b.fresnel.c_k = S(M.ks);
      b.dist = TRDist(ur, vr, true);
      bsdf->add(b);
      break;
    }
    case PT_MAT_UBER: {  // uber.rs:73-128
      Float e = M.eta;
      Spec<NC> op = S(M.opacity);
      Spec<NC> t = (Spec<NC>(1.0f) - op).clamp_zero();
      bsdf->init(si, !t.is_black() ? 1.0f : e);
      if (!t.is_black()) { BxDF<NC> b; b.kind = BX_SPEC_TRANS; b.t = t; b.eta_a = 1.0f; b.eta_b = 1.0f; bsdf->add(b); }
      Spec<NC> kd = op * S(M.kd).clamp_zero();
      if (!kd.is_black()) { BxDF<NC> b; b.kind = BX_LAMBERT; b.r = kd; bsdf->add(b); }
      Spec<NC> ks = op * S(M.ks).clamp_zero();
      if (!ks.is_black()) {
        Float ur = M.uroughness, vr = M.vroughness;
        if (M.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        BxDF<NC> b; b.kind = BX_MF_REFL; b.r = ks;
        b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.0f; b.fresnel.eta_t = e;
        b.dist = TRDist(ur, vr, true);
        bsdf->add(b);
      }
      Spec<NC> kr = op * S(M.kr).clamp_zero();
      if (!kr.is_black()) {
        BxDF<NC> b; b.kind = BX_SPEC_REFL; b.r = kr;
        b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.0f; b.fresnel.eta_t = e;
        bsdf->add(b);
      }
      Spec<NC> kt = op * S(M.kt).clamp_zero();
      if (!kt.is_black()) { BxDF<NC> b; b.kind = BX_SPEC_TRANS; b.t = kt; b.eta_a = 1.0f; b.eta_b = e; bsdf->add(b); }
      break;
    }
    case PT_MAT_SUBSTRATE: {  // substrate.rs:33-60
      Spec<NC> d = S(M.kd).clamp_zero(), sp = S(M.ks).clamp_zero();
      bsdf->init(si, 1.0f);
      if (!d.is_black() && !sp.is_black()) {
        Float ur = M.uroughness, vr = M.vroughness;
        if (M.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        BxDF<NC> b; b.kind = BX_FRESNEL_BLEND; b.r = d; b.t = sp; b.dist = TRDist(ur, vr, true);
        bsdf->add(b);
      }
      break;
    }
    case PT_MAT_TRANSLUCENT: {  // translucent.rs:37-100
      const Float eta = 1.5f;
      Spec<NC> r = S(M.kr).clamp_zero(), t = S(M.kt).clamp_zero();
      if (r.is_black() && t.is_black()) return;  // returns before si.bsdf is set
      bsdf->init(si, eta);
      Spec<NC> kd = S(M.kd).clamp_zero();
      if (!kd.is_black()) {
        if (!r.is_black()) { BxDF<NC> b; b.kind = BX_LAMBERT; b.r = r * kd; bsdf->add(b); }
        if (!t.is_black()) { BxDF<NC> b; b.kind = BX_LAMBERT_T; b.t = t * kd; bsdf->add(b); }
      }
      Spec<NC> ks = S(M.ks).clamp_zero();
      if (!ks.is_black()) {
        Float rough = M.uroughness;
        if (M.remap_roughness) rough = roughness_to_alpha(rough);
        if (!r.is_black()) {
          BxDF<NC> b; b.kind = BX_MF_REFL; b.r = r * ks;
          b.fresnel.kind = FR_DIELECTRIC; b.fresnel.eta_i = 1.0f; b.fresnel.eta_t = eta;
          b.dist = TRDist(rough, rough, true);
          bsdf->add(b);
        }
        if (!t.is_black()) { BxDF<NC> b; b.kind = BX_MF_TRANS; b.t = t * ks; b.eta_a = 1.0f; b.eta_b = eta; b.dist = TRDist(rough, rough, true); bsdf->add(b); }
      }
      break;
    }
    default: return;
  }
}

// ---- Lights ---------------------------------------------------------------------------------------------
struct RefPoint {  // the Interaction fields sample_li / pdf_li read
  Vec3 p, p_error, n;
  Float time = 0;
};
struct ShadowSeg { Ray ray; };  // VisibilityTester reduced to its spawn_ray_to ray

inline Ray spawn_ray(const RefPoint& it, Vec3 d) { return Ray(offset_ray_origin(it.p, it.p_error, it.n, d), d, INF, it.time); }
// Interaction::spawn_ray_to(it) (interaction.rs:118-127)
inline Ray spawn_ray_to(const RefPoint& a, const RefPoint& b) {
  Vec3 origin = offset_ray_origin(a.p, a.p_error, a.n, b.p - a.p);
  Vec3 target = offset_ray_origin(b.p, b.p_error, b.n, origin - b.p);
  Vec3 d = target - origin;
  return Ray(origin, d, 1.0f - SHADOW_EPSILON, a.time);
}

// Triangle::sample (triangle.rs:590-619)
inline void tri_sample(const pt_scene_desc* d, int tri, Vec2 u, RefPoint* out, Float* pdf) {
  TriRef T = get_tri(d, tri);
  Vec2 b = uniform_sample_triangle(u);
  Vec3 p0 = T.p0, p1 = T.p1, p2 = T.p2;
  Vec3 p = b.x * p0 + b.y * p1 + (1.0f - b.x - b.y) * p2;
  Vec3 n = normalize(cross(p1 - p0, p2 - p0));
  if ((T.flags & PT_TRI_HAS_N) && d->tri_n) {
    const float* q = d->tri_n + (size_t)tri * 9;
    Vec3 ns = b.x * Vec3(q[0], q[1], q[2]) + b.y * Vec3(q[3], q[4], q[5]) + (1.0f - b.x - b.y) * Vec3(q[6], q[7], q[8]);
    n = face_forward(n, ns);
  } else if (((T.flags & PT_TRI_REVERSE_ORIENTATION) != 0) ^ ((T.flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) {
    n = n * -1.0f;
  }
  Vec3 pas = vabs(b.x * p0) + vabs(b.y * p1) + vabs((1.0f - b.x - b.y) * p2);
  out->p = p;
  out->p_error = GAMMA6 * Vec3(pas.x, pas.y, pas.z);
  out->n = n;
  Float area = 0.5f * length(cross(p1 - p0, p2 - p0));
  *pdf = 1.0f / area;
}
inline Float tri_area(const pt_scene_desc* d, int tri) {
  TriRef T = get_tri(d, tri);
  return 0.5f * length(cross(T.p1 - T.p0, T.p2 - T.p0));
}
// Triangle::sample_from (triangle.rs:621-651)
inline bool tri_sample_from(const pt_scene_desc* d, int tri, const RefPoint& ref, Vec2 u, RefPoint* out, Float* pdf) {
  Float p_area;
  tri_sample(d, tri, u, out, &p_area);
  Vec3 wi = out->p - ref.p;
  if (length_squared(wi) <= 0.0f) return false;
  wi = normalize(wi);
  if (!(d->tri_flags[tri] & PT_TRI_TWO_SIDED)) {
    if (dot(out->n, -wi) <= 0.0f) return false;
  }
  Float p = p_area * distance_squared(ref.p, out->p) / abs_dot(out->n, -wi);
  if (p <= 0.0f || std::isinf(p)) return false;
  *pdf = p;
  return true;
}

// ---- InfiniteAreaLight (lights/infinite.rs) ------------------------------------------------------------------------
// MIPMap<RGBSpectrum> (core/texture/mipmap.rs: resample_image :316-404 for sides that are not powers of two, make_pyramid :406-441 with downsample_half :225-275,
// triangle :711-767, lookup :620-637; create_spectrum_mipmap :1013-1026 = trilinear off, s repeats, t clamps) and the
// Distribution2D built by make_distribution (infinite.rs:67-93; core/sampling/distribution.rs:109-145).
struct MipLevel { int w = 0, h = 0; std::vector<Float> rgb; };
struct EnvMap {
  Float l2w[9], w2l[9];
  Float xf[16], screen[4], hitcher = 0, cos_total_width = 0;  // ProjectionLight / GonioPhotometricLight extras (pt_envmap)
  std::vector<MipLevel> pyramid;
  std::vector<Distribution1D> conditional_v;
  Distribution1D marginal;

  static std::vector<Float> downsample_half(const std::vector<Float>& img, int width, int height) {  // halves the width
    int nw = width / 2;
    std::vector<Float> out((size_t)3 * nw * height);
    for (int y = 0; y < height; y++)
      for (int x = 0; x < nw; x++)
        for (int c = 0; c < 3; c++) {
          size_t i0 = (size_t)y * width + (2 * x + 0), i1 = (size_t)y * width + (2 * x + 1);
          out[3 * ((size_t)y * nw + x) + c] = img[3 * i0 + c] * 0.5f + img[3 * i1 + c] * 0.5f;
        }
    return out;
  }
  static std::vector<Float> transpose(const std::vector<Float>& img, int width, int height) {
    std::vector<Float> out(img.size());
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++)
        for (int c = 0; c < 3; c++) out[3 * ((size_t)x * height + y) + c] = img[3 * ((size_t)y * width + x) + c];
    return out;
  }
  void make_pyramid(int w, int h, std::vector<Float> data) {
    MipLevel L; L.w = w; L.h = h; L.rgb = data;
    pyramid.push_back(L);
    if (w * h == 1) return;
    std::vector<Float> s = data;
    int sw = w, sh = h;
    if (w > 1) { s = downsample_half(s, w, h); sw = w / 2; }
    if (h > 1) {
      s = transpose(s, sw, sh);          // now sh wide, sw high
      s = downsample_half(s, sh, sw);
      s = transpose(s, sh / 2, sw);
      sh = h / 2;
    }
    make_pyramid(sw, sh, s);
  }
  // ---- resampling of maps whose sides are not powers of two (mipmap.rs:291-404) ----
  static Float lanczos(Float x, Float tau) {  // core/texture/noise.rs:150-162
    x = std::fabs(x);
    if (x < 1e-5f) return 1.0f;
    if (x > 1.0f) return 0.0f;
    x *= PI;
    Float s = std::sin(x * tau) / (x * tau);
    Float l = std::sin(x) / x;
    return s * l;
  }
  struct ResampleWeight { int first_texel; Float weight[4]; };
  static std::vector<ResampleWeight> resample_weights(size_t old_res, size_t new_res) {  // mipmap.rs:291-314
    std::vector<ResampleWeight> wt(new_res);
    const Float filter_width = 2.0f;
    for (size_t i = 0; i < new_res; i++) {
      Float center = ((Float)i + 0.5f) * ((Float)old_res / (Float)new_res);
      Float first_texel = floorf((center - filter_width) + 0.5f);
      for (int j = 0; j < 4; j++) {
        Float pos = first_texel + (Float)j + 0.5f;
        wt[i].weight[j] = lanczos((pos - center) / filter_width, filter_width);
      }
      Float inv_sum_wts = 1.0f / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
      for (int j = 0; j < 4; j++) wt[i].weight[j] *= inv_sum_wts;
      wt[i].first_texel = (int)first_texel;
    }
    return wt;
  }
  static int math_mod(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }
  // resample_image (mipmap.rs:316-404) with swrap = Repeat, twrap = Clamp
  static std::vector<Float> resample_image(const std::vector<Float>& img, int rw, int rh, int* ow, int* oh) {
    int pw = 1, ph = 1;
    while (pw < rw) pw *= 2;
    while (ph < rh) ph *= 2;
    *ow = pw; *oh = ph;
    if (pw == rw && ph == rh) return img;
    std::vector<Float> out((size_t)3 * pw * ph, 0.0f);
    std::vector<ResampleWeight> sw = resample_weights((size_t)rw, (size_t)pw);
    for (int t = 0; t < rh; t++)
      for (int s = 0; s < pw; s++)
        for (int j = 0; j < 4; j++) {
          int orig_s = math_mod(sw[(size_t)s].first_texel + j, rw);
          if (orig_s >= 0 && orig_s < rw) {
            Float w = sw[(size_t)s].weight[j];
            size_t src = (size_t)t * rw + orig_s, dst = (size_t)t * pw + s;
            for (int c = 0; c < 3; c++) out[3 * dst + c] += img[3 * src + c] * w;
          }
        }
    std::vector<ResampleWeight> tw = resample_weights((size_t)rh, (size_t)ph);
    for (int s = 0; s < pw; s++) {
      std::vector<Float> buffer((size_t)3 * rh);
      for (int t = 0; t < rh; t++)
        for (int c = 0; c < 3; c++) buffer[3 * (size_t)t + c] = out[3 * ((size_t)t * pw + s) + c];
      for (int t = 0; t < ph; t++) {
        Float l[3] = {0.0f, 0.0f, 0.0f};
        for (int j = 0; j < 4; j++) {
          int orig_t = std::min(std::max(tw[(size_t)t].first_texel + j, 0), rh - 1);
          if (orig_t >= 0 && orig_t < rh) {
            Float w = tw[(size_t)t].weight[j];
            for (int c = 0; c < 3; c++) l[c] += buffer[3 * (size_t)orig_t + c] * w;
          }
        }
        for (int c = 0; c < 3; c++) out[3 * ((size_t)t * pw + s) + c] = l[c];
      }
    }
    for (Float& v : out) v = std::max(v, 0.0f);
    return out;
  }
  void triangle(size_t level, Vec2 st, Float out[3]) const {
    level = std::min(level, pyramid.size() - 1);
    const MipLevel& I = pyramid[level];
    Float s = st.x * (Float)I.w - 0.5f, t = st.y * (Float)I.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    Float ds = s - (Float)s0, dt = t - (Float)t0;
    int s1 = s0 + 1, t1 = t0 + 1;
    s0 = s0 & (I.w - 1); s1 = s1 & (I.w - 1);                                      // ImageWrap::Repeat
    t0 = std::min(std::max(t0, 0), I.h - 1); t1 = std::min(std::max(t1, 0), I.h - 1);  // ImageWrap::Clamp
    for (int c = 0; c < 3; c++) {
      Float a = I.rgb[3 * ((size_t)t0 * I.w + s0) + c] * ((1.0f - ds) * (1.0f - dt));
      Float b = I.rgb[3 * ((size_t)t1 * I.w + s0) + c] * ((1.0f - ds) * dt);
      Float cc = I.rgb[3 * ((size_t)t0 * I.w + s1) + c] * (ds * (1.0f - dt));
      Float d = I.rgb[3 * ((size_t)t1 * I.w + s1) + c] * (ds * dt);
      out[c] = a + b + cc + d;
    }
  }
  void lookup(Vec2 st, Float width, Float out[3]) const {
    Float max_level = (Float)(pyramid.size() - 1);
    Float level = max_level + log2f(rmax(width, 1e-8f));
    if (level < 0.0f) { triangle(0, st, out); return; }
    if (level >= max_level) { for (int c = 0; c < 3; c++) out[c] = pyramid.back().rgb[c]; return; }
    size_t il = (size_t)floorf(level);
    Float delta = rclamp(level - (Float)il, 0.0f, 1.0f);
    Float a[3], b[3];
    triangle(il, st, a);
    triangle(il + 1, st, b);
    for (int c = 0; c < 3; c++) out[c] = a[c] * (1.0f - delta) + b[c] * delta;
  }
  static Float rgb_y(const Float* c) { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }
  void init(const pt_envmap& M) {
    for (int i = 0; i < 9; i++) { l2w[i] = M.l2w[i]; w2l[i] = M.w2l[i]; }
    for (int i = 0; i < 16; i++) xf[i] = M.xf[i];
    for (int i = 0; i < 4; i++) screen[i] = M.screen_bounds[i];
    hitcher = M.hitcher; cos_total_width = M.cos_total_width;
    int rw = 0, rh = 0;
    std::vector<Float> data = resample_image(std::vector<Float>(M.texels, M.texels + (size_t)M.width * M.height * 3), M.width, M.height, &rw, &rh);
    make_pyramid(rw, rh, data);
    size_t width = (size_t)rw * 2, height = (size_t)rh * 2;   // make_distribution: lmap.width() / height() (infinite.rs:70-72)
    std::vector<Float> img(width * height);
    Float fwidth = 0.5f / (Float)std::min(width, height);
    for (size_t v = 0; v < height; v++) {
      Float vp = ((Float)v + 0.5f) / (Float)height;
      Float sin_theta = sinf(PI * vp);
      for (size_t u = 0; u < width; u++) {
        Float up = ((Float)u + 0.5f) / (Float)width;
        Float c[3];
        lookup(Vec2(up, vp), fwidth, c);
        img[v * width + u] = rmax(rgb_y(c), 0.0f) * sin_theta;
      }
    }
    std::vector<Float> mf;
    for (size_t v = 0; v < height; v++) {
      conditional_v.emplace_back(std::vector<Float>(img.begin() + v * width, img.begin() + (v + 1) * width));
      mf.push_back(conditional_v.back().func_int);
    }
    marginal = Distribution1D(mf);
  }
  Vec2 sample_continuous(Vec2 u, Float* pdf) const {
    Float pdf1, pdf0;
    size_t v, dummy;
    Float d1 = marginal.sample_continuous(u.y, &pdf1, &v);
    Float d0 = conditional_v[v].sample_continuous(u.x, &pdf0, &dummy);
    *pdf = pdf0 * pdf1;
    return Vec2(d0, d1);
  }
  Float pdf(Vec2 p) const {
    size_t uc = conditional_v[0].func.size(), vc = marginal.func.size();
    Float fu = p.x * (Float)uc, fv = p.y * (Float)vc;
    size_t iu = fu > 0.0f ? (size_t)fu : 0, iv = fv > 0.0f ? (size_t)fv : 0;   // `as usize` saturates
    iu = std::min(iu, uc - 1); iv = std::min(iv, vc - 1);
    return conditional_v[iv].func[iu] / marginal.func_int;
  }
  Vec3 to_light(Vec3 v) const { return Vec3(w2l[0] * v.x + w2l[1] * v.y + w2l[2] * v.z, w2l[3] * v.x + w2l[4] * v.y + w2l[5] * v.z, w2l[6] * v.x + w2l[7] * v.y + w2l[8] * v.z); }
  Vec3 to_world(Vec3 v) const { return Vec3(l2w[0] * v.x + l2w[1] * v.y + l2w[2] * v.z, l2w[3] * v.x + l2w[4] * v.y + l2w[5] * v.z, l2w[6] * v.x + l2w[7] * v.y + l2w[8] * v.z); }
};
inline Float spherical_theta(Vec3 v) { return acosf(rclamp(v.z, -1.0f, 1.0f)); }                         // geometry/misc.rs:96-104
inline Float spherical_phi(Vec3 v) { Float p = atan2f(v.y, v.x); return p < 0.0f ? p + 2.0f * PI : p; }
// InfiniteAreaLight::le (infinite.rs:110-121); RGB spectra only
template <int NC>
inline Spec<NC> env_le(const EnvMap& E, Vec3 ray_d) {
  Vec3 w = normalize(E.to_light(ray_d));
  Float c[3];
  E.lookup(Vec2(spherical_phi(w) * INV_2_PI, spherical_theta(w) * INV_PI), 0.0f, c);
  Spec<NC> r;
  for (int k = 0; k < 3 && k < NC; k++) r.c[k] = c[k];
  return r;
}

template <int NC>
struct LightSample {
  Spec<NC> li;
  Vec3 wi;
  Float pdf;
  Ray shadow;
};

template <int NC>
inline Spec<NC> area_l(const pt_scene_desc* d, const pt_light& L, Vec3 n, Vec3 w) {  // DiffuseAreaLight::l (diffuse.rs:156-163)
  if (L.two_sided || dot(n, w) > 0.0f) return Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC);
  return Spec<NC>();
}

template <int NC>
inline bool light_sample_li(const pt_scene_desc* d, Float world_radius, int li, const RefPoint& it, Vec2 u, LightSample<NC>* out, const EnvMap* envs = nullptr) {
  const pt_light& L = d->lights[li];
  if (L.type == PT_LIGHT_INFINITE) {  // infinite.rs:123-166
    const EnvMap& E = envs[L.tri];
    Float map_pdf;
    Vec2 uv = E.sample_continuous(u, &map_pdf);
    if (map_pdf <= 0.0f) return false;
    Float theta = uv.y * PI, phi = uv.x * 2.0f * PI;
    Float cos_theta = cosf(theta), sin_theta = rclamp(sinf(theta), 0.0f, 1.0f);
    Float cos_phi = cosf(phi), sin_phi = sinf(phi);
    Vec3 wi = E.to_world(Vec3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
    out->wi = wi;
    out->pdf = sin_theta == 0.0f ? 0.0f : map_pdf / (2.0f * PI * PI * sin_theta);
    RefPoint pl;
    pl.p = it.p + wi * (2.0f * world_radius);
    pl.time = it.time;
    out->shadow = spawn_ray_to(it, pl);
    Float c[3];
    E.lookup(uv, 0.0f, c);
    out->li = Spec<NC>();
    for (int k = 0; k < 3 && k < NC; k++) out->li.c[k] = c[k];
    return true;
  }
  if (L.type == PT_LIGHT_AREA) {  // diffuse.rs:65-82
    RefPoint ps;
    Float pdf;
    if (!tri_sample_from(d, L.tri, it, u, &ps, &pdf)) return false;
    if (pdf <= 0.0f || length_squared(ps.p - it.p) <= 0.0f) return false;
    Vec3 wi = normalize(ps.p - it.p);
    out->shadow = spawn_ray_to(it, ps);
    out->li = area_l<NC>(d, L, ps.n, -wi);
    out->wi = wi;
    out->pdf = pdf;
    return true;
  }
  Vec3 v(L.v[0], L.v[1], L.v[2]);
  RefPoint pl;
  if (L.type == PT_LIGHT_POINT) {  // point.rs:44-58
    out->li = Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC) / distance_squared(v, it.p);
    out->wi = normalize(v - it.p);
    out->pdf = 1.0f;
    pl.p = v;
  } else if (L.type == PT_LIGHT_PROJECTION) {  // projection.rs:81-99, projection() :64-79
    const EnvMap& E = envs[L.tri];
    Vec3 wi = normalize(v - it.p);
    Vec3 w = -wi;
    Float rgb[3] = {0.0f, 0.0f, 0.0f};
    Vec3 wl(E.w2l[0] * w.x + E.w2l[1] * w.y + E.w2l[2] * w.z, E.w2l[3] * w.x + E.w2l[4] * w.y + E.w2l[5] * w.z, E.w2l[6] * w.x + E.w2l[7] * w.y + E.w2l[8] * w.z);
    if (!(wl.z < E.hitcher)) {
      Vec3 p = mat_of(E.xf).point(wl);
      if (p.x >= E.screen[0] && p.x <= E.screen[2] && p.y >= E.screen[1] && p.y <= E.screen[3]) {
        Vec2 st(p.x - E.screen[0], p.y - E.screen[1]);  // Bounds2::offset (bounds2.rs:44-53)
        if (E.screen[2] > E.screen[0]) st.x = st.x / (E.screen[2] - E.screen[0]);
        if (E.screen[3] > E.screen[1]) st.y = st.y / (E.screen[3] - E.screen[1]);
        E.lookup(st, 0.0f, rgb);
      }
    }
    Spec<NC> m;
    for (int k = 0; k < 3 && k < NC; k++) m.c[k] = rgb[k];
    out->li = (Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC) * m) / distance_squared(v, it.p);
    out->wi = wi;
    out->pdf = 1.0f;
    pl.p = v;
  } else if (L.type == PT_LIGHT_GONIOMETRIC) {  // goniometric.rs:47-62, scale() :33-41
    const EnvMap& E = envs[L.tri];
    Vec3 wi = normalize(v - it.p);
    Vec3 wp = mat_of(E.xf).point(-wi);  // world_to_light.transform_point(w): a point transform, as the reference does
    Float rgb[3];
    E.lookup(Vec2(spherical_phi(wp) * INV_2_PI, spherical_theta(wp) * INV_PI), 0.0f, rgb);
    Spec<NC> m;
    for (int k = 0; k < 3 && k < NC; k++) m.c[k] = rgb[k];
    out->li = (Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC) * m) / distance_squared(v, it.p);
    out->wi = wi;
    out->pdf = 1.0f;
    pl.p = v;
  } else if (L.type == PT_LIGHT_SPOT) {  // spot.rs:62-80, falloff :44-57
    const pt_spot& S = d->spots[L.tri];
    Vec3 wi = normalize(v - it.p);
    Vec3 w = -wi;
    Vec3 wl = normalize(Vec3(S.w2l[0] * w.x + S.w2l[1] * w.y + S.w2l[2] * w.z, S.w2l[3] * w.x + S.w2l[4] * w.y + S.w2l[5] * w.z,
                             S.w2l[6] * w.x + S.w2l[7] * w.y + S.w2l[8] * w.z));
    Float ct = wl.z, fall;
    if (ct < S.cos_total_width) fall = 0.0f;
    else if (ct >= S.cos_falloff_start) fall = 1.0f;
    else {
      Float delta = (ct - S.cos_total_width) / (S.cos_falloff_start - S.cos_total_width);
      fall = (delta * delta) * (delta * delta);
    }
    out->li = Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC) * (fall / distance_squared(v, it.p));
    out->wi = wi;
    out->pdf = 1.0f;
    pl.p = v;
  } else {  // distant.rs:62-80
    out->li = Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC);
    out->wi = v;
    out->pdf = 1.0f;
    pl.p = it.p + v * (2.0f * world_radius);
  }
  pl.time = it.time;
  out->shadow = spawn_ray_to(it, pl);
  return true;
}
// Shape::pdf_from (shape.rs:40-54) for the light's triangle
inline Float light_pdf_li(const pt_scene_desc* d, int li, const RefPoint& it, Vec3 wi, const EnvMap* envs = nullptr) {
  const pt_light& L = d->lights[li];
  if (L.type == PT_LIGHT_INFINITE) {  // infinite.rs:168-185
    const EnvMap& E = envs[L.tri];
    Vec3 wl = E.to_light(wi);
    Float theta = spherical_theta(wl), phi = spherical_phi(wl);
    Float sin_theta = rclamp(sinf(theta), 0.0f, 1.0f);
    if (sin_theta == 0.0f) return 0.0f;
    return E.pdf(Vec2(phi * INV_2_PI, theta * INV_PI)) / (2.0f * PI * PI * sin_theta);
  }
  if (L.type != PT_LIGHT_AREA) return 0.0f;
  Ray ray = spawn_ray(it, wi);
  TriHit th;
  if (!tri_intersect(get_tri(d, L.tri), ray, &th)) return 0.0f;
  SurfInt s = tri_interaction(d, L.tri, ray, th.b0, th.b1, th.b2);
  Float pdf = distance_squared(it.p, s.p) / (abs_dot(s.n, -wi) * tri_area(d, L.tri));
  if (std::isinf(pdf)) return 0.0f;
  return pdf;
}
inline bool light_is_delta(const pt_light& L) { return L.type != PT_LIGHT_AREA && L.type != PT_LIGHT_INFINITE; }

template <int NC>
inline Spec<NC> light_power(const pt_scene_desc* d, Float world_radius, int li, const EnvMap* envs = nullptr) {
  const pt_light& L = d->lights[li];
  Spec<NC> s = Spec<NC>::load(d->spectra + (size_t)L.spectrum * NC);
  if (L.type == PT_LIGHT_INFINITE) {  // infinite.rs:103-108
    Float c[3];
    envs[L.tri].lookup(Vec2(0.5f, 0.5f), 0.5f, c);
    Spec<NC> r;
    for (int k = 0; k < 3 && k < NC; k++) r.c[k] = c[k];
    return r * (PI * world_radius * world_radius);
  }
  if (L.type == PT_LIGHT_PROJECTION || L.type == PT_LIGHT_GONIOMETRIC) {
    Float c[3];
    envs[L.tri].lookup(Vec2(0.5f, 0.5f), 0.5f, c);
    Spec<NC> r;
    for (int k = 0; k < 3 && k < NC; k++) r.c[k] = c[k];
    if (L.type == PT_LIGHT_PROJECTION) return r * s * (2.0f * PI * (1.0f - envs[L.tri].cos_total_width));  // projection.rs:113-117
    return (s * (4.0f * PI)) * r;                                                                          // goniometric.rs:64-68
  }
  if (L.type == PT_LIGHT_AREA) return s * ((L.two_sided ? 2.0f : 1.0f) * tri_area(d, L.tri) * PI);
  if (L.type == PT_LIGHT_POINT) return s * (4.0f * PI);
  if (L.type == PT_LIGHT_SPOT) {  // spot.rs:82-85 (minus, unlike pbrt-v3's plus)
    const pt_spot& S = d->spots[L.tri];
    return s * (2.0f * PI * (1.0f - 0.5f * (S.cos_falloff_start - S.cos_total_width)));
  }
  return s * (PI * world_radius * world_radius);
}

// ---- light distributions (lightdistrib/{spatial,power,uniform}.rs) -----------------------------------------
template <int NC>
struct LightDistrib {
  const pt_scene_desc* d = nullptr;
  Float world_radius = 0;
  Bounds3 wb;
  int strategy = PT_LIGHTS_SPATIAL;
  uint32_t voxels[3] = {1, 1, 1};
  Distribution1D fixed;                                      // uniform / power
  mutable std::vector<std::unique_ptr<Distribution1D>> grid;  // spatial: dense, lazily filled (a pure function of the voxel)

  const EnvMap* envs = nullptr;
  void init(const pt_scene_desc* desc, const Bounds3& world_bound, Float radius, const EnvMap* env_maps = nullptr) {
    d = desc; wb = world_bound; world_radius = radius; envs = env_maps;
    int strat = d->integrator.light_strategy;
    if (strat == PT_LIGHTS_UNIFORM && d->n_lights != 1) strat = PT_LIGHTS_SPATIAL;  // create_light_sample_distribution.rs:23-29
    strategy = strat;
    if (strat == PT_LIGHTS_UNIFORM) fixed = Distribution1D(std::vector<Float>(d->n_lights, 1.0f));
    else if (strat == PT_LIGHTS_POWER) {
      std::vector<Float> pw;
      for (uint32_t i = 0; i < d->n_lights; i++) pw.push_back(spec_y(d, light_power<NC>(d, world_radius, (int)i, envs)));
      fixed = Distribution1D(pw);
    } else {
      const uint32_t max_voxels = 64;
      Vec3 diag = wb.diagonal();
      Float bmax = diag[wb.maximum_extent()];
      for (int i = 0; i < 3; i++) {
        uint32_t v = f2u(ceilf(diag[i] / bmax * (Float)max_voxels));
        voxels[i] = std::min(std::max(v, 1u), max_voxels);
      }
      grid.resize((size_t)voxels[0] * voxels[1] * voxels[2]);
    }
  }
  void voxel_of(Vec3 p, uint32_t pi[3]) const {  // spatial.rs:86-111
    Vec3 o = wb.offset(p);
    Float off[3] = {rclamp(o.x, 0.0f, 1.0f), rclamp(o.y, 0.0f, 1.0f), rclamp(o.z, 0.0f, 1.0f)};
    for (int i = 0; i < 3; i++) pi[i] = std::min(f2u(off[i] * (Float)voxels[i]), voxels[i] - 1);
  }
  std::vector<Float> voxel_func(const uint32_t pi[3]) const {  // spatial.rs:113-197
    Vec3 p0((Float)pi[0] / (Float)voxels[0], (Float)pi[1] / (Float)voxels[1], (Float)pi[2] / (Float)voxels[2]);
    Vec3 p1((Float)(pi[0] + 1) / (Float)voxels[0], (Float)(pi[1] + 1) / (Float)voxels[1], (Float)(pi[2] + 1) / (Float)voxels[2]);
    Bounds3 vb(wb.lerp3(p0), wb.lerp3(p1));
    const size_t n_samples = 128;
    size_t nl = d->n_lights;
    std::vector<Float> contrib(nl, 0.0f);
    for (size_t i = 0; i < n_samples; i++) {
      Vec3 t(radical_inverse(0, i), radical_inverse(1, i), radical_inverse(2, i));
      RefPoint it;
      it.p = vb.lerp3(t);
      Vec2 u(radical_inverse(3, i), radical_inverse(4, i));
      for (size_t j = 0; j < nl; j++) {
        LightSample<NC> ls;
        if (light_sample_li<NC>(d, world_radius, (int)j, it, u, &ls, envs))
          if (ls.pdf > 0.0f) contrib[j] += spec_y(d, ls.li) / ls.pdf;
      }
    }
    Float sum = 0.0f;
    for (size_t j = 0; j < nl; j++) sum += contrib[j];
    Float avg = sum / (Float)(n_samples * nl);
    Float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (size_t j = 0; j < nl; j++) contrib[j] = rmax(min_contrib, contrib[j]);
    return contrib;
  }
  const Distribution1D* lookup(Vec3 p) const {
    if (strategy != PT_LIGHTS_SPATIAL) return &fixed;
    uint32_t pi[3];
    voxel_of(p, pi);
    size_t idx = ((size_t)pi[0] * voxels[1] + pi[1]) * voxels[2] + pi[2];
    if (!grid[idx]) grid[idx] = std::make_unique<Distribution1D>(voxel_func(pi));
    return grid[idx].get();
  }
};

}  // namespace orc
