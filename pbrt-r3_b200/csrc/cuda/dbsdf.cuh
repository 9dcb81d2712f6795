// dbsdf.cuh — surface interaction, material -> lobe set, BSDF eval / pdf / sample on the device.
// Replaces reference src/shapes/triangle.rs:132-186,354-450 (interaction part of Triangle::intersect),
// src/core/interaction/surface_interaction.rs:140-161, src/core/transform/transform.rs:299-323,
// src/materials/{matte,plastic,glass,metal,mirror}.rs (compute_scattering_functions, constant textures),
// src/core/reflection/{bsdf,bxdf,math,lambertian,oren_nayar,specular,fresnel,microfacet}.rs,
// src/core/distribution/{microfacet,trowbridge_reitz}.rs.
// Design: no per-hit heap objects (the reference allocates Arc<BSDF> + Vec<Arc<dyn BxDF>> per hit). A material
// expands to at most two lobes described by small PODs; spectra stay in HBM/L2 and are fetched per use.
// PathIntegrator calls compute_scattering_functions with allow_multiple_lobes = true (path.rs:106), so smooth
// glass is always FresnelSpecular; SpecularTransmission / dielectric SpecularReflection are unreachable.
#pragma once
#include "dscene.cuh"

namespace ptd {

template <int NC>
struct Sp {
  float c[NC];
};
template <int NC> PT_HD Sp<NC> sp_const(float v) { Sp<NC> r;
#pragma unroll
  for (int i = 0; i < NC; i++) r.c[i] = v; return r; }
template <int NC> PT_HD Sp<NC> sp_load(const float* p) { Sp<NC> r;
#pragma unroll
  for (int i = 0; i < NC; i++) r.c[i] = ldg(p + i); return r; }
template <int NC> PT_HD Sp<NC> operator+(Sp<NC> a, const Sp<NC>& b) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] + b.c[i]; return a; }
template <int NC> PT_HD Sp<NC> operator-(Sp<NC> a, const Sp<NC>& b) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] - b.c[i]; return a; }
template <int NC> PT_HD Sp<NC> operator*(Sp<NC> a, const Sp<NC>& b) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] * b.c[i]; return a; }
template <int NC> PT_HD Sp<NC> operator/(Sp<NC> a, const Sp<NC>& b) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] / b.c[i]; return a; }
template <int NC> PT_HD Sp<NC> operator*(Sp<NC> a, float s) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] * s; return a; }
template <int NC> PT_HD Sp<NC> operator/(Sp<NC> a, float s) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = a.c[i] / s; return a; }
template <int NC> PT_HD bool sp_black(const Sp<NC>& a) {
  if (NC == 3) { return a.c[0] == 0.0f && a.c[1] == 0.0f && a.c[2] == 0.0f; }  // RGBSpectrum::is_black: all == 0
  bool blk = true;                                                               // SampledSpectrum::is_black: all <= 0
#pragma unroll
  for (int i = 0; i < NC; i++) blk = blk && (a.c[i] <= 0.0f);
  return blk;
}
template <int NC> PT_HD float sp_max(const Sp<NC>& a) { float m = a.c[0];
#pragma unroll
  for (int i = 1; i < NC; i++) m = fmaxf(m, a.c[i]); return m; }
template <int NC> PT_HD Sp<NC> sp_clamp0(Sp<NC> a) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = clampf(a.c[i], 0.0f, pt_inf()); return a; }
template <int NC> PT_HD Sp<NC> sp_sqrt(Sp<NC> a) {
#pragma unroll
  for (int i = 0; i < NC; i++) a.c[i] = sqrtf(a.c[i]); return a; }
template <int NC> PT_HD bool sp_valid(const Sp<NC>& a) { bool ok = true;
#pragma unroll
  for (int i = 0; i < NC; i++) ok = ok && isfinite_(a.c[i]); return ok; }
template <int NC> PT_HD float sp_y(const DScene& S, const Sp<NC>& a) {
  const float* w = S.xyz_w + 60;
  float acc = w[0] * a.c[0];
#pragma unroll
  for (int i = 1; i < NC; i++) acc += w[i] * a.c[i];
  return acc * S.xyz_scale;
}
template <int NC> PT_HD void sp_xyz(const DScene& S, const Sp<NC>& a, float xyz[3]) {
  for (int r = 0; r < 3; r++) {
    const float* w = S.xyz_w + r * 60;
    float acc = w[0] * a.c[0];
#pragma unroll
    for (int i = 1; i < NC; i++) acc += w[i] * a.c[i];
    xyz[r] = acc * S.xyz_scale;
  }
}

// ---- local-frame helpers (reflection/math.rs) -----------------------------------------------------------
PT_HD float cth(f3 w) { return w.z; }
PT_HD float c2th(f3 w) { return w.z * w.z; }
PT_HD float acth(f3 w) { return fabsf(w.z); }
PT_HD float s2th(f3 w) { return fmaxf(0.0f, 1.0f - c2th(w)); }
PT_HD float sth(f3 w) { return sqrtf(s2th(w)); }
PT_HD float tth(f3 w) { return sth(w) / cth(w); }
PT_HD float t2th(f3 w) { return s2th(w) / c2th(w); }
PT_HD float cphi(f3 w) { float s = sth(w); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
PT_HD float sphi(f3 w) { float s = sth(w); return s == 0.0f ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
PT_HD bool same_hemi(f3 a, f3 b) { return a.z * b.z > 0.0f; }
PT_HD f3 reflect3(f3 wo, f3 n) { f3 a = (2.0f * dot3(wo, n)) * n; f3 b = -wo; return a + b; }
PT_HD bool refract3(f3 wi, f3 n, float eta, f3* wt) {
  float ci = dot3(n, wi);
  float s2i = fmaxf(0.0f, 1.0f - ci * ci);
  float s2t = eta * eta * s2i;
  if (s2t >= 1.0f) return false;
  float ct = sqrtf(1.0f - s2t);
  f3 a = eta * (-wi);
  f3 b = (eta * ci - ct) * n;
  *wt = a + b;
  return true;
}

// sampling.rs:106-145
PT_HD f2 concentric_disk(f2 u) {
  float ox = u.x * 2.0f - 1.0f, oy = u.y * 2.0f - 1.0f;
  if (ox == 0.0f && oy == 0.0f) return mk2(0.0f, 0.0f);
  float r, theta;
  if (fabsf(ox) > fabsf(oy)) { r = ox; theta = PT_PI_OVER_4 * (oy / ox); }
  else { r = oy; theta = PT_PI_OVER_2 - PT_PI_OVER_4 * (ox / oy); }
  return mk2(r * pt_cosf(theta), r * pt_sinf(theta));
}
PT_HD f3 cosine_hemisphere(f2 u) {
  f2 d = concentric_disk(u);
  float z = sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y));
  return mk3(d.x, d.y, z);
}
PT_HD f3 uniform_hemisphere(f2 u) {  // sampling.rs:85-90
  float z = u.x;
  float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  float phi = 2.0f * PT_PI * u.y;
  return mk3(r * pt_cosf(phi), r * pt_sinf(phi), z);
}
PT_HD float power_heuristic1(float f, float g) { return (f * f) / (f * f + g * g); }  // nf = ng = 1

// ---- Fresnel (fresnel.rs:16-82) ---------------------------------------------------------------------------
PT_HD float fr_diel(float ci, float ei, float et) {
  ci = clampf(ci, -1.0f, 1.0f);
  if (!(ci > 0.0f)) { float t = ei; ei = et; et = t; ci = fabsf(ci); }
  float si = sqrtf(fmaxf(0.0f, 1.0f - ci * ci));
  float st = ei / et * si;
  if (st >= 1.0f) return 1.0f;
  float ct = sqrtf(fmaxf(0.0f, 1.0f - st * st));
  float rparl = ((et * ci) - (ei * ct)) / ((et * ci) + (ei * ct));
  float rperp = ((ei * ci) - (et * ct)) / ((ei * ci) + (et * ct));
  return (rparl * rparl + rperp * rperp) / 2.0f;
}
// FresnelConductor with eta_i = 1 (metal.rs:73): per channel
PT_HD float fr_cond1(float ci, float eta_t, float kk) {
  float eta = eta_t / 1.0f, etak = kk / 1.0f;
  float c2 = ci * ci;
  float s2 = 1.0f - c2;
  float s22 = s2 * s2;
  float eta2 = eta * eta, etak2 = etak * etak;
  float t0 = eta2 - etak2 - s2;
  float a2b2 = sqrtf(t0 * t0 + eta2 * etak2 * 4.0f);
  float t1 = a2b2 + c2;
  float a = sqrtf((a2b2 + t0) * 0.5f);
  float t2 = ci * a * 2.0f;
  float rs = (t1 - t2) / (t1 + t2);
  float t3 = c2 * a2b2 + s22;
  float t4 = t2 * s2;
  float rp = rs * (t3 - t4) / (t3 + t4);
  return (rp + rs) * 0.5f;
}

// ---- Trowbridge-Reitz (trowbridge_reitz.rs), samplevis = true everywhere on this path ---------------------
struct TR {
  float ax, ay;
};
PT_HD float tr_d(TR D, f3 wh) {
  float t2 = t2th(wh);
  if (isinf_(t2)) return 0.0f;
  float c2 = c2th(wh);
  float c4 = c2 * c2;
  float cp = cphi(wh), sp = sphi(wh);
  float e = ((cp * cp) / (D.ax * D.ax) + (sp * sp) / (D.ay * D.ay)) * t2;
  float e2 = (1.0f + e) * (1.0f + e);
  return 1.0f / (PT_PI * D.ax * D.ay * c4 * e2);
}
PT_HD float tr_lambda(TR D, f3 w) {
  float att = fabsf(tth(w));
  if (isinf_(att)) return 0.0f;
  float cp = cphi(w), sp = sphi(w);
  float alpha = sqrtf((cp * cp) * D.ax * D.ax + (sp * sp) * D.ay * D.ay);
  float a2t2 = (alpha * att) * (alpha * att);
  return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
PT_HD float tr_g1(TR D, f3 w) { return 1.0f / (1.0f + tr_lambda(D, w)); }
PT_HD float tr_g(TR D, f3 wo, f3 wi) { return 1.0f / (1.0f + tr_lambda(D, wo) + tr_lambda(D, wi)); }
PT_HD float tr_pdf(TR D, f3 wo, f3 wh) { return tr_d(D, wh) * tr_g1(D, wo) * absdot3(wo, wh) / acth(wo); }
PT_HD void tr_sample11(float ct, float u1, float u2, float* sx, float* sy) {
  if (ct > 0.9999f) {
    float r = sqrtf(u1 / (1.0f - u1));
    float phi = 2.0f * PT_PI * u2;
    *sx = r * pt_cosf(phi);
    *sy = r * pt_sinf(phi);
    return;
  }
  float st = sqrtf(fmaxf(0.0f, 1.0f - ct * ct));
  float tt = st / ct;
  float a = 1.0f / tt;
  float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
  a = 2.0f * u1 / g1 - 1.0f;
  float tmp = fminf(1e10f, 1.0f / (a * a - 1.0f));
  float b = tt;
  float d = sqrtf(fmaxf(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
  float s1 = b * tmp - d, s2 = b * tmp + d;
  float slope_x = (a < 0.0f || s2 > 1.0f / tt) ? s1 : s2;
  float s;
  if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
  else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
  float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *sx = slope_x;
  *sy = s * z * sqrtf(1.0f + slope_x * slope_x);
}
PT_HD f3 tr_sample_wh(TR D, f3 wo, f2 u) {
  bool flip = wo.z < 0.0f;
  f3 w = flip ? -wo : wo;
  f3 ws = norm3(mk3(D.ax * w.x, D.ay * w.y, w.z));
  float sx, sy;
  tr_sample11(cth(ws), u.x, u.y, &sx, &sy);
  float cp = cphi(ws), sp = sphi(ws);
  float tmp = cp * sx - sp * sy;
  sy = sp * sx + cp * sy;
  sx = tmp;
  sx *= D.ax;
  sy *= D.ay;
  f3 wh = norm3(mk3(-sx, -sy, 1.0f));
  return flip ? -wh : wh;
}

// ---- lobes --------------------------------------------------------------------------------------------------
enum { LB_LAMBERT = 0, LB_OREN = 1, LB_MIRROR = 2, LB_FRESNEL_SPEC = 3, LB_MF_REFL = 4, LB_MF_TRANS = 5,
       // smooth glass without allow_multiple_lobes (directlighting / whitted; glass.rs:80-99): SpecularReflection with a
       // dielectric Fresnel term and SpecularTransmission (specular.rs) as two lobes
       LB_SPEC_REFL_D = 6, LB_SPEC_TRANS = 7,
       // translucent: LambertianTransmission (lambertian.rs:59-99); substrate: FresnelBlend (fresnel_blend.rs)
       LB_LAMBERT_T = 8, LB_FRESNEL_BLEND = 9 };
enum { F_REFL = 1, F_TRANS = 2, F_DIFF = 4, F_GLOSSY = 8, F_SPEC = 16, F_ALL = 31 };
enum { FR_NOOP = 0, FR_DIEL = 1, FR_COND = 2 };

struct Lobe {
  int kind;
  int sp0, sp1;   // spectrum indices: reflectance R (or T) | conductor: eta, k ; FresnelSpecular: R, T
  int fres;       // FR_*
  float a, b;     // Oren-Nayar A,B | alpha_x, alpha_y
  float ei, et;   // dielectric Fresnel eta_i, eta_t | transmission eta_a, eta_b
  int scale = -1; // ScaledBxDF (MixMaterial): spectrum index of the factor on f, -1 = none; only read by the five-lobe BSDFs
};
// Compile-time lobe sets: a shade kernel specialised for one material class only keeps the code of the lobes and
// Fresnel terms that class can produce (bits 0-9: LB_* kinds, bits 12-14: FR_* kinds).
#define PT_CM_ANY 0x73ff
#define PT_CM_MATTE 0x1003
#define PT_CM_PLASTIC 0x3011
#define PT_CM_MIRROR 0x1004
#define PT_CM_GLASS 0x3038
#define PT_CM_METAL 0x4010
#if defined(__CUDACC__) || defined(__GNUC__)
#define PT_UNREACHABLE() __builtin_unreachable()
#else
#define PT_UNREACHABLE()
#endif
#define PT_KIND(K) case K: if (!((CM >> (K)) & 1)) { PT_UNREACHABLE(); }
#define PT_HAS_FRES(F) (((CM >> (12 + (F))) & 1) != 0)

// CM: the class's lobe set, so that a specialised kernel only compares against the kinds it can meet
template <int CM = PT_CM_ANY>
PT_HD int lobe_type(int kind) {
  switch (kind) {
    case LB_LAMBERT: case LB_OREN: if (!(CM & 3)) { PT_UNREACHABLE(); } return F_REFL | F_DIFF;
    PT_KIND(LB_MIRROR) return F_REFL | F_SPEC;
    PT_KIND(LB_FRESNEL_SPEC) return F_REFL | F_TRANS | F_SPEC;
    PT_KIND(LB_MF_REFL) return F_REFL | F_GLOSSY;
    PT_KIND(LB_SPEC_REFL_D) return F_REFL | F_SPEC;
    PT_KIND(LB_SPEC_TRANS) return F_TRANS | F_SPEC;
    PT_KIND(LB_LAMBERT_T) return F_TRANS | F_DIFF;
    PT_KIND(LB_FRESNEL_BLEND) return F_REFL | F_GLOSSY;
    default: if (!((CM >> LB_MF_TRANS) & 1)) { PT_UNREACHABLE(); } return F_TRANS | F_GLOSSY;
  }
}
template <int CM = PT_CM_ANY>
PT_HD bool lobe_matches(int kind, int flags) { int t = lobe_type<CM>(kind); return (t & flags) == t; }

template <int NC, int CM = PT_CM_ANY>
PT_HD Sp<NC> lobe_fresnel(const DScene& S, const Lobe& L, float ci) {
  if (PT_HAS_FRES(FR_NOOP) && (L.fres == FR_NOOP || (!PT_HAS_FRES(FR_DIEL) && !PT_HAS_FRES(FR_COND)))) return sp_const<NC>(1.0f);
  if (PT_HAS_FRES(FR_DIEL) && (L.fres == FR_DIEL || !PT_HAS_FRES(FR_COND))) return sp_const<NC>(fr_diel(ci, L.ei, L.et));
  if (!PT_HAS_FRES(FR_COND)) return sp_const<NC>(1.0f);
  Sp<NC> r;
  float c = clampf(fabsf(ci), -1.0f, 1.0f);
  const float* e = S.spectra + (size_t)L.sp0 * NC;
  const float* k = S.spectra + (size_t)L.sp1 * NC;
#pragma unroll
  for (int i = 0; i < NC; i++) r.c[i] = fr_cond1(c, ldg(e + i), ldg(k + i));
  return r;
}

// FresnelBlend's pow5 is `v.powf(5.0)` (fresnel_blend.rs:13-15) = libm powf: v^5 formed in double and rounded once gives the
// correctly rounded value except in near-tie cases, which is what glibc's powf returns (CUDA's powf does not)
PT_HD float pow5(float v) { double d = (double)v; return (float)((d * d) * (d * d) * d); }

template <int NC, int CM = PT_CM_ANY>
PT_HDN Sp<NC> lobe_f(const DScene& S, const Lobe L, f3 wo, f3 wi) {
  switch (L.kind) {
    PT_KIND(LB_LAMBERT) return sp_load<NC>(S.spectra + (size_t)L.sp0 * NC) * PT_INV_PI;
    PT_KIND(LB_OREN) {
      float sti = sth(wi), sto = sth(wo);
      float max_cos = 0.0f;
      if (sti > 1e-4f && sto > 1e-4f) {
        float dcos = cphi(wi) * cphi(wo) + sphi(wi) * sphi(wo);
        max_cos = fmaxf(dcos, 0.0f);
      }
      float sin_alpha, tan_beta;
      if (acth(wi) > acth(wo)) { sin_alpha = sto; tan_beta = sti / acth(wi); }
      else { sin_alpha = sti; tan_beta = sto / acth(wo); }
      return sp_load<NC>(S.spectra + (size_t)L.sp0 * NC) * PT_INV_PI * (L.a + L.b * max_cos * sin_alpha * tan_beta);
    }
    PT_KIND(LB_MF_REFL) {
      float cto = acth(wo), cti = acth(wi);
      f3 wh = wi + wo;
      if (cti == 0.0f || cto == 0.0f) return sp_const<NC>(0.0f);
      if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return sp_const<NC>(0.0f);
      wh = norm3(wh);
      Sp<NC> F = lobe_fresnel<NC, CM>(S, L, dot3(wi, faceforward(wh, mk3(0.0f, 0.0f, 1.0f))));
      TR D; D.ax = L.a; D.ay = L.b;
      float k = tr_d(D, wh) * tr_g(D, wo, wi) / (4.0f * cti * cto);
      Sp<NC> R = (L.fres == FR_COND) ? sp_const<NC>(1.0f) : sp_load<NC>(S.spectra + (size_t)L.sp0 * NC);
      return R * F * k;
    }
    PT_KIND(LB_MF_TRANS) {
      if (same_hemi(wo, wi)) return sp_const<NC>(0.0f);
      float cto = cth(wo), cti = cth(wi);
      if (cti == 0.0f || cto == 0.0f) return sp_const<NC>(0.0f);
      float eta = cto > 0.0f ? L.et / L.ei : L.ei / L.et;
      f3 wh = norm3(wo + (wi * eta));
      if (wh.z < 0.0f) wh = -wh;
      float wo_wh = dot3(wo, wh), wi_wh = dot3(wi, wh);
      if (wo_wh * wi_wh > 0.0f) return sp_const<NC>(0.0f);
      float F = fr_diel(dot3(wo, wh), L.ei, L.et);
      float sq = wo_wh + eta * wi_wh;
      float factor = 1.0f / eta;
      TR D; D.ax = L.a; D.ay = L.b;
      float dd = fabsf(tr_d(D, wh) * tr_g(D, wo, wi) * eta * eta * absdot3(wi, wh) * absdot3(wo, wh) * factor * factor / (cti * cto * sq * sq));
      return ((sp_const<NC>(1.0f) - sp_const<NC>(F)) * sp_load<NC>(S.spectra + (size_t)L.sp0 * NC)) * dd;
    }
    PT_KIND(LB_LAMBERT_T) return sp_load<NC>(S.spectra + (size_t)L.sp0 * NC) * PT_INV_PI;
    PT_KIND(LB_FRESNEL_BLEND) {  // fresnel_blend.rs:37-57
      Sp<NC> rd = sp_load<NC>(S.spectra + (size_t)L.sp0 * NC), rs = sp_load<NC>(S.spectra + (size_t)L.sp1 * NC);
      Sp<NC> diffuse = rd * (sp_const<NC>(1.0f) - rs) * (1.0f - pow5(1.0f - 0.5f * acth(wi))) * (1.0f - pow5(1.0f - 0.5f * acth(wo))) * (28.0f / (23.0f * PT_PI));
      f3 wh = wi + wo;
      if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return sp_const<NC>(0.0f);
      wh = norm3(wh);
      TR D; D.ax = L.a; D.ay = L.b;
      Sp<NC> schlick = rs + (sp_const<NC>(1.0f) - rs) * pow5(1.0f - dot3(wi, wh));
      Sp<NC> specular = schlick * (tr_d(D, wh) / (4.0f * absdot3(wi, wh) * fmaxf(acth(wi), acth(wo))));
      return diffuse + specular;
    }
    default: return sp_const<NC>(0.0f);
  }
}
template <int CM = PT_CM_ANY>
PT_HDN float lobe_pdf(const Lobe L, f3 wo, f3 wi) {
  switch (L.kind) {
    case LB_LAMBERT: case LB_OREN: if (!(CM & 3)) { PT_UNREACHABLE(); } return same_hemi(wo, wi) ? acth(wi) * PT_INV_PI : 0.0f;
    PT_KIND(LB_MF_REFL) {
      if (!same_hemi(wo, wi)) return 0.0f;
      f3 wh = norm3(wo + wi);
      if (dot3(wo, wh) < 0.0f) return 0.0f;
      TR D; D.ax = L.a; D.ay = L.b;
      return tr_pdf(D, wo, wh) / (4.0f * dot3(wo, wh));
    }
    PT_KIND(LB_MF_TRANS) {
      if (same_hemi(wo, wi)) return 0.0f;
      float eta = cth(wo) > 0.0f ? L.et / L.ei : L.ei / L.et;
      f3 wh = norm3(wo + (wi * eta));
      float wo_wh = dot3(wo, wh), wi_wh = dot3(wi, wh);
      if (wo_wh * wi_wh > 0.0f) return 0.0f;
      float sq = wo_wh + eta * wi_wh;
      float dwh = fabsf((eta * eta * wi_wh) / (sq * sq));
      TR D; D.ax = L.a; D.ay = L.b;
      return tr_pdf(D, wo, wh) * dwh;
    }
    PT_KIND(LB_LAMBERT_T) return !same_hemi(wo, wi) ? acth(wi) : 0.0f;  // lambertian.rs:80-86 (no 1 / pi in the reference)
    PT_KIND(LB_FRESNEL_BLEND) {  // fresnel_blend.rs:92-103
      if (!same_hemi(wo, wi)) return 0.0f;
      f3 wh = norm3(wo + wi);
      if (dot3(wo, wh) < 0.0f) return 0.0f;
      TR D; D.ax = L.a; D.ay = L.b;
      return 0.5f * (acth(wi) * PT_INV_PI + tr_pdf(D, wo, wh) / (4.0f * dot3(wo, wh)));
    }
    default: return 0.0f;
  }
}
// lobe_f and lobe_pdf of one lobe for the same pair (wo, wi), sharing what both compute (half vector, D(wh), Lambda(wo)).
// Every shared value is produced by the same operations on the same operands as in the two separate functions
// (wi + wo and wo + wi are the same float sums), so f and pdf are bit-identical to lobe_f / lobe_pdf.
template <int NC, int CM = PT_CM_ANY>
PT_HDN void lobe_f_pdf(const DScene& S, const Lobe L, f3 wo, f3 wi, bool want_f, Sp<NC>* f, float* pdf) {
  if (((CM >> LB_MF_REFL) & 1) && L.kind == LB_MF_REFL) {
    *f = sp_const<NC>(0.0f);
    *pdf = 0.0f;
    const float cto = acth(wo), cti = acth(wi);
    f3 wh = wi + wo;
    const bool f_ok = want_f && !(cti == 0.0f || cto == 0.0f) && !(wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f);
    const bool hemi = same_hemi(wo, wi);
    if (!f_ok && !hemi) return;
    wh = norm3(wh);
    const bool pdf_ok = hemi && !(dot3(wo, wh) < 0.0f);
    if (!f_ok && !pdf_ok) return;
    TR D; D.ax = L.a; D.ay = L.b;
    const float d = tr_d(D, wh), lo = tr_lambda(D, wo);
    if (f_ok) {
      Sp<NC> F = lobe_fresnel<NC, CM>(S, L, dot3(wi, faceforward(wh, mk3(0.0f, 0.0f, 1.0f))));
      float k = d * (1.0f / (1.0f + lo + tr_lambda(D, wi))) / (4.0f * cti * cto);
      Sp<NC> R = (L.fres == FR_COND) ? sp_const<NC>(1.0f) : sp_load<NC>(S.spectra + (size_t)L.sp0 * NC);
      *f = R * F * k;
    }
    if (pdf_ok) *pdf = (d * (1.0f / (1.0f + lo)) * absdot3(wo, wh) / acth(wo)) / (4.0f * dot3(wo, wh));
    return;
  }
  *f = want_f ? lobe_f<NC, CM>(S, L, wo, wi) : sp_const<NC>(0.0f);
  *pdf = lobe_pdf<CM>(L, wo, wi);
}

// BxDF::sample_f of one lobe; stype = 0 means "the lobe's own type".
// For the non-specular lobes *f is NOT evaluated: BSDF::sample_f recomputes f over all matching lobes right after
// (bsdf.rs:178-196; bsdf_sample below), so the reference's own value is always discarded, and no decision depends on it.
template <int NC, int CM = PT_CM_ANY>
PT_HDN bool lobe_sample(const DScene& S, const Lobe L, f3 wo, f2 u, Sp<NC>* f, f3* wi_out, float* pdf, int* stype) {
  *stype = 0;
  switch (L.kind) {
    case LB_LAMBERT: case LB_OREN: {
      if (!(CM & 3)) { PT_UNREACHABLE(); }
      f3 wi = cosine_hemisphere(u);
      if (wo.z < 0.0f) wi.z *= -1.0f;
      *pdf = lobe_pdf<CM>(L, wo, wi);
      *wi_out = wi;   // f: see the note above lobe_sample
      return true;
    }
    PT_KIND(LB_MIRROR) {
      f3 wi = mk3(-wo.x, -wo.y, wo.z);
      *pdf = 1.0f;
      *f = (sp_const<NC>(1.0f) * sp_load<NC>(S.spectra + (size_t)L.sp0 * NC)) / acth(wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_FRESNEL_SPEC) {
      float F = fr_diel(cth(wo), L.ei, L.et);
      if (u.x < F) {
        f3 wi = mk3(-wo.x, -wo.y, wo.z);
        *pdf = F;
        *stype = F_SPEC | F_REFL;
        float k = F / acth(wi);
        *f = sp_load<NC>(S.spectra + (size_t)L.sp0 * NC) * k;
        *wi_out = wi;
        return true;
      }
      bool entering = cth(wo) > 0.0f;
      float ei = entering ? L.ei : L.et, et = entering ? L.et : L.ei;
      f3 wi;
      if (!refract3(wo, faceforward(mk3(0.0f, 0.0f, 1.0f), wo), ei / et, &wi)) return false;
      Sp<NC> ft = sp_load<NC>(S.spectra + (size_t)L.sp1 * NC) * (1.0f - F);
      ft = ft * ((ei * ei) / (et * et));
      *stype = F_SPEC | F_TRANS;
      *pdf = 1.0f - F;
      *f = ft / acth(wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_SPEC_REFL_D) {  // SpecularReflection::sample_f (specular.rs) with FresnelDielectric(ei, et)
      f3 wi = mk3(-wo.x, -wo.y, wo.z);
      *pdf = 1.0f;
      *f = (sp_const<NC>(fr_diel(cth(wi), L.ei, L.et)) * sp_load<NC>(S.spectra + (size_t)L.sp0 * NC)) / acth(wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_SPEC_TRANS) {  // SpecularTransmission::sample_f (specular.rs:78-107), TransportMode::Radiance
      bool entering = cth(wo) > 0.0f;
      float ei = entering ? L.ei : L.et, et = entering ? L.et : L.ei;
      f3 wi;
      if (!refract3(wo, faceforward(mk3(0.0f, 0.0f, 1.0f), wo), ei / et, &wi)) return false;
      *pdf = 1.0f;
      Sp<NC> ft = sp_load<NC>(S.spectra + (size_t)L.sp0 * NC) * (sp_const<NC>(1.0f) - sp_const<NC>(fr_diel(cth(wi), L.ei, L.et)));
      ft = ft * ((ei * ei) / (et * et));
      *f = ft / acth(wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_LAMBERT_T) {  // lambertian.rs:63-79
      f3 wi = cosine_hemisphere(u);
      if (wo.z > 0.0f) wi.z *= -1.0f;
      *pdf = lobe_pdf<CM>(L, wo, wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_FRESNEL_BLEND) {  // fresnel_blend.rs:59-90
      float ux = u.x;
      f3 wi;
      if (ux < 0.5f) {
        ux = fminf(2.0f * ux, PT_ONE_MINUS_EPS);
        wi = cosine_hemisphere(mk2(ux, u.y));
        if (wo.z < 0.0f) wi.z *= -1.0f;
      } else {
        ux = fminf(2.0f * (ux - 0.5f), PT_ONE_MINUS_EPS);
        TR D; D.ax = L.a; D.ay = L.b;
        f3 wh = tr_sample_wh(D, wo, mk2(ux, u.y));
        wi = reflect3(wo, wh);
        if (!same_hemi(wo, wi)) return false;
      }
      *pdf = lobe_pdf<CM>(L, wo, wi);
      *wi_out = wi;
      return true;
    }
    PT_KIND(LB_MF_REFL) {
      if (wo.z == 0.0f) return false;
      TR D; D.ax = L.a; D.ay = L.b;
      f3 wh = tr_sample_wh(D, wo, u);
      if (dot3(wo, wh) < 0.0f) return false;
      f3 wi = reflect3(wo, wh);
      if (!same_hemi(wo, wi)) return false;
      float p = tr_pdf(D, wo, wh) / (4.0f * dot3(wo, wh));
      if (p == 0.0f) return false;
      *pdf = p;
      *wi_out = wi;
      return true;
    }
    default: {
      if (!((CM >> LB_MF_TRANS) & 1)) { PT_UNREACHABLE(); }
      if (wo.z == 0.0f) return false;
      TR D; D.ax = L.a; D.ay = L.b;
      f3 wh = tr_sample_wh(D, wo, u);
      if (dot3(wo, wh) < 0.0f) return false;
      float eta = cth(wo) > 0.0f ? L.ei / L.et : L.et / L.ei;
      f3 wi;
      if (!refract3(wo, wh, eta, &wi)) return false;
      float p = lobe_pdf<CM>(L, wo, wi);
      if (!(p > 0.0f)) return false;
      *pdf = p;
      *wi_out = wi;
      return true;
    }
  }
}

// ---- surface interaction --------------------------------------------------------------------------------------
struct SurfD {
  f3 p, perr, n, wo, sn, sdpdu;
  f3 dpdu;  // of the true geometry (read by the ambientocclusion integrator only)
  int tri;
};

PT_HD void tri_verts(const DScene& S, int tri, f3* p0, f3* p1, f3* p2) {
  const f4* v = S.tri_v + (size_t)tri * 3;
  *p0 = xyz(ldg4(v)); *p1 = xyz(ldg4(v + 1)); *p2 = xyz(ldg4(v + 2));
}

// interaction part of Triangle::intersect in the triangle's own space (rd = ray direction in that space)
PT_HD1 SurfD tri_surface(const DScene& S, int tri, f3 rd, float b0, float b1, float b2) {
  f3 p0, p1, p2;
  tri_verts(S, tri, &p0, &p1, &p2);
  uint32_t flags = ldg(S.tri_flags + tri);
  bool rev = (flags & PT_TRI_REVERSE_ORIENTATION) != 0, swp = (flags & PT_TRI_SWAPS_HANDEDNESS) != 0;
  f3 dp02 = p0 - p2, dp12 = p1 - p2;
  f3 n = cross3(dp02, dp12);
  if (rev != swp) n = n * -1.0f;
  n = norm3(n);
  f2 uv0 = mk2(0.0f, 0.0f), uv1 = mk2(1.0f, 0.0f), uv2 = mk2(1.0f, 1.0f);
  if ((flags & PT_TRI_HAS_UV) && S.tri_uv) {
    const float* q = S.tri_uv + (size_t)tri * 6;
    uv0 = mk2(ldg(q), ldg(q + 1)); uv1 = mk2(ldg(q + 2), ldg(q + 3)); uv2 = mk2(ldg(q + 4), ldg(q + 5));
  }
  f2 duv02 = uv0 - uv2, duv12 = uv1 - uv2;
  float det = duv02.x * duv12.y - duv02.y * duv12.x;
  f3 dpdu, dpdv;
  bool ok = false;
  if (!(fabsf(det) < 1e-8f)) {
    float invdet = 1.0f / det;
    dpdu = (duv12.y * dp02 - duv02.y * dp12) * invdet;
    dpdv = ((-duv12.x) * dp02 + duv02.x * dp12) * invdet;
    ok = !(len2(cross3(dpdu, dpdv)) <= 0.0f);
  }
  if (!ok) {
    f3 ng = cross3(p2 - p0, p1 - p0);
    coord_system(norm3(ng), &dpdu, &dpdv);
  }
  SurfD si;
  float xs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
  float ys = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
  float zs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
  si.perr = PT_GAMMA(7.0f) * mk3(xs, ys, zs);
  si.p = b0 * p0 + b1 * p1 + b2 * p2;
  si.wo = -rd;
  si.n = n;
  si.sn = n;
  si.sdpdu = dpdu;
  si.dpdu = dpdu;
  si.tri = tri;
  const bool has_n = (flags & PT_TRI_HAS_N) && S.tri_n, has_s = (flags & PT_TRI_HAS_S) && S.tri_s;
  if (has_n || has_s) {  // triangle.rs:376-444
    f3 ns = si.n;
    if (has_n) {
      const float* q = S.tri_n + (size_t)tri * 9;
      f3 n0 = mk3(ldg(q), ldg(q + 1), ldg(q + 2)), n1 = mk3(ldg(q + 3), ldg(q + 4), ldg(q + 5)), n2 = mk3(ldg(q + 6), ldg(q + 7), ldg(q + 8));
      f3 nns = b0 * n0 + b1 * n1 + b2 * n2;
      if (len2(nns) > 0.0f) ns = norm3(nns);
    }
    f3 ss = norm3(dpdu);
    if (has_s) {
      const float* q = S.tri_s + (size_t)tri * 9;
      f3 s0 = mk3(ldg(q), ldg(q + 1), ldg(q + 2)), s1 = mk3(ldg(q + 3), ldg(q + 4), ldg(q + 5)), s2 = mk3(ldg(q + 6), ldg(q + 7), ldg(q + 8));
      f3 nss = b0 * s0 + b1 * s1 + b2 * s2;
      if (len2(nss) > 0.0f) ss = norm3(nss);
    }
    f3 ts = cross3(ns, ss);
    if (len2(ts) > 0.0f) {
      ts = norm3(ts);
      ss = norm3(cross3(ts, ns));
    } else {
      coord_system(ns, &ss, &ts);
    }
    if (rev) ts = ts * -1.0f;
    si.sn = norm3(cross3(ss, ts));
    si.n = faceforward(si.n, si.sn);
    si.sdpdu = ss;
  }
  return si;
}

// Only the point and the geometric normal of tri_surface (what Shape::pdf_from and the area-light side test read), by the
// same operations; meshes with per-vertex normals or tangents, whose normal is flipped towards the shading normal, take
// the full path.
PT_HD void tri_point_normal(const DScene& S, int tri, f3 rd, float b0, float b1, float b2, f3* p_out, f3* n_out) {
  uint32_t flags = ldg(S.tri_flags + tri);
  if (((flags & PT_TRI_HAS_N) && S.tri_n) || ((flags & PT_TRI_HAS_S) && S.tri_s)) {
    SurfD s = tri_surface(S, tri, rd, b0, b1, b2);
    *p_out = s.p; *n_out = s.n;
    return;
  }
  f3 p0, p1, p2;
  tri_verts(S, tri, &p0, &p1, &p2);
  bool rev = (flags & PT_TRI_REVERSE_ORIENTATION) != 0, swp = (flags & PT_TRI_SWAPS_HANDEDNESS) != 0;
  f3 n = cross3(p0 - p2, p1 - p2);
  if (rev != swp) n = n * -1.0f;
  *n_out = norm3(n);
  *p_out = b0 * p0 + b1 * p1 + b2 * p2;
}

// Transform::transform_surface_interaction (instance -> world)
PT_HD1 SurfD surf_to_world(const InstanceRec& I, const SurfD& s) {
  SurfD r = s;
  const float* m = I.m;
  float x = s.p.x, y = s.p.y, z = s.p.z;
  r.p = mat_point(m, s.p);
  const float g3 = PT_GAMMA(3.0f);
  float ex = (g3 + 1.0f) * (fabsf(m[0]) * s.perr.x + fabsf(m[1]) * s.perr.y + fabsf(m[2]) * s.perr.z) +
             g3 * (fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z) + fabsf(m[3]));
  float ey = (g3 + 1.0f) * (fabsf(m[4]) * s.perr.x + fabsf(m[5]) * s.perr.y + fabsf(m[6]) * s.perr.z) +
             g3 * (fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z) + fabsf(m[7]));
  float ez = (g3 + 1.0f) * (fabsf(m[8]) * s.perr.x + fabsf(m[9]) * s.perr.y + fabsf(m[10]) * s.perr.z) +
             g3 * (fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z) + fabsf(m[11]));
  r.perr = mk3(ex, ey, ez);
  r.n = norm3(mat_normal_inv(I.minv, s.n));
  r.wo = norm3(mat_vector(m, s.wo));
  r.sn = norm3(mat_normal_inv(I.minv, s.sn));
  r.sdpdu = mat_vector(m, s.sdpdu);
  r.dpdu = mat_vector(m, s.dpdu);
  r.sn = faceforward(r.sn, r.n);
  return r;
}

// world-space interaction for a hit found by scene_intersect on world ray (ro, rd)
PT_HD SurfD make_surface(const DScene& S, int tri, int instance, f3 ro, f3 rd, float rtime, float b0, float b1, float b2) {
  if (instance < 0) return tri_surface(S, tri, rd, b0, b1, b2);
  const InstanceRec& I = S.instances[instance];
  DRay r;
  r.o = ro; r.d = rd; r.tmax = pt_inf(); r.time = rtime;
  DRay q = xform_ray(I.minv, r);
  SurfD s = tri_surface(S, tri, q.d, b0, b1, b2);
  return surf_to_world(I, s);
}

// ---- BSDF ---------------------------------------------------------------------------------------------------------
// at most two lobes per BSDF (plastic: diffuse + glossy; rough glass: reflection + transmission).  Loops over the lobes have a
// constant trip count and the lobe functions take the lobe by value, so that a BsdfD lives in registers: indexed through
// a pointer it sat in the local-memory frame (ncu, matte kernel: 5 % of the stall samples on the first load of B.lobe[i].kind)
// The five materials of the path's own classes need two; uber has up to five (uber.rs:85-125), translucent four: those run
// with BsdfT<PT_LOBES_EXTRA> in their own shade class (CL_EXTRA) and in the tree integrators' kernel.
#define PT_MAX_LOBES 2
#define PT_LOBES_EXTRA 5
#define PT_LOBE_LOOP(i) _Pragma("unroll") for (int i = 0; i < BS::ML; i++) if (i < B.n)
template <int MLOBES>
struct BsdfT {
  static constexpr int ML = MLOBES;
  f3 ns, ng, ss, ts;
  float eta;
  int n;       // number of lobes; -1 = no BSDF (ray passes through)
  Lobe lobe[MLOBES];
};
typedef BsdfT<PT_MAX_LOBES> BsdfD;
// lobe i of the BSDF by value (select chain: an indexed read would put the BsdfT into the local-memory frame)
template <class BS>
PT_HD Lobe bsdf_lobe(const BS& B, int index) {
  Lobe L = B.lobe[0];
#pragma unroll
  for (int k = 1; k < BS::ML; k++)
    if (index == k) L = B.lobe[k];
  return L;
}
template <class BS> PT_HD f3 to_local(const BS& B, f3 v) { return mk3(dot3(v, B.ss), dot3(v, B.ts), dot3(v, B.ns)); }
template <class BS> PT_HD f3 to_world(const BS& B, f3 v) {
  return mk3(B.ss.x * v.x + B.ts.x * v.y + B.ns.x * v.z, B.ss.y * v.x + B.ts.y * v.y + B.ns.y * v.z, B.ss.z * v.x + B.ts.z * v.y + B.ns.z * v.z);
}
PT_HD bool fin3(f3 v) { return isfinite_(v.x) && isfinite_(v.y) && isfinite_(v.z); }
template <int CM = PT_CM_ANY, class BS = BsdfD> PT_HD int bsdf_count(const BS& B, int flags) {
  int c = 0;
  PT_LOBE_LOOP(i) if (lobe_matches<CM>(B.lobe[i].kind, flags)) c++;
  return c;
}

template <int NC>
PT_HD bool spectrum_black(const DScene& S, int idx, bool clamp0) {
  Sp<NC> s = sp_load<NC>(S.spectra + (size_t)idx * NC);
  if (clamp0) s = sp_clamp0(s);
  return sp_black(s);
}

// Material::compute_scattering_functions for the five in-scope materials.
// Spectra that the reference clamps (clamp_zero) are clamped on the HOST at upload when a negative value is
// present, so the device can treat every stored reflectance as already clamped except matte Kd (never clamped).
// appends a lobe (the slot index is 0 or 1: two selects instead of an indexed store into the frame)
template <class BS> PT_HD void bsdf_add(BS* B, int& n, const Lobe& L) {
  if (BS::ML == 2) {
    if (n == 0) B->lobe[0] = L;
    else B->lobe[1] = L;
  } else {
#pragma unroll
    for (int k = 0; k < BS::ML; k++)
      if (n == k) B->lobe[k] = L;
  }
  n++;
}
// MULTI = allow_multiple_lobes (true: PathIntegrator, path.rs:106; false: directlighting / whitted)
template <int NC, int MT = -1, bool MULTI = true, class BS = BsdfD>
PT_HD1 void make_bsdf(const DScene& S, const SurfD& si, BS* B) {
  B->n = -1;
  int mi = ldg(S.tri_material + si.tri);
  if (mi < 0) return;
  const pt_material& M = S.materials[mi];
  float eta = 1.0f;
  int n = 0;
  switch (M.type) {
    case PT_MAT_MATTE: {
      if (MT != -1 && MT != PT_MAT_MATTE) { PT_UNREACHABLE(); }
      if (!spectrum_black<NC>(S, M.kd, false)) {
        Lobe L;
        L.sp0 = M.kd; L.sp1 = -1; L.fres = FR_NOOP; L.ei = L.et = 1.0f;
        float sig = clampf(M.sigma, 0.0f, 90.0f);
        if (sig == 0.0f) { L.kind = LB_LAMBERT; L.a = L.b = 0.0f; }
        else {
          L.kind = LB_OREN;
          float s = sig * (PT_PI / 180.0f);
          float s2 = s * s;
          L.a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
          L.b = 0.45f * s2 / (s2 + 0.09f);
        }
        bsdf_add(B, n, L);
      }
      break;
    }
    case PT_MAT_PLASTIC: {
      if (MT != -1 && MT != PT_MAT_PLASTIC) { PT_UNREACHABLE(); }
      if (!spectrum_black<NC>(S, M.kd, true)) {
        Lobe L;
        L.kind = LB_LAMBERT; L.sp0 = M.kd; L.sp1 = -1; L.fres = FR_NOOP; L.a = L.b = 0.0f; L.ei = L.et = 1.0f;
        bsdf_add(B, n, L);
      }
      if (!spectrum_black<NC>(S, M.ks, true)) {
        Lobe L;
        float a = ldg(S.alpha + 2 * mi);
        L.kind = LB_MF_REFL; L.sp0 = M.ks; L.sp1 = -1; L.fres = FR_DIEL; L.ei = 1.5f; L.et = 1.0f;
        L.a = fmaxf(0.001f, a); L.b = fmaxf(0.001f, a);
        bsdf_add(B, n, L);
      }
      break;
    }
    case PT_MAT_MIRROR: {
      if (MT != -1 && MT != PT_MAT_MIRROR) { PT_UNREACHABLE(); }
      if (!spectrum_black<NC>(S, M.kr, true)) {
        Lobe L;
        L.kind = LB_MIRROR; L.sp0 = M.kr; L.sp1 = -1; L.fres = FR_NOOP; L.a = L.b = 0.0f; L.ei = L.et = 1.0f;
        bsdf_add(B, n, L);
      }
      break;
    }
    case PT_MAT_GLASS: {
      if (MT != -1 && MT != PT_MAT_GLASS) { PT_UNREACHABLE(); }
      bool rb = spectrum_black<NC>(S, M.kr, false), tb = spectrum_black<NC>(S, M.ks, false);
      if (rb && tb) return;  // no BSDF at all
      eta = M.eta;
      if (M.uroughness == 0.0f && M.vroughness == 0.0f) {
        if (MULTI) {
          Lobe L;
          L.kind = LB_FRESNEL_SPEC; L.sp0 = M.kr; L.sp1 = M.ks; L.fres = FR_NOOP; L.a = L.b = 0.0f; L.ei = 1.0f; L.et = eta;
          bsdf_add(B, n, L);
        } else {
          if (!rb) {
            Lobe L;
            L.kind = LB_SPEC_REFL_D; L.sp0 = M.kr; L.sp1 = -1; L.fres = FR_DIEL; L.a = L.b = 0.0f; L.ei = 1.0f; L.et = eta;
            bsdf_add(B, n, L);
          }
          if (!tb) {
            Lobe L;
            L.kind = LB_SPEC_TRANS; L.sp0 = M.ks; L.sp1 = -1; L.fres = FR_NOOP; L.a = L.b = 0.0f; L.ei = 1.0f; L.et = eta;
            bsdf_add(B, n, L);
          }
        }
      } else {
        float ax = fmaxf(0.001f, ldg(S.alpha + 2 * mi)), ay = fmaxf(0.001f, ldg(S.alpha + 2 * mi + 1));
        if (!rb) {
          Lobe L;
          L.kind = LB_MF_REFL; L.sp0 = M.kr; L.sp1 = -1; L.fres = FR_DIEL; L.ei = 1.0f; L.et = eta; L.a = ax; L.b = ay;
          bsdf_add(B, n, L);
        }
        if (!tb) {
          Lobe L;
          L.kind = LB_MF_TRANS; L.sp0 = M.ks; L.sp1 = -1; L.fres = FR_NOOP; L.ei = 1.0f; L.et = eta; L.a = ax; L.b = ay;
          bsdf_add(B, n, L);
        }
      }
      break;
    }
    case PT_MAT_METAL: {
      if (MT != -1 && MT != PT_MAT_METAL) { PT_UNREACHABLE(); }
      Lobe L;
      L.kind = LB_MF_REFL; L.sp0 = M.kr; L.sp1 = M.ks; L.fres = FR_COND; L.ei = L.et = 1.0f;
      L.a = fmaxf(0.001f, ldg(S.alpha + 2 * mi)); L.b = fmaxf(0.001f, ldg(S.alpha + 2 * mi + 1));
      bsdf_add(B, n, L);
      break;
    }
    case PT_MAT_UBER: case PT_MAT_SUBSTRATE: case PT_MAT_TRANSLUCENT: case PT_MAT_MIX: {
      // uber.rs:85-125, substrate.rs:44-59, translucent.rs:48-98: every parameter is a constant, so the host has already
      // listed the lobes (device_api.cu material_lobes)
      if (MT != -1) { PT_UNREACHABLE(); }
      const MatLobes& T = S.mat_lobes[mi];
      if (T.n < 0) return;  // translucent with reflect = transmit = 0 leaves si.bsdf unset
      eta = T.eta;
#pragma unroll
      for (int k = 0; k < PT_LOBES_EXTRA; k++)
        if (k < T.n && k < BS::ML) {
          Lobe L;
          L.kind = T.kind[k]; L.sp0 = T.sp0[k]; L.sp1 = T.sp1[k]; L.fres = T.fres[k]; L.a = T.a[k]; L.b = T.b[k]; L.ei = T.ei[k]; L.et = T.et[k]; L.scale = T.scale[k];
          bsdf_add(B, n, L);
        }
      break;
    }
    default: return;
  }
  B->n = n;
  B->eta = eta;
  B->ns = si.sn;
  B->ng = si.n;
  B->ss = norm3(si.sdpdu);
  B->ts = norm3(cross3(B->ns, B->ss));
}

// ScaledBxDF::f / sample_f (reflection/scaled.rs:32-47): scale * f; pdf and type are the wrapped lobe's
template <int NC, class BS>
PT_HD Sp<NC> lobe_scaled(const DScene& S, const Lobe& L, Sp<NC> f) {
  if constexpr (BS::ML == PT_LOBES_EXTRA) {
    if (L.scale >= 0) return sp_load<NC>(S.spectra + (size_t)L.scale * NC) * f;
  }
  return f;
}
template <int NC, int CM = PT_CM_ANY, class BS = BsdfD>
PT_HD Sp<NC> bsdf_sum_f(const DScene& S, const BS& B, f3 wo, f3 wi, f3 wo_w, f3 wi_w, int flags) {
  bool refl = (dot3(wi_w, B.ng) * dot3(wo_w, B.ng)) > 0.0f;
  Sp<NC> r = sp_const<NC>(0.0f);
  PT_LOBE_LOOP(i) {
    int tp = lobe_type<CM>(B.lobe[i].kind);
    if (lobe_matches<CM>(B.lobe[i].kind, flags) && ((refl && (tp & F_REFL)) || (!refl && (tp & F_TRANS)))) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], lobe_f<NC, CM>(S, B.lobe[i], wo, wi));
  }
  return r;
}
template <int NC, int CM = PT_CM_ANY, class BS = BsdfD>
PT_HDN Sp<NC> bsdf_f(const DScene& S, const BS& B, f3 wo_w, f3 wi_w, int flags) {
  f3 wi = to_local(B, wi_w), wo = to_local(B, wo_w);
  if (wo.z == 0.0f || !fin3(wo)) return sp_const<NC>(0.0f);
  return bsdf_sum_f<NC, CM>(S, B, wo, wi, wo_w, wi_w, flags);
}
template <int CM = PT_CM_ANY, class BS = BsdfD>
PT_HDN float bsdf_pdf(const BS& B, f3 wo_w, f3 wi_w, int flags) {
  f3 wi = to_local(B, wi_w), wo = to_local(B, wo_w);
  if (wo.z == 0.0f || !fin3(wo)) return 0.0f;
  int count = 0;
  float p = 0.0f;
  PT_LOBE_LOOP(i)
    if (lobe_matches<CM>(B.lobe[i].kind, flags)) { p = p + lobe_pdf<CM>(B.lobe[i], wo, wi); count++; }
  return count > 0 ? p / (float)count : 0.0f;
}
// BSDF::f and BSDF::pdf for the same pair of directions (estimate_direct's light sample needs both, sample_lights.rs:
// 366-380): one pass over the lobes with lobe_f_pdf.  Same sums in the same lobe order as bsdf_f / bsdf_pdf.
template <int NC, int CM = PT_CM_ANY, class BS = BsdfD>
PT_HD1 void bsdf_f_pdf(const DScene& S, const BS& B, f3 wo_w, f3 wi_w, int flags, Sp<NC>* f_out, float* pdf_out) {
  *f_out = sp_const<NC>(0.0f);
  *pdf_out = 0.0f;
  f3 wi = to_local(B, wi_w), wo = to_local(B, wo_w);
  if (wo.z == 0.0f || !fin3(wo)) return;
  const bool refl = (dot3(wi_w, B.ng) * dot3(wo_w, B.ng)) > 0.0f;
  Sp<NC> r = sp_const<NC>(0.0f);
  int count = 0;
  float p = 0.0f;
  PT_LOBE_LOOP(i) {
    if (!lobe_matches<CM>(B.lobe[i].kind, flags)) continue;
    const int tp = lobe_type<CM>(B.lobe[i].kind);
    const bool want_f = (refl && (tp & F_REFL)) || (!refl && (tp & F_TRANS));
    if (((CM >> LB_MF_REFL) & 1) && B.lobe[i].kind == LB_MF_REFL) {
      Sp<NC> fl;
      float pl;
      lobe_f_pdf<NC, CM>(S, B.lobe[i], wo, wi, want_f, &fl, &pl);
      if (want_f) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], fl);
      p = p + pl;
    } else {  // nothing to share: no spectrum-sized temporary (matters for 60 bins)
      if (want_f) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], lobe_f<NC, CM>(S, B.lobe[i], wo, wi));
      p = p + lobe_pdf<CM>(B.lobe[i], wo, wi);
    }
    count++;
  }
  *f_out = r;
  *pdf_out = count > 0 ? p / (float)count : 0.0f;
}

template <int NC, int CM = PT_CM_ANY, class BS = BsdfD>
// (inlined at its two call sites, the MIS sample and the continuation sample: +5 % code, but the BsdfD and the four result
// slots leave the local-memory frame: k_shade -2 % on config 2, -3 % on the instanced scene, profiles/r2f_ab_sync_and_shade.txt)
PT_HD bool bsdf_sample(const DScene& S, const BS& B, f3 wo_w, f2 u, int flags, Sp<NC>* f, f3* wi_w_out, float* pdf_out, int* stype) {
  int matching = bsdf_count<CM>(B, flags);
  if (matching == 0) return false;
  int comp = (int)floorf(u.x * (float)matching);
  if (comp > matching - 1) comp = matching - 1;
  int index = -1, cnt = comp;
  PT_LOBE_LOOP(i)
    if (lobe_matches<CM>(B.lobe[i].kind, flags)) { if (cnt == 0) { index = i; break; } cnt--; }
  const Lobe L = BS::ML == 2 ? (index == 0 ? B.lobe[0] : B.lobe[1]) : bsdf_lobe(B, index);
  f2 ur = mk2(fminf((u.x * (float)matching) - (float)comp, PT_ONE_MINUS_EPS), u.y);
  f3 wo = to_local(B, wo_w);
  if (wo.z == 0.0f || !fin3(wo)) return false;
  Sp<NC> fv;
  f3 wi;
  float pdf;
  int t;
  if (!lobe_sample<NC, CM>(S, L, wo, ur, &fv, &wi, &pdf, &t)) return false;
  fv = lobe_scaled<NC, BS>(S, L, fv);
  if (pdf <= 0.0f) return false;
  int ltype = lobe_type<CM>(L.kind);
  int sampled = t != 0 ? t : ltype;
  f3 wi_w = to_world(B, wi);
  if (NC != 3 && (ltype & F_SPEC) == 0) {  // 60 bins: the plain form (no spectrum-sized temporaries)
    if (matching > 1)
      PT_LOBE_LOOP(i)
        if (i != index && lobe_matches<CM>(B.lobe[i].kind, flags)) pdf += lobe_pdf<CM>(B.lobe[i], wo, wi);
    fv = bsdf_sum_f<NC, CM>(S, B, wo, wi, wo_w, wi_w, flags);
  } else if ((ltype & F_SPEC) == 0) {
    // pdf += the other matching lobes' pdfs (bsdf.rs:167-176); f = sum over the matching lobes on the right side of the
    // geometric normal (bsdf.rs:178-196), the sampled lobe included: one fused evaluation per other lobe, both in lobe order
    const bool refl = (dot3(wi_w, B.ng) * dot3(wo_w, B.ng)) > 0.0f;
    Sp<NC> r = sp_const<NC>(0.0f);
    PT_LOBE_LOOP(i) {
      if (!lobe_matches<CM>(B.lobe[i].kind, flags)) continue;
      const int tp = lobe_type<CM>(B.lobe[i].kind);
      const bool want_f = (refl && (tp & F_REFL)) || (!refl && (tp & F_TRANS));
      if (i == index) {
        if (want_f) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], lobe_f<NC, CM>(S, B.lobe[i], wo, wi));
      } else if (((CM >> LB_MF_REFL) & 1) && B.lobe[i].kind == LB_MF_REFL) {
        Sp<NC> fl;
        float pl;
        lobe_f_pdf<NC, CM>(S, B.lobe[i], wo, wi, want_f, &fl, &pl);
        if (want_f) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], fl);
        pdf += pl;
      } else {
        if (want_f) r = r + lobe_scaled<NC, BS>(S, B.lobe[i], lobe_f<NC, CM>(S, B.lobe[i], wo, wi));
        pdf += lobe_pdf<CM>(B.lobe[i], wo, wi);
      }
    }
    fv = r;
  }
  if (matching > 1) pdf /= (float)matching;
  *f = fv; *wi_w_out = wi_w; *pdf_out = pdf; *stype = sampled;
  return true;
}

}  // namespace ptd
