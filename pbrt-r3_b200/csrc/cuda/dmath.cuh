// dmath.cuh — device math for the wavefront path tracer (f32, no FMA contraction: built with -fmad=false).
// All functions are PT_HD so that tests/hostcheck can also compile them with g++ as a logic check
// (never as a product path: the library's entry points need the CUDA kernels).
// Numerics follow reference src/core/geometry/{vector3,misc}.rs, src/core/misc/float.rs,
// src/core/base/{constants,functions}.rs.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef PT_INLINE_SINGLE_SITE
#define PT_INLINE_SINGLE_SITE 1
#endif
#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__
/* large stage functions: ONE out-of-line copy each keeps the shade kernel inside the instruction cache (ncu: the
   fully inlined kernel was 34K instructions and stalled on no_instruction for 19-63 cycles per issue) */
#define PT_HDN inline __host__ __device__ __noinline__
/* stage functions with ONE call site per shade kernel: inlining them does not grow the code, and their argument / result
   structures (SurfD, LightSampleD, RefPt) stay in registers instead of the local-memory stack frame (ncu: 41 % of the
   local loads of the shade kernels missed L1, the frames of 16 warps being as large as the cache) */
#if PT_INLINE_SINGLE_SITE
#define PT_HD1 __host__ __device__ __forceinline__
#else
#define PT_HD1 inline __host__ __device__ __noinline__
#endif
#else
#define PT_HD inline
#define PT_D inline
#define PT_HDN inline
#define PT_HD1 inline
#endif

namespace ptd {

struct alignas(16) f4 {
  float x, y, z, w;
};
struct alignas(16) i4 {
  int x, y, z, w;
};
struct f2 {
  float x, y;
};
struct f3 {
  float x, y, z;
};

PT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
PT_HD f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PT_HD f3 xyz(const f4& v) { return mk3(v.x, v.y, v.z); }
PT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PT_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PT_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PT_HD f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
PT_HD f2 operator-(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
PT_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float absdot3(f3 a, f3 b) { return fabsf(dot3(a, b)); }
PT_HD float len2(f3 a) { return dot3(a, a); }
PT_HD float len3(f3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
PT_HD f3 norm3(f3 a) { float l = len3(a); return mk3(a.x / l, a.y / l, a.z / l); }
PT_HD f3 cross3(f3 a, f3 b) { return mk3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
PT_HD f3 abs3(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
PT_HD float dist2(f3 a, f3 b) { f3 v = a - b; return dot3(v, v); }
PT_HD float comp(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
PT_HD f3 faceforward(f3 n, f3 v) { return dot3(n, v) < 0.0f ? n * -1.0f : n; }

#define PT_PI 3.14159265358979323846f
#define PT_INV_PI 0.31830988618379067154f
#define PT_PI_OVER_2 (PT_PI / 2.0f)
#define PT_PI_OVER_4 (PT_PI / 4.0f)
#define PT_ONE_MINUS_EPS 0.99999994f
#define PT_SHADOW_EPS 0.0001f
#define PT_MACH_EPS 5.9604644775390625e-8f /* f32::EPSILON * 0.5 */
#define PT_GAMMA(n) (((n)*PT_MACH_EPS) / (1.0f - (n)*PT_MACH_EPS))
#define PT_GAMMA6_REF ((6.0f * PT_MACH_EPS) / (1.0f - 5.0f * PT_MACH_EPS)) /* shapes/triangle.rs:34 (sic) */

PT_HD float pt_inf() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7f800000);
#else
  return INFINITY;
#endif
}
PT_HD uint32_t fbits(float f) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PT_HD float bitsf(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((int)u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PT_HD bool isinf_(float v) { return (fbits(v) & 0x7fffffffu) == 0x7f800000u; }
PT_HD bool isfinite_(float v) { return (fbits(v) & 0x7f800000u) != 0x7f800000u; }
PT_HD bool signbit_(float v) { return (fbits(v) >> 31) != 0; }
// IEEE division whatever -prec-div says: the shade translation unit is compiled with approximate division / sqrt, but
// sampler values and the voxel of the light distribution must stay bit-exact (SURVEY 8d parity gates)
PT_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
PT_HD int popc_(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
PT_HD int ffs_(unsigned v) {  // 1-based index of the lowest set bit, 0 for v == 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}
PT_HD float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }  // f32::clamp
PT_HD float lerpf(float t, float a, float b) { return (1.0f - t) * a + t * b; }

// sin / cos with the host libm's results.  The reference calls f32::sin / f32::cos, i.e. the platform libm; on
// Linux/glibc (>= 2.28) that is the published Arm "optimized routines" sincosf algorithm: argument reduction and
// a degree-7/8 polynomial evaluated in f64, rounded once to f32.  CUDA's sinf/cosf differ from it in the last ulp
// often enough that specular chains amplify the difference into visible per-pixel noise (RMSE > 1e-3 on the
// glass/mirror Cornell test), so the same algorithm is evaluated here (f64, no FMA).  Checked bit-for-bit against
// glibc sinf/cosf on 2e8 random arguments in [-7, 7] (tools/check_sincos.cpp).  Valid for |x| < 120, which
// covers every call site of the path (theta in [-pi/4, 3pi/4], phi in [0, 2pi)).
PT_HD float pt_sincos_poly(double x, double x2, bool neg_cos, int n) {
  if ((n & 1) == 0) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x3 = x * x2;
    double t = s2 + x2 * s3;
    double x7 = x3 * x2;
    double s = x + x3 * s1;
    return (float)(s + x7 * t);
  }
  const double sg = neg_cos ? -1.0 : 1.0;
  const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10,
               c4 = sg * 0x1.99343027bf8c3p-16;
  double x4 = x2 * x2;
  double t2 = c3 + x2 * c4;
  double t1 = c1 + x2 * c2;
  double x6 = x4 * x2;
  double c = c0 + x2 * t1;
  return (float)(c + x6 * t2);
}
PT_HD float pt_sincos(float y, int is_cos) {
  double x = (double)y;
  uint32_t top = (fbits(y) >> 20) & 0x7ffu;
  if (top < 0x3f4u) {  // |y| < pi/4
    if (top < 0x398u) return is_cos ? 1.0f : y;  // |y| < 2^-12
    return pt_sincos_poly(x, x * x, false, is_cos);
  }
  double r = x * 0x1.45F306DC9C883p+23;
  int n = ((int)r + 0x800000) >> 24;
  x = x - (double)n * 0x1.921FB54442D18p0;
  double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return pt_sincos_poly(x * s, x * x, (n & 2) != 0, n ^ is_cos);
}
PT_HD float pt_sinf(float y) { return pt_sincos(y, 0); }
PT_HD float pt_cosf(float y) { return pt_sincos(y, 1); }

// core/misc/float.rs:23-56
PT_HD float next_up(float v) {
  if (isinf_(v) && v > 0.0f) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = fbits(v);
  if (v >= 0.0f) ui = ui == 0xffffffffu ? ui : ui + 1;
  else ui = ui == 0 ? 0 : ui - 1;
  return bitsf(ui);
}
PT_HD float next_down(float v) {
  if (isinf_(v) && v < 0.0f) return v;
  if (v == 0.0f) v = -0.0f;
  uint32_t ui = fbits(v);
  if (v > 0.0f) ui = ui == 0 ? 0 : ui - 1;
  else ui = ui == 0xffffffffu ? ui : ui + 1;
  return bitsf(ui);
}
// core/geometry/misc.rs:5-23
PT_HD f3 offset_origin(f3 p, f3 perr, f3 n, f3 w) {
  float d = dot3(abs3(n), perr);
  f3 off = d * n;
  if (dot3(w, n) < 0.0f) off = -off;
  f3 po = p + off;
  if (off.x > 0.0f) po.x = next_up(po.x); else if (off.x < 0.0f) po.x = next_down(po.x);
  if (off.y > 0.0f) po.y = next_up(po.y); else if (off.y < 0.0f) po.y = next_down(po.y);
  if (off.z > 0.0f) po.z = next_up(po.z); else if (off.z < 0.0f) po.z = next_down(po.z);
  return po;
}
// core/geometry/misc.rs:62-71
PT_HD void coord_system(f3 v1, f3* v2, f3* v3) {
  if (fabsf(v1.x) > fabsf(v1.y)) *v2 = mk3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
  else *v2 = mk3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
  *v3 = norm3(cross3(v1, *v2));
}
PT_HD int max_dim(f3 v) {  // misc.rs:35-50
  if (v.x > v.y) return v.x > v.z ? 0 : 2;
  return v.y > v.z ? 1 : 2;
}

// 4x4 row-major matrix application (core/transform/matrix4x4.rs:284-317)
PT_HD f3 mat_point(const float* m, f3 p) {
  float x = p.x, y = p.y, z = p.z;
  float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
  float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
  float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
  float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
  if (wp == 1.0f) return mk3(xp, yp, zp);
  return mk3(xp / wp, yp / wp, zp / wp);
}
PT_HD f3 mat_vector(const float* m, f3 p) {
  float x = p.x, y = p.y, z = p.z;
  return mk3(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
PT_HD f3 mat_normal_inv(const float* minv, f3 p) {
  float x = p.x, y = p.y, z = p.z;
  return mk3(minv[0] * x + minv[4] * y + minv[8] * z, minv[1] * x + minv[5] * y + minv[9] * z, minv[2] * x + minv[6] * y + minv[10] * z);
}

struct DRay {
  f3 o, d;
  float tmax, time;
};
// Transform::transform_ray (core/transform/transform.rs:184-203, 270-282)
PT_HD DRay xform_ray(const float* m, const DRay& r) {
  float x = r.o.x, y = r.o.y, z = r.o.z;
  f3 o = mat_point(m, r.o);
  float xs = fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z) + fabsf(m[3]);
  float ys = fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z) + fabsf(m[7]);
  float zs = fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z) + fabsf(m[11]);
  f3 oerr = PT_GAMMA(3.0f) * mk3(xs, ys, zs);
  f3 d = mat_vector(m, r.d);
  float ls = len2(d);
  if (ls > 0.0f) {
    float dt = dot3(abs3(d), oerr) / ls;
    o = o + d * dt;
  }
  DRay q;
  q.o = o; q.d = d; q.tmax = r.tmax; q.time = r.time;
  return q;
}

// global loads through the read-only path on device
template <typename T>
PT_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
PT_HD f4 ldg4(const f4* p) {
#if defined(__CUDA_ARCH__)
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return mk4(v.x, v.y, v.z, v.w);
#else
  return *p;
#endif
}
PT_HD i4 ldgi4(const i4* p) {
#if defined(__CUDA_ARCH__)
  int4 v = __ldg(reinterpret_cast<const int4*>(p));
  i4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
  return r;
#else
  return *p;
#endif
}

}  // namespace ptd
