// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of pbrt-r3's PathIntegrator hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything under oracle/.  The product (pbrt-r3_b200/) never does.
//
// PARITY UNPINNED at the image level: the reference is Rust and cannot be built here (no
// rustc/cargo), and it ships no golden images.  The leaf arithmetic is pinned against the
// reference's own known-answer tests (tests/test_oracle_known_answers.py).
//
// o_math.h — f32 vectors, bounds, matrices, float bit tricks.
// Follows reference src/core/geometry/{vector3,bounds3,misc}.rs, src/core/transform/matrix4x4.rs,
// src/core/misc/float.rs, src/core/base/{constants,functions}.rs.
// Build with -ffp-contract=off (Rust never fuses a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

typedef float Float;

// core/base/constants.rs:13-32
static const Float SHADOW_EPSILON = 0.0001f;
static const Float PI = 3.14159265358979323846f;
static const Float INV_PI = 0.31830988618379067154f;
static const Float INV_2_PI = INV_PI * 0.5f;
static const Float INV_4_PI = INV_PI * 0.25f;
static const Float PI_OVER_2 = PI / 2.0f;
static const Float PI_OVER_4 = PI / 4.0f;
static const Float ONE_MINUS_EPSILON = 0.99999994f;
static const Float INF = std::numeric_limits<float>::infinity();
static const Float MACHINE_EPSILON = std::numeric_limits<float>::epsilon() * 0.5f;

// core/base/functions.rs:3-10 ; shapes/triangle.rs:30-35 (note GAMMA6's typo is the reference's)
inline constexpr Float gamma_n(int n) { return (n * (std::numeric_limits<float>::epsilon() * 0.5f)) / (1.0f - n * (std::numeric_limits<float>::epsilon() * 0.5f)); }
static const Float GAMMA2 = (2.0f * MACHINE_EPSILON) / (1.0f - (2.0f * MACHINE_EPSILON));
static const Float GAMMA3 = (3.0f * MACHINE_EPSILON) / (1.0f - (3.0f * MACHINE_EPSILON));
static const Float GAMMA5 = (5.0f * MACHINE_EPSILON) / (1.0f - (5.0f * MACHINE_EPSILON));
static const Float GAMMA6 = (6.0f * MACHINE_EPSILON) / (1.0f - (5.0f * MACHINE_EPSILON));
static const Float GAMMA7 = (7.0f * MACHINE_EPSILON) / (1.0f - (7.0f * MACHINE_EPSILON));

// Rust semantics helpers -------------------------------------------------------------
inline Float rmax(Float a, Float b) { return fmaxf(a, b); }   // f32::max (NaN-ignoring)
inline Float rmin(Float a, Float b) { return fminf(a, b); }   // f32::min
inline Float rclamp(Float x, Float lo, Float hi) {            // f32::clamp
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}
inline int32_t f2i(Float v) {                                  // `as i32` (saturating, NaN -> 0)
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}
inline uint32_t f2u(Float v) {                                 // `as u32`
  if (v != v) return 0;
  if (v >= 4294967296.0f) return UINT32_MAX;
  if (v <= 0.0f) return 0;
  return (uint32_t)v;
}
inline Float lerp(Float t, Float v1, Float v2) { return (1.0f - t) * v1 + t * v2; }  // functions.rs:31

inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// core/misc/float.rs:23-56
inline float next_float_down(float x) {
  float v = x;
  if (std::isinf(v) && v < 0.0f) return v;
  if (v == 0.0f) v = -0.0f;
  uint32_t ui = f2bits(v);
  if (v > 0.0f) ui = ui == 0 ? 0 : ui - 1;
  else ui = ui == UINT32_MAX ? ui : ui + 1;
  return bits2f(ui);
}
inline float next_float_up(float x) {
  float v = x;
  if (std::isinf(v) && v > 0.0f) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = f2bits(v);
  if (v >= 0.0f) ui = ui == UINT32_MAX ? ui : ui + 1;
  else ui = ui == 0 ? 0 : ui - 1;
  return bits2f(ui);
}

// core/geometry/vector2.rs, vector3.rs ---------------------------------------------------
struct Vec2 {
  Float x, y;
  Vec2() : x(0), y(0) {}
  Vec2(Float x_, Float y_) : x(x_), y(y_) {}
  Float operator[](int i) const { return i == 0 ? x : y; }
};
inline Vec2 operator+(Vec2 a, Vec2 b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(Vec2 a, Vec2 b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(Vec2 a, Float s) { return Vec2(a.x * s, a.y * s); }
inline Vec2 operator*(Float s, Vec2 a) { return Vec2(s * a.x, s * a.y); }

struct Vec3 {
  Float x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
  Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(Vec3 a, Float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(Float s, Vec3 a) { return Vec3(s * a.x, s * a.y, s * a.z); }
inline Vec3 operator/(Vec3 a, Float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline Float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }           // vector3.rs:98
inline Float abs_dot(Vec3 a, Vec3 b) { return fabsf(dot(a, b)); }
inline Float length_squared(Vec3 a) { return dot(a, a); }
inline Float length(Vec3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }           // :108
inline Vec3 normalize(Vec3 a) { Float l = length(a); return Vec3(a.x / l, a.y / l, a.z / l); }  // :113-121
inline Vec3 cross(Vec3 a, Vec3 b) {                                                          // :134-142 (f32)
  return Vec3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
inline Vec3 vabs(Vec3 a) { return Vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
inline Float distance_squared(Vec3 a, Vec3 b) { Vec3 v = a - b; return dot(v, v); }

// core/geometry/misc.rs -------------------------------------------------------------------
inline Vec3 face_forward(Vec3 n, Vec3 v) { return dot(n, v) < 0.0f ? n * -1.0f : n; }        // :26-32
inline int max_dimension(Vec3 v) {                                                            // :35-50
  if (v.x > v.y) return v.x > v.z ? 0 : 2;
  return v.y > v.z ? 1 : 2;
}
inline Vec3 permute(Vec3 v, int x, int y, int z) { return Vec3(v[x], v[y], v[z]); }
inline Float max_component(Vec3 v) { return rmax(v.x, rmax(v.y, v.z)); }
inline void coordinate_system(Vec3 v1, Vec3* v2, Vec3* v3) {                                  // :62-71
  if (fabsf(v1.x) > fabsf(v1.y)) *v2 = Vec3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
  else *v2 = Vec3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
  *v3 = normalize(cross(v1, *v2));
}
inline Vec3 offset_ray_origin(Vec3 p, Vec3 p_error, Vec3 n, Vec3 w) {                         // :5-23
  Float d = dot(vabs(n), p_error);
  Vec3 offset = d * n;
  if (dot(w, n) < 0.0f) offset = -offset;
  Vec3 po = p + offset;
  Float a[3] = {po.x, po.y, po.z};
  for (int i = 0; i < 3; i++) {
    if (offset[i] > 0.0f) a[i] = next_float_up(a[i]);
    else if (offset[i] < 0.0f) a[i] = next_float_down(a[i]);
  }
  return Vec3(a[0], a[1], a[2]);
}

// core/geometry/bounds3.rs ---------------------------------------------------------------
inline Float bmin_(Float a, Float b) { return a <= b ? a : b; }   // :38-44 (PartialOrd forms)
inline Float bmax_(Float a, Float b) { return a >= b ? a : b; }
struct Bounds3 {
  Vec3 mn, mx;
  Bounds3() : mn(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()),
              mx(std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()) {}  // :224-231
  Bounds3(Vec3 a, Vec3 b) : mn(bmin_(a.x, b.x), bmin_(a.y, b.y), bmin_(a.z, b.z)), mx(bmax_(a.x, b.x), bmax_(a.y, b.y), bmax_(a.z, b.z)) {}
  static Bounds3 raw(Vec3 a, Vec3 b) { Bounds3 r; r.mn = a; r.mx = b; return r; }
  Vec3 diagonal() const { return mx - mn; }
  int maximum_extent() const {                                     // :61-70
    Vec3 d = diagonal();
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
  }
  Vec3 offset(Vec3 p) const {                                      // :72-84
    Vec3 o = p - mn;
    if (mx.x > mn.x) o.x = o.x / (mx.x - mn.x);
    if (mx.y > mn.y) o.y = o.y / (mx.y - mn.y);
    if (mx.z > mn.z) o.z = o.z / (mx.z - mn.z);
    return o;
  }
  Float surface_area() const { Vec3 d = diagonal(); return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }  // :196-199
  Vec3 lerp3(Vec3 t) const { return Vec3(lerp(t.x, mn.x, mx.x), lerp(t.y, mn.y, mx.y), lerp(t.z, mn.z, mx.z)); }  // :176-182
};
inline Bounds3 bunion(const Bounds3& a, const Bounds3& b) {       // :94-106
  return Bounds3::raw(Vec3(bmin_(a.mn.x, b.mn.x), bmin_(a.mn.y, b.mn.y), bmin_(a.mn.z, b.mn.z)),
                      Vec3(bmax_(a.mx.x, b.mx.x), bmax_(a.mx.y, b.mx.y), bmax_(a.mx.z, b.mx.z)));
}
inline Bounds3 bunion_p(const Bounds3& a, Vec3 p) {                // :122-134
  return Bounds3::raw(Vec3(bmin_(a.mn.x, p.x), bmin_(a.mn.y, p.y), bmin_(a.mn.z, p.z)),
                      Vec3(bmax_(a.mx.x, p.x), bmax_(a.mx.y, p.y), bmax_(a.mx.z, p.z)));
}

// core/transform/matrix4x4.rs (only what the hot path touches at run time) ---------------
struct Mat4 {
  Float m[16];
  Vec3 point(Vec3 p) const {                                      // :284-297
    Float x = p.x, y = p.y, z = p.z;
    Float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    Float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    Float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    Float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.0f) return Vec3(xp, yp, zp);
    return Vec3(xp / wp, yp / wp, zp / wp);
  }
  Vec3 vector(Vec3 p) const {                                     // :299-307
    Float x = p.x, y = p.y, z = p.z;
    return Vec3(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
  }
  Vec3 normal_with(Vec3 p) const {                                // :309-317 (caller passes minv)
    Float x = p.x, y = p.y, z = p.z;
    return Vec3(m[0] * x + m[4] * y + m[8] * z, m[1] * x + m[5] * y + m[9] * z, m[2] * x + m[6] * y + m[10] * z);
  }
};

struct Ray {
  Vec3 o, d;
  mutable Float t_max;   // Cell<Float> in the reference (core/geometry/ray/ray.rs:7-16)
  Float time;
  Ray() : t_max(INF), time(0) {}
  Ray(Vec3 o_, Vec3 d_, Float tm, Float ti) : o(o_), d(d_), t_max(tm), time(ti) {}
};

// core/transform/transform.rs:270-282 (+ :184-203, :240-268)
inline Ray transform_ray(const Mat4& M, const Ray& r) {
  Float x = r.o.x, y = r.o.y, z = r.o.z;
  Vec3 o = M.point(r.o);
  Float xs = fabsf(M.m[0] * x) + fabsf(M.m[1] * y) + fabsf(M.m[2] * z) + fabsf(M.m[3]);
  Float ys = fabsf(M.m[4] * x) + fabsf(M.m[5] * y) + fabsf(M.m[6] * z) + fabsf(M.m[7]);
  Float zs = fabsf(M.m[8] * x) + fabsf(M.m[9] * y) + fabsf(M.m[10] * z) + fabsf(M.m[11]);
  Vec3 o_error = GAMMA3 * Vec3(xs, ys, zs);
  Vec3 d = M.vector(r.d);
  Float ls = length_squared(d);
  if (ls > 0.0f) {
    Float dt = dot(vabs(d), o_error) / ls;
    o = o + d * dt;
  }
  return Ray(o, d, r.t_max, r.time);
}

}  // namespace orc
