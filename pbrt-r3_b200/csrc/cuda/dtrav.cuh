// dtrav.cuh — watertight ray/triangle test, world-bound slab test and the shared traversal definitions (the 4-wide
// traversal itself is the resumable state machine of dtrav2.cuh).
// Replaces reference src/shapes/triangle.rs:226-352,452-577 (Triangle::intersect / intersect_p decision part),
// src/accelerators/bvh/accel/qbvh/qbvh_x86.rs:26-67,187-343,373-399 (test_aabb, ORDER_TABLE, intersect_simd[_p],
// QBVHAccel::intersect[_p]), src/core/geometry/intersect.rs:46-65 (world-bound slab test),
// src/core/primitive/transformed_primitive.rs:26-46.
// Traversal order, tie-breaking and every float operation follow the reference so that primitive ids match
// bit for bit: separate sub and mul (no FMA), SSE max/min operand order, far-to-near push, leaf order.
#pragma once
#include "dscene.cuh"

namespace ptd {

struct TriHitD {
  float t, b0, b1, b2;
};

PT_HD bool tri_test(f3 p0, f3 p1, f3 p2, uint32_t flags, f3 ro, f3 rd, float tmax, TriHitD* out) {
  f3 dp02 = p0 - p2, dp12 = p1 - p2;
  if (!(flags & PT_TRI_TWO_SIDED)) {
    f3 n = cross3(dp02, dp12);
    if (((flags & PT_TRI_REVERSE_ORIENTATION) != 0) != ((flags & PT_TRI_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    if (dot3(n, rd) >= 0.0f) return false;
  }
  f3 a = p0 - ro, b = p1 - ro, c = p2 - ro;
  int kz = max_dim(abs3(rd));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float dx = comp(rd, kx), dy = comp(rd, ky), dz = comp(rd, kz);
  float ax = comp(a, kx), ay = comp(a, ky), az = comp(a, kz);
  float bx = comp(b, kx), by = comp(b, ky), bz = comp(b, kz);
  float cx = comp(c, kx), cy = comp(c, ky), cz = comp(c, kz);
  float sx = -dx / dz, sy = -dy / dz, sz = 1.0f / dz;
  ax += sx * az; ay += sy * az;
  bx += sx * bz; by += sy * bz;
  cx += sx * cz; cy += sy * cz;
  float e0 = bx * cy - by * cx;
  float e1 = cx * ay - cy * ax;
  float e2 = ax * by - ay * bx;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)((double)cy * (double)bx - (double)cx * (double)by);
    e1 = (float)((double)ay * (double)cx - (double)ax * (double)cy);
    e2 = (float)((double)by * (double)ax - (double)bx * (double)ay);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  az *= sz; bz *= sz; cz *= sz;
  float ts = e0 * az + e1 * bz + e2 * cz;
  if (det < 0.0f && (ts >= 0.0f || ts < tmax * det)) return false;
  else if (det > 0.0f && (ts <= 0.0f || ts > tmax * det)) return false;
  float inv_det = 1.0f / det;
  float t = ts * inv_det;
  float max_zt = fmaxf(fabsf(az), fmaxf(fabsf(bz), fabsf(cz)));
  float delta_z = PT_GAMMA(3.0f) * max_zt;
  float max_xt = fmaxf(fabsf(ax), fmaxf(fabsf(bx), fabsf(cx)));
  float max_yt = fmaxf(fabsf(ay), fmaxf(fabsf(by), fabsf(cy)));
  float delta_x = PT_GAMMA(5.0f) * (max_xt + max_zt);
  float delta_y = PT_GAMMA(5.0f) * (max_yt + max_zt);
  float delta_e = 2.0f * (PT_GAMMA(2.0f) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
  float max_e = fmaxf(fabsf(e0), fmaxf(fabsf(e1), fabsf(e2)));
  float delta_t = 3.0f * (PT_GAMMA(3.0f) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
  if (t <= delta_t) return false;
  out->t = t;
  out->b0 = e0 * inv_det;
  out->b1 = e1 * inv_det;
  out->b2 = e2 * inv_det;
  return true;
}

// little-endian nibble lists, terminated by nibble 4 (qbvh_x86.rs:187-204); index = hit_mask * 8 + sign bits
#if defined(__CUDACC__)
__device__ __constant__
#endif
static const uint32_t kOrderTable[128] = {
    0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44444, 0x44440, 0x44440, 0x44440, 0x44440, 0x44440, 0x44440,
    0x44440, 0x44440, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44441, 0x44401, 0x44401, 0x44410, 0x44410,
    0x44401, 0x44401, 0x44410, 0x44410, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44442, 0x44402, 0x44402,
    0x44402, 0x44402, 0x44420, 0x44420, 0x44420, 0x44420, 0x44412, 0x44412, 0x44412, 0x44412, 0x44421, 0x44421, 0x44421, 0x44421,
    0x44012, 0x44012, 0x44102, 0x44102, 0x44201, 0x44201, 0x44210, 0x44210, 0x44443, 0x44443, 0x44443, 0x44443, 0x44443, 0x44443,
    0x44443, 0x44443, 0x44403, 0x44403, 0x44403, 0x44403, 0x44430, 0x44430, 0x44430, 0x44430, 0x44413, 0x44413, 0x44413, 0x44413,
    0x44431, 0x44431, 0x44431, 0x44431, 0x44013, 0x44013, 0x44103, 0x44103, 0x44301, 0x44301, 0x44310, 0x44310, 0x44423, 0x44432,
    0x44423, 0x44432, 0x44423, 0x44432, 0x44423, 0x44432, 0x44023, 0x44032, 0x44023, 0x44032, 0x44230, 0x44320, 0x44230, 0x44320,
    0x44123, 0x44132, 0x44123, 0x44132, 0x44231, 0x44321, 0x44231, 0x44321, 0x40123, 0x40132, 0x41023, 0x41032, 0x42301, 0x43201,
    0x42310, 0x43210};

struct HitD {
  int tri;       // global triangle id, -1 = miss
  int instance;  // -1 = top level
  float t, b0, b1, b2;
};

struct TravCounters {
  unsigned long long nodes, tris;
};

// world-bound slab test with f32::max / f32::min semantics (intersect.rs:46-65, bounds3.rs:154-163)
PT_HD bool bounds_range(const float* bmin, const float* bmax, f3 o, f3 d, float tmax, float* t0o, float* t1o) {
  float t0 = 0.0f, t1 = tmax;
  float idx = 1.0f / d.x, idy = 1.0f / d.y, idz = 1.0f / d.z;
  bool nx = signbit_(d.x), ny = signbit_(d.y), nz = signbit_(d.z);
  t0 = fmaxf(t0, ((nx ? bmax[0] : bmin[0]) - o.x) * idx);
  t1 = fminf(t1, ((nx ? bmin[0] : bmax[0]) - o.x) * idx);
  t0 = fmaxf(t0, ((ny ? bmax[1] : bmin[1]) - o.y) * idy);
  t1 = fminf(t1, ((ny ? bmin[1] : bmax[1]) - o.y) * idy);
  t0 = fmaxf(t0, ((nz ? bmax[2] : bmin[2]) - o.z) * idz);
  t1 = fminf(t1, ((nz ? bmin[2] : bmax[2]) - o.z) * idz);
  *t0o = t0;
  *t1o = t1;
  return t0 <= t1;
}

#define PT_SSE_MAX(a, b) ((a) > (b) ? (a) : (b))
#define PT_SSE_MIN(a, b) ((a) < (b) ? (a) : (b))

#define PT_STACK_BARRIER ((int)0x7fffffff)

}  // namespace ptd
