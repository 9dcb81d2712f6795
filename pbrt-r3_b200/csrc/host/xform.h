// xform.h — host-side 4x4 matrices and Transform pairs for the scene front end.
// Mirrors the conventions of reference src/core/transform/{matrix4x4,transform}.rs:
// row-major m[4*r+c], column vectors, Transform{m,minv}, a*b = {a.m*b.m, b.minv*a.minv}.
// f32 throughout; no FMA contraction (the library is built with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstring>

namespace pth {

struct V3 {
  float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 vcross(V3 a, V3 b) {
  return V3{(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
inline float vlen(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 vnorm(V3 a) {
  float l = vlen(a);
  return V3{a.x / l, a.y / l, a.z / l};
}

struct M4 {
  float m[16];
  static M4 identity() {
    M4 r;
    for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    return r;
  }
  bool operator==(const M4& o) const { return memcmp(m, o.m, sizeof(m)) == 0; }
};

inline M4 mmul(const M4& a, const M4& b) {  // matrix4x4.rs Mul: row . column, left-to-right sum
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r.m[4 * i + j] = a.m[4 * i + 0] * b.m[0 + j] + a.m[4 * i + 1] * b.m[4 + j] + a.m[4 * i + 2] * b.m[8 + j] +
                       a.m[4 * i + 3] * b.m[12 + j];
  return r;
}
inline M4 mtranspose(const M4& a) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r.m[4 * i + j] = a.m[4 * j + i];
  return r;
}
// Gauss-Jordan with full pivoting in f32 (matrix4x4.rs:209-282). Returns false when singular.
inline bool minverse(const M4& a, M4* out) {
  int indxc[4] = {0, 0, 0, 0}, indxr[4] = {0, 0, 0, 0}, ipiv[4] = {0, 0, 0, 0};
  float minv[16];
  memcpy(minv, a.m, sizeof(minv));
  for (int i = 0; i < 4; i++) {
    int irow = 0, icol = 0;
    float big = 0.0f;
    for (int j = 0; j < 4; j++) {
      if (ipiv[j] != 1) {
        for (int k = 0; k < 4; k++) {
          if (ipiv[k] == 0) {
            if (fabsf(minv[4 * j + k]) >= big) {
              big = fabsf(minv[4 * j + k]);
              irow = j;
              icol = k;
            }
          } else if (ipiv[k] > 1) {
            return false;
          }
        }
      }
    }
    ipiv[icol] += 1;
    if (irow != icol)
      for (int k = 0; k < 4; k++) {
        float t = minv[4 * irow + k];
        minv[4 * irow + k] = minv[4 * icol + k];
        minv[4 * icol + k] = t;
      }
    indxr[i] = irow;
    indxc[i] = icol;
    if (minv[4 * icol + icol] == 0.0f) return false;
    float pivinv = 1.0f / minv[4 * icol + icol];
    minv[4 * icol + icol] = 1.0f;
    for (int j = 0; j < 4; j++) minv[4 * icol + j] *= pivinv;
    for (int j = 0; j < 4; j++) {
      if (j != icol) {
        float save = minv[4 * j + icol];
        minv[4 * j + icol] = 0.0f;
        for (int k = 0; k < 4; k++) minv[4 * j + k] -= minv[4 * icol + k] * save;
      }
    }
  }
  for (int j = 3; j >= 0; j--) {
    if (indxr[j] != indxc[j])
      for (int k = 0; k < 4; k++) {
        float t = minv[4 * k + indxr[j]];
        minv[4 * k + indxr[j]] = minv[4 * k + indxc[j]];
        minv[4 * k + indxc[j]] = t;
      }
  }
  memcpy(out->m, minv, sizeof(minv));
  return true;
}

struct Xf {
  M4 m, minv;
  static Xf identity() { return Xf{M4::identity(), M4::identity()}; }
  bool operator==(const Xf& o) const { return m == o.m && minv == o.minv; }
};
inline Xf operator*(const Xf& a, const Xf& b) { return Xf{mmul(a.m, b.m), mmul(b.minv, a.minv)}; }
inline Xf xinverse(const Xf& a) { return Xf{a.minv, a.m}; }

inline float radians(float d) { return d * (3.14159265358979323846f / 180.0f); }

inline M4 m_translate(float x, float y, float z) {
  M4 r = M4::identity();
  r.m[3] = x;
  r.m[7] = y;
  r.m[11] = z;
  return r;
}
inline M4 m_scale(float x, float y, float z) {
  M4 r = M4::identity();
  r.m[0] = x;
  r.m[5] = y;
  r.m[10] = z;
  return r;
}
inline Xf x_translate(float x, float y, float z) { return Xf{m_translate(x, y, z), m_translate(-x, -y, -z)}; }
inline Xf x_scale(float x, float y, float z) { return Xf{m_scale(x, y, z), m_scale(1.0f / x, 1.0f / y, 1.0f / z)}; }
inline Xf x_rotate(float theta, float x, float y, float z) {  // matrix4x4.rs:104-128
  V3 a = vnorm(v3(x, y, z));
  float s = sinf(radians(theta)), c = cosf(radians(theta));
  M4 m = M4::identity();
  m.m[0] = a.x * a.x + (1.0f - a.x * a.x) * c;
  m.m[1] = a.x * a.y * (1.0f - c) - a.z * s;
  m.m[2] = a.x * a.z * (1.0f - c) + a.y * s;
  m.m[3] = 0.0f;
  m.m[4] = a.x * a.y * (1.0f - c) + a.z * s;
  m.m[5] = a.y * a.y + (1.0f - a.y * a.y) * c;
  m.m[6] = a.y * a.z * (1.0f - c) - a.x * s;
  m.m[7] = 0.0f;
  m.m[8] = a.x * a.z * (1.0f - c) - a.y * s;
  m.m[9] = a.y * a.z * (1.0f - c) + a.x * s;
  m.m[10] = a.z * a.z + (1.0f - a.z * a.z) * c;
  m.m[11] = 0.0f;
  return Xf{m, mtranspose(m)};
}
// Transform::look_at (transform.rs:70-85): m = world-to-camera, minv = camera-to-world
inline bool x_look_at(const float e[3], const float l[3], const float u[3], Xf* out) {
  V3 pos = v3(e[0], e[1], e[2]);
  V3 look = v3(l[0], l[1], l[2]);
  V3 up = vnorm(v3(u[0], u[1], u[2]));
  V3 dir = vnorm(look - pos);
  V3 right = vcross(up, dir);
  if (vlen(right) == 0.0f) return false;
  right = vnorm(right);
  V3 new_up = vnorm(vcross(dir, right));
  M4 c2w;
  float mm[16] = {right.x, new_up.x, dir.x, pos.x, right.y, new_up.y, dir.y, pos.y,
                  right.z, new_up.z, dir.z, pos.z, 0.0f,    0.0f,     0.0f,  1.0f};
  memcpy(c2w.m, mm, sizeof(mm));
  M4 w2c;
  if (!minverse(c2w, &w2c)) return false;
  *out = Xf{w2c, c2w};
  return true;
}
inline Xf x_from_matrix(const M4& m) {
  M4 inv = M4::identity();
  minverse(m, &inv);
  return Xf{m, inv};
}
inline Xf x_perspective(float fov, float n, float f) {  // transform.rs:89-99
  M4 persp = M4::identity();
  persp.m[10] = f / (f - n);
  persp.m[11] = -f * n / (f - n);
  persp.m[14] = 1.0f;
  persp.m[15] = 0.0f;
  float inv_tan_ang = 1.0f / tanf(radians(fov) / 2.0f);
  return x_scale(inv_tan_ang, inv_tan_ang, 1.0f) * x_from_matrix(persp);
}
inline V3 m_point(const M4& M, V3 p) {
  float x = p.x, y = p.y, z = p.z;
  float xp = M.m[0] * x + M.m[1] * y + M.m[2] * z + M.m[3];
  float yp = M.m[4] * x + M.m[5] * y + M.m[6] * z + M.m[7];
  float zp = M.m[8] * x + M.m[9] * y + M.m[10] * z + M.m[11];
  float wp = M.m[12] * x + M.m[13] * y + M.m[14] * z + M.m[15];
  if (wp == 1.0f) return v3(xp, yp, zp);
  return v3(xp / wp, yp / wp, zp / wp);
}
inline V3 m_vector(const M4& M, V3 p) {
  float x = p.x, y = p.y, z = p.z;
  return v3(M.m[0] * x + M.m[1] * y + M.m[2] * z, M.m[4] * x + M.m[5] * y + M.m[6] * z,
            M.m[8] * x + M.m[9] * y + M.m[10] * z);
}
inline V3 m_normal_with_inv(const M4& Minv, V3 p) {  // Transform::transform_normal uses minv transposed
  float x = p.x, y = p.y, z = p.z;
  return v3(Minv.m[0] * x + Minv.m[4] * y + Minv.m[8] * z, Minv.m[1] * x + Minv.m[5] * y + Minv.m[9] * z,
            Minv.m[2] * x + Minv.m[6] * y + Minv.m[10] * z);
}
inline bool x_swaps_handedness(const Xf& t) {  // transform.rs:113-121
  const float* m = t.m.m;
  float det = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) +
              m[2] * (m[4] * m[9] - m[5] * m[8]);
  return det < 0.0f;
}

}  // namespace pth
