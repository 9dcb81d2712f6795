// shape_loaders.h — shapes of the reference that are triangle meshes by construction: "heightfield", "loopsubdiv", "nurbs".
// Replaces reference src/shapes/heightfield.rs:6-82, src/shapes/loopsubdiv.rs:283-760 (Loop subdivision surfaces: the
// pbrt-v3 algorithm on SDVertex / SDFace records, limit positions and limit-surface normals) and src/shapes/nurbs.rs:9-404
// (de Boor evaluation of the rational surface and its derivatives, diced into diceu x dicev vertices).  Index-based records
// instead of the reference's Arc<RefCell<..>> graph; the same faces in the same order, the same f32 operations in the same
// order (the reference keys its edge maps by pointer, which only decides which end of an edge is called p0: every sum
// over the two ends is symmetric).  Host code: runs once per shape while the scene is parsed.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pth {

struct LoadedMesh {
  std::vector<int32_t> vi;
  std::vector<float> P, N, UV;  // 3, 3, 2 floats per vertex; N / UV may be empty
};

// create_heightfield (heightfield.rs:6-82): nu x nv grid over [0,1]^2, two triangles per cell, uv = (x, y)
inline LoadedMesh heightfield_mesh(int nx, int ny, const std::vector<float>& z) {
  if (nx == -1 || ny == -1) throw std::runtime_error("Must provide \"nu\" and \"nv\" parameters to heightfield shape.");
  if (nx < 2 || ny < 2) throw std::runtime_error("heightfield: \"nu\" and \"nv\" must be at least 2");
  if (z.empty()) throw std::runtime_error("No vertex positions provided for heightfield shape.");
  if (z.size() != (size_t)nx * (size_t)ny) throw std::runtime_error("Number of \"Pz\" values doesn't match resolution.");
  LoadedMesh M;
  for (int y = 0; y < ny; y++)
    for (int x = 0; x < nx; x++) {
      const float xx = (float)x / (float)(nx - 1), yy = (float)y / (float)(ny - 1);
      M.P.push_back(xx); M.P.push_back(yy); M.P.push_back(z[(size_t)nx * y + x]);
      M.UV.push_back(xx); M.UV.push_back(yy);
    }
  auto vert = [&](int x, int y) { return (int32_t)(x + y * nx); };
  for (int y = 0; y < ny - 1; y++)
    for (int x = 0; x < nx - 1; x++) {
      const int32_t q[6] = {vert(x, y), vert(x + 1, y), vert(x + 1, y + 1), vert(x, y), vert(x + 1, y + 1), vert(x, y + 1)};
      M.vi.insert(M.vi.end(), q, q + 6);
    }
  return M;
}

namespace loop {
struct P3 { float x = 0, y = 0, z = 0; };
inline P3 operator*(float s, P3 a) { return P3{s * a.x, s * a.y, s * a.z}; }
inline P3 operator+(P3 a, P3 b) { return P3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator-(P3 a, P3 b) { return P3{a.x - b.x, a.y - b.y, a.z - b.z}; }
struct SDVertex { P3 p; int start_face = -1, child = -1; bool regular = false, boundary = false; };
struct SDFace { int v[3] = {-1, -1, -1}, f[3] = {-1, -1, -1}, children[4] = {-1, -1, -1, -1}; };
inline int NEXT(int i) { return (i + 1) % 3; }
inline int PREV(int i) { return (i + 2) % 3; }

struct Mesh {
  std::vector<SDVertex> V;
  std::vector<SDFace> F;
  int vnum(int f, int v) const {
    for (int i = 0; i < 3; i++) if (F[(size_t)f].v[i] == v) return i;
    throw std::runtime_error("loopsubdiv: inconsistent mesh (a face does not hold its vertex)");
  }
  int next_face(int f, int v) const { return F[(size_t)f].f[vnum(f, v)]; }
  int prev_face(int f, int v) const { return F[(size_t)f].f[PREV(vnum(f, v))]; }
  int next_vert(int f, int v) const { return F[(size_t)f].v[NEXT(vnum(f, v))]; }
  int prev_vert(int f, int v) const { return F[(size_t)f].v[PREV(vnum(f, v))]; }
  int other_vert(int f, int v0, int v1) const {
    for (int i = 0; i < 3; i++) if (F[(size_t)f].v[i] != v0 && F[(size_t)f].v[i] != v1) return F[(size_t)f].v[i];
    throw std::runtime_error("loopsubdiv: degenerate face");
  }
  int valence(int v) const {  // loopsubdiv.rs:56-94
    const SDVertex& s = V[(size_t)v];
    int nf = 1, f = s.start_face, guard = 0;
    if (!s.boundary) {
      for (;;) {
        int f2 = next_face(f, v);
        if (f2 == s.start_face) break;
        if (f2 < 0 || ++guard > 1000000) throw std::runtime_error("loopsubdiv: inconsistent mesh");
        f = f2; nf++;
      }
      return nf;
    }
    for (int f2; (f2 = next_face(f, v)) >= 0; f = f2) { nf++; if (++guard > 1000000) throw std::runtime_error("loopsubdiv: inconsistent mesh"); }
    f = s.start_face;
    for (int f2; (f2 = prev_face(f, v)) >= 0; f = f2) { nf++; if (++guard > 1000000) throw std::runtime_error("loopsubdiv: inconsistent mesh"); }
    return nf + 1;
  }
  std::vector<P3> one_ring(int v) const {  // loopsubdiv.rs:96-157
    const SDVertex& s = V[(size_t)v];
    std::vector<P3> pts;
    int face = s.start_face;
    if (!s.boundary) {
      for (;;) {
        pts.push_back(V[(size_t)next_vert(face, v)].p);
        int f2 = next_face(face, v);
        if (f2 == s.start_face) break;
        face = f2;
      }
    } else {
      for (int f2; (f2 = next_face(face, v)) >= 0;) face = f2;
      pts.push_back(V[(size_t)next_vert(face, v)].p);
      for (;;) {
        pts.push_back(V[(size_t)prev_vert(face, v)].p);
        int f2 = prev_face(face, v);
        if (f2 < 0) break;
        face = f2;
      }
    }
    return pts;
  }
};
inline P3 weight_one_ring(const Mesh& M, int v, float beta) {  // :283-291
  std::vector<P3> ring = M.one_ring(v);
  const float valence = (float)ring.size();
  P3 p = (1.0f - valence * beta) * M.V[(size_t)v].p;
  for (const P3& q : ring) p = p + beta * q;
  return p;
}
inline P3 weight_boundary(const Mesh& M, int v, float beta) {  // :293-299
  std::vector<P3> ring = M.one_ring(v);
  P3 p = (1.0f - 2.0f * beta) * M.V[(size_t)v].p;
  p = p + beta * ring[0];
  p = p + beta * ring[ring.size() - 1];
  return p;
}
inline float beta(int valence) { return valence == 3 ? 3.0f / 16.0f : 3.0f / (8.0f * (float)valence); }
inline float loop_gamma(int valence) { return 1.0f / ((float)valence + 3.0f / (8.0f * beta(valence))); }
}  // namespace loop

// loop_subdiv (loopsubdiv.rs:311-690)
inline LoadedMesh loop_subdiv_mesh(int n_levels, const std::vector<int32_t>& indices, const std::vector<float>& P) {
  using namespace loop;
  if (indices.empty()) throw std::runtime_error("Vertex indices \"indices\" not provided for LoopSubdiv shape.");
  if (P.empty()) throw std::runtime_error("Vertex positions \"P\" not provided for LoopSubdiv shape.");
  Mesh M;
  const size_t n_vertices = P.size() / 3, n_faces = indices.size() / 3;
  M.V.resize(n_vertices);
  for (size_t i = 0; i < n_vertices; i++) M.V[i].p = P3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
  M.F.resize(n_faces);
  for (size_t i = 0; i < n_faces; i++)
    for (int j = 0; j < 3; j++) {
      const int32_t v = indices[3 * i + j];
      if (v < 0 || (size_t)v >= n_vertices) throw std::runtime_error("loopsubdiv: vertex index out of range");
      M.F[i].v[j] = v;
      M.V[(size_t)v].start_face = (int)i;
    }
  for (size_t i = 0; i < n_vertices; i++) if (M.V[i].start_face < 0) throw std::runtime_error("loopsubdiv: a vertex belongs to no face");
  {  // face neighbours through the edges (:354-391)
    std::map<std::pair<int, int>, std::pair<int, int>> edges;  // (v0 < v1) -> (first face, its edge number)
    for (size_t i = 0; i < n_faces; i++)
      for (int en = 0; en < 3; en++) {
        int a = M.F[i].v[en], b = M.F[i].v[NEXT(en)];
        std::pair<int, int> key(a < b ? a : b, a < b ? b : a);
        auto it = edges.find(key);
        if (it == edges.end()) edges[key] = std::make_pair((int)i, en);
        else {
          M.F[(size_t)it->second.first].f[it->second.second] = (int)i;
          M.F[i].f[en] = it->second.first;
          edges.erase(it);
        }
      }
  }
  for (size_t i = 0; i < n_vertices; i++) {  // boundary / regular (:393-417)
    SDVertex& v = M.V[i];
    int face = v.start_face, guard = 0;
    for (;;) {
      int f2 = M.next_face(face, (int)i);
      if (f2 < 0) { v.boundary = true; break; }
      if (f2 == v.start_face) break;
      if (++guard > 1000000) throw std::runtime_error("loopsubdiv: inconsistent mesh");
      face = f2;
    }
    const int val = M.valence((int)i);
    v.regular = (!v.boundary && val == 6) || (v.boundary && val == 4);
  }
  for (int level = 0; level < n_levels; level++) {
    Mesh N;
    N.V.resize(M.V.size());
    for (size_t i = 0; i < M.V.size(); i++) {  // one child per vertex (:426-437)
      N.V[i].p = M.V[i].p; N.V[i].regular = M.V[i].regular; N.V[i].boundary = M.V[i].boundary;
      M.V[i].child = (int)i;
    }
    N.F.resize(4 * M.F.size());
    for (size_t i = 0; i < M.F.size(); i++)
      for (int k = 0; k < 4; k++) M.F[i].children[k] = (int)(4 * i + (size_t)k);
    for (size_t i = 0; i < M.V.size(); i++) {  // even vertices (:449-468)
      const SDVertex& v = M.V[i];
      if (!v.boundary) N.V[i].p = v.regular ? weight_one_ring(M, (int)i, 1.0f / 16.0f) : weight_one_ring(M, (int)i, beta(M.valence((int)i)));
      else N.V[i].p = weight_boundary(M, (int)i, 1.0f / 8.0f);
    }
    std::map<std::pair<int, int>, int> edge_verts;  // odd vertices (:470-520)
    for (size_t i = 0; i < M.F.size(); i++) {
      const SDFace& face = M.F[i];
      for (int k = 0; k < 3; k++) {
        int a = face.v[k], b = face.v[NEXT(k)];
        std::pair<int, int> key(a < b ? a : b, a < b ? b : a);
        if (edge_verts.count(key)) continue;
        SDVertex vert;
        vert.regular = true;
        vert.boundary = face.f[k] < 0;
        vert.start_face = face.children[3];
        const P3 p0 = M.V[(size_t)key.first].p, p1 = M.V[(size_t)key.second].p;
        if (vert.boundary) vert.p = 0.5f * p0 + 0.5f * p1;
        else {
          const P3 po0 = M.V[(size_t)M.other_vert((int)i, key.first, key.second)].p;
          const P3 po1 = M.V[(size_t)M.other_vert(face.f[k], key.first, key.second)].p;
          vert.p = (3.0f / 8.0f) * p0 + (3.0f / 8.0f) * p1 + (1.0f / 8.0f) * po0 + (1.0f / 8.0f) * po1;
        }
        edge_verts[key] = (int)N.V.size();
        N.V.push_back(vert);
      }
    }
    for (size_t i = 0; i < M.V.size(); i++) {  // start faces of the even vertices (:523-530)
      const int sf = M.V[i].start_face;
      N.V[i].start_face = M.F[(size_t)sf].children[M.vnum(sf, (int)i)];
    }
    for (size_t i = 0; i < M.F.size(); i++) {  // face neighbours (:532-567)
      const SDFace& face = M.F[i];
      for (int j = 0; j < 3; j++) {
        N.F[(size_t)face.children[3]].f[j] = face.children[NEXT(j)];
        N.F[(size_t)face.children[j]].f[NEXT(j)] = face.children[3];
        int f2 = face.f[j];
        if (f2 >= 0) N.F[(size_t)face.children[j]].f[j] = M.F[(size_t)f2].children[M.vnum(f2, face.v[j])];
        f2 = face.f[PREV(j)];
        if (f2 >= 0) N.F[(size_t)face.children[j]].f[PREV(j)] = M.F[(size_t)f2].children[M.vnum(f2, face.v[j])];
      }
    }
    for (size_t i = 0; i < M.F.size(); i++) {  // face vertices (:569-604)
      const SDFace& face = M.F[i];
      for (int j = 0; j < 3; j++) {
        N.F[(size_t)face.children[j]].v[j] = M.V[(size_t)face.v[j]].child;
        int a = face.v[j], b = face.v[NEXT(j)];
        std::pair<int, int> key(a < b ? a : b, a < b ? b : a);
        auto it = edge_verts.find(key);
        if (it != edge_verts.end()) {
          N.F[(size_t)face.children[j]].v[NEXT(j)] = it->second;
          N.F[(size_t)face.children[NEXT(j)]].v[j] = it->second;
          N.F[(size_t)face.children[3]].v[j] = it->second;
        }
      }
    }
    M = std::move(N);
  }
  // limit positions (:609-619): computed in place, vertex after vertex, like the reference (a later vertex sees the
  // already moved positions of its earlier neighbours)
  LoadedMesh out;
  for (size_t i = 0; i < M.V.size(); i++) {
    M.V[i].p = M.V[i].boundary ? weight_boundary(M, (int)i, 1.0f / 5.0f) : weight_one_ring(M, (int)i, loop_gamma(M.valence((int)i)));
    out.P.push_back(M.V[i].p.x); out.P.push_back(M.V[i].p.y); out.P.push_back(M.V[i].p.z);
  }
  const float PI = 3.14159265358979323846f;
  for (size_t i = 0; i < M.V.size(); i++) {  // limit-surface normals (:621-661)
    const SDVertex& v = M.V[i];
    P3 s, t;
    std::vector<P3> ring = M.one_ring((int)i);
    const size_t valence = ring.size();
    if (!v.boundary) {
      for (size_t j = 0; j < valence; j++) {
        s = s + cosf(2.0f * PI * (float)j / (float)valence) * ring[j];
        t = t + sinf(2.0f * PI * (float)j / (float)valence) * ring[j];
      }
    } else {
      s = ring[valence - 1] - ring[0];
      if (valence == 2) t = ring[0] + ring[1] - 2.0f * v.p;
      else if (valence == 3) t = ring[1] - v.p;
      else if (valence == 4) t = -1.0f * ring[0] + 2.0f * ring[1] + 2.0f * ring[2] + -1.0f * ring[3] + -2.0f * v.p;
      else {
        const float theta = PI / (float)(valence - 1);
        t = sinf(theta) * (ring[0] + ring[valence - 1]);
        for (size_t k = 1; k < valence - 1; k++) {
          const float wt = (2.0f * cosf(theta) - 2.0f) * sinf((float)k * theta);
          t = t + wt * ring[k];
        }
        t = -1.0f * t;
      }
    }
    P3 n{s.y * t.z - s.z * t.y, s.z * t.x - s.x * t.z, s.x * t.y - s.y * t.x};
    const float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    out.N.push_back(n.x / len); out.N.push_back(n.y / len); out.N.push_back(n.z / len);
  }
  for (const SDFace& f : M.F)
    for (int j = 0; j < 3; j++) out.vi.push_back(f.v[j]);
  return out;
}

// ---- NURBS (nurbs.rs) ------------------------------------------------------------------------------------------------
namespace nurbs {
struct H3 { float x = 0, y = 0, z = 0, w = 0; };
inline int knot_offset(const std::vector<float>& knot, int order, float t) {  // :9-19
  int k = order - 1;
  while ((size_t)k + 1 < knot.size() && t > knot[(size_t)k + 1]) k++;
  return k;
}
// nurbs_evaluate (:61-117): de Boor on the `order` control points cp[(cp_offset + i - bias) * stride]; value and derivative
// (bias = the reference's negative OffsetArray offset when the points are an iso curve that starts at control point `bias`)
inline H3 evaluate(int order, const std::vector<float>& knots, const H3* cp, int np, int stride, float t, float deriv[3], int bias = 0) {
  const int ko = knot_offset(knots, order, t);
  const float* knot = knots.data() + ko;
  const int cp_offset = ko - order + 1;
  if (cp_offset - bias < 0 || (cp_offset - bias + order - 1) * stride >= np) throw std::runtime_error("NURBS: parameter outside the knot vector");
  std::vector<H3> w((size_t)order);
  for (int i = 0; i < order; i++) w[(size_t)i] = cp[(size_t)((cp_offset + i - bias) * stride)];
  for (int i = 0; i < order - 2; i++)
    for (int j = 0; j < order - 1 - i; j++) {
      const float alpha = (knot[1 + j] - t) / (knot[1 + j] - knot[j + 2 - order + i]);
      H3& a = w[(size_t)j];
      const H3& b = w[(size_t)j + 1];
      a.x = alpha * a.x + (1.0f - alpha) * b.x;
      a.y = alpha * a.y + (1.0f - alpha) * b.y;
      a.z = alpha * a.z + (1.0f - alpha) * b.z;
      a.w = alpha * a.w + (1.0f - alpha) * b.w;
    }
  const float alpha = (knot[1] - t) / (knot[1] - knot[0]);
  H3 val;
  val.x = alpha * w[0].x + (1.0f - alpha) * w[1].x;
  val.y = alpha * w[0].y + (1.0f - alpha) * w[1].y;
  val.z = alpha * w[0].z + (1.0f - alpha) * w[1].z;
  val.w = alpha * w[0].w + (1.0f - alpha) * w[1].w;
  if (deriv) {
    const float factor = (float)(order - 1) / (knot[1] - knot[0]);
    const float dx = factor * (w[1].x - w[0].x), dy = factor * (w[1].y - w[0].y), dz = factor * (w[1].z - w[0].z), dw = factor * (w[1].w - w[0].w);
    deriv[0] = (dx / val.w) - (val.x * dw / (val.w * val.w));
    deriv[1] = (dy / val.w) - (val.y * dw / (val.w * val.w));
    deriv[2] = (dz / val.w) - (val.z * dw / (val.w * val.w));
  }
  return val;
}
}  // namespace nurbs

// create_nurbs + create_tesselated_mesh (nurbs.rs:119-404); pw = nu * nv homogeneous control points, u fastest
inline LoadedMesh nurbs_mesh(int nu, int uorder, const std::vector<float>& uknots, int nv, int vorder, const std::vector<float>& vknots,
                             const std::vector<float>& P, bool homogeneous, const float* u0p, const float* u1p, const float* v0p, const float* v1p, int diceu,
                             int dicev) {
  using nurbs::H3;
  if (nu == -1) throw std::runtime_error("Must provide number of control points \"nu\" with NURBS shape.");
  if (uorder == -1) throw std::runtime_error("Must provide u order \"uorder\" with NURBS shape.");
  if (uknots.empty()) throw std::runtime_error("Must provide u knot vector \"uknots\" with NURBS shape.");
  if (nu < 1 || uorder < 2 || (int)uknots.size() != nu + uorder) throw std::runtime_error("Number of knots in u knot vector doesn't match sum of number of u control points and u order.");
  if (nv == -1) throw std::runtime_error("Must provide number of control points \"nv\" with NURBS shape.");
  if (vorder == -1) throw std::runtime_error("Must provide v order \"vorder\" with NURBS shape.");
  if (vknots.empty()) throw std::runtime_error("Must provide v knot vector \"vknots\" with NURBS shape.");
  if (nv < 1 || vorder < 2 || (int)vknots.size() != nv + vorder) throw std::runtime_error("Number of knots in v knot vector doesn't match sum of number of v control points and v order.");
  if (P.empty()) throw std::runtime_error("Must provide control points via \"P\" or \"Pw\" parameter to NURBS shape.");
  const size_t comps = homogeneous ? 4 : 3;
  if (P.size() % comps != 0) throw std::runtime_error("Number of control points must be multiple of 3 or 4.");
  const size_t npts = P.size() / comps;
  if (npts != (size_t)nu * (size_t)nv) throw std::runtime_error("Number of control points doesn't match nu * nv.");
  std::vector<H3> pw(npts);
  for (size_t i = 0; i < npts; i++) {
    pw[i].x = P[comps * i]; pw[i].y = P[comps * i + 1]; pw[i].z = P[comps * i + 2];
    pw[i].w = homogeneous ? P[comps * i + 3] : 1.0f;
  }
  auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
  const float u0x = uknots[(size_t)uorder - 1], u1x = uknots[(size_t)nu], v0x = vknots[(size_t)vorder - 1], v1x = vknots[(size_t)nv];
  const float u0 = clampf(u0p ? *u0p : u0x, u0x, u1x), u1 = clampf(u1p ? *u1p : u1x, u0x, u1x);
  const float v0 = clampf(v0p ? *v0p : v0x, v0x, v1x), v1 = clampf(v1p ? *v1p : v1x, v0x, v1x);
  diceu = diceu < 2 ? 2 : diceu;
  dicev = dicev < 2 ? 2 : dicev;
  auto lerp = [](float t, float a, float b) { return (1.0f - t) * a + t * b; };
  LoadedMesh M;
  std::vector<H3> iso((size_t)(uorder > vorder ? uorder : vorder));
  for (int v = 0; v < dicev; v++)
    for (int u = 0; u < diceu; u++) {
      const float uu = lerp((float)u / (float)(diceu - 1), u0, u1), vv = lerp((float)v / (float)(dicev - 1), v0, v1);
      M.UV.push_back(uu); M.UV.push_back(vv);
      // nurbs_evaluate_surface (:119-159)
      const int u_first = nurbs::knot_offset(uknots, uorder, uu) - uorder + 1;
      for (int i = 0; i < uorder; i++) iso[(size_t)i] = nurbs::evaluate(vorder, vknots, pw.data() + (u_first + i), (int)npts - (u_first + i), nu, vv, nullptr);
      float dpdu[3], dpdv[3];
      const H3 p = nurbs::evaluate(uorder, uknots, iso.data(), (int)iso.size(), 1, uu, dpdu, u_first);
      const int v_first = nurbs::knot_offset(vknots, vorder, vv) - vorder + 1;
      for (int i = 0; i < vorder; i++) iso[(size_t)i] = nurbs::evaluate(uorder, uknots, pw.data() + (size_t)(v_first + i) * nu, (int)npts - (v_first + i) * nu, 1, uu, nullptr);
      nurbs::evaluate(vorder, vknots, iso.data(), (int)iso.size(), 1, vv, dpdv, v_first);
      M.P.push_back(p.x / p.w); M.P.push_back(p.y / p.w); M.P.push_back(p.z / p.w);
      float n[3] = {dpdu[1] * dpdv[2] - dpdu[2] * dpdv[1], dpdu[2] * dpdv[0] - dpdu[0] * dpdv[2], dpdu[0] * dpdv[1] - dpdu[1] * dpdv[0]};
      const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      M.N.push_back(n[0] / len); M.N.push_back(n[1] / len); M.N.push_back(n[2] / len);
    }
  auto vn = [&](int u, int v) { return (int32_t)(v * diceu + u); };
  for (int v = 0; v < dicev - 1; v++)
    for (int u = 0; u < diceu - 1; u++) {
      const int32_t q[6] = {vn(u, v), vn(u + 1, v), vn(u + 1, v + 1), vn(u, v), vn(u + 1, v + 1), vn(u, v + 1)};
      M.vi.insert(M.vi.end(), q, q + 6);
    }
  return M;
}

}  // namespace pth
