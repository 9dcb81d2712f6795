// device_api.cu — the CUDA half of the C ABI (include/pbrt_r3_b200.h): scene upload, the wavefront render loop,
// stage entry points.  Replaces SceneContext::make_scene/make_integrator (scene_context.rs:675-731), QBVHAccel::new
// (qbvh_x86.rs:352-371), Integrator::render (core/integrator/sampler.rs:259-325) and Film::write_image (film.rs:440-472).
// There is no CPU fallback: every entry point fails with an error when CUDA is unavailable.
#include "launch.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <functional>
#include <thread>
#include <vector>

#include "../host/bvh_build.h"
#include "../host/envmap.h"
#include "../host/scene_accel.h"
#include "../host/spectrum_host.h"
#include "kernels.cuh"

using namespace ptd;

void pt_set_error(const std::string& s);
// shade_tu.cu: k_shade<n_channels, cls> (PT_SHADE_BLOCK threads per block) on `stream`
void launch_shade(int n_channels, int cls, int grid, cudaStream_t stream, const DScene& S, const WaveState& W, const Queues& Q, unsigned long long* vertex_counter);

#define CK(call)                                                                                             \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) {                                                                                 \
      pt_set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return 1;                                                                                              \
    }                                                                                                        \
  } while (0)

// default path slots per wave (RGB): 6 waves of 6.4 M paths in flight = 10 GB of path state on a 180 GB part; measured
// optimum of the wave-size / waves-in-flight sweep (profiles/r1_wave_pipelining.md)
// shade kernels of one bounce on one stream per shading class (1) or back to back on the render stream (0)
#ifndef PT_SHADE_STREAMS
#define PT_SHADE_STREAMS 1
#endif
#ifndef PT_WAVE_PATHS_RGB
#define PT_WAVE_PATHS_RGB 6400000u
#endif

namespace {

struct DevBufs {
  std::vector<void*> ptrs;
  // keep = true: every upload also keeps a PINNED host staging copy so that the scene can be re-sent to the device
  // (pt_scene_reupload: the end-to-end benchmark's host->device leg)
  bool keep = false;
  struct Staged { void* dev; void* host; size_t bytes; };
  std::vector<Staged> staged;
  ~DevBufs() {
    for (void* p : ptrs) cudaFree(p);
    for (auto& st : staged) cudaFreeHost(st.host);
  }
  template <typename T>
  cudaError_t upload(const T* host, size_t n, const T** out) {
    *out = nullptr;
    if (n == 0) return cudaSuccess;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    const void* src = host;
    if (keep) {
      void* h = nullptr;
      e = cudaMallocHost(&h, n * sizeof(T));
      if (e != cudaSuccess) return e;
      memcpy(h, host, n * sizeof(T));
      staged.push_back(Staged{d, h, n * sizeof(T)});
      src = h;
    }
    e = cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T*)d;
    return e;
  }
  // copy into an existing device range (staged like upload() when keep is set)
  template <typename T>
  cudaError_t upload_into(T* dev, const T* host, size_t n) {
    if (n == 0) return cudaSuccess;
    const void* src = host;
    if (keep) {
      void* h = nullptr;
      cudaError_t e = cudaMallocHost(&h, n * sizeof(T));
      if (e != cudaSuccess) return e;
      memcpy(h, host, n * sizeof(T));
      staged.push_back(Staged{dev, h, n * sizeof(T)});
      src = h;
    }
    return cudaMemcpy(dev, src, n * sizeof(T), cudaMemcpyHostToDevice);
  }
  template <typename T>
  cudaError_t alloc(size_t n, T** out) {
    *out = nullptr;
    if (n == 0) return cudaSuccess;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    *out = (T*)d;
    return cudaSuccess;
  }
};

using pth::ObjectHost;

}  // namespace

static uint64_t next_scene_uid() {
  static uint64_t n = 0;
  return ++n;  // scenes are created by a single-threaded caller (include/pbrt_r3_b200.h)
}

struct pt_scene {
  uint64_t uid = next_scene_uid();  // never reused, unlike the address (pt_bvh_node's cache key)
  int device = 0;
  int n_sm = 148;
  DScene h;            // host copy (device pointers inside)
  DScene* d = nullptr;
  DevBufs bufs;        // scene-lifetime buffers
  std::vector<ObjectHost> objects;
  uint32_t max_stack = 0;
  cudaStream_t stream = nullptr;
  // The shade kernels of one bounce (one per shading class) run side by side on their own streams, forked from and
  // joined to `stream` with events, so that the tail of one class overlaps the body of the next.
  cudaStream_t cls_stream[CL_COUNT] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[CL_COUNT] = {};
  // film + wave state
  f4* film = nullptr;
  float* rgb_dev = nullptr;
  uint32_t film_pixels = 0;
  std::unique_ptr<DevBufs> wave_bufs;
  WaveState W;
  Queues Q;
  Queues Q_cur;                            // Q with the extend queues in the orientation of the bounce being launched
  int* slot_of = nullptr;                  // pixel -> slot maps of the waves in flight (slot_segs x sample-bounds pixels)
  std::unique_ptr<DevBufs> slot_bufs;
  uint32_t slot_segs = 0;
  uint32_t* counts_host = nullptr;         // pinned: 2 x 32 queue counters, copied after every shade stage (render_waves)
  cudaEvent_t ev_iter[2] = {}, ev_time[2][4] = {};
  unsigned long long* counters = nullptr;  // [0] nodes [1] tris [2] vertices
  int* overflow = nullptr;
  uint32_t* fetch = nullptr;               // ray fetch counter of the stage entry points
  uint32_t class_mask = 1;                 // shading classes present in the scene (bit CL_*); CL_NONE always
  uint32_t wave_cap = 0;
  // pixel list cache
  int* pixlist = nullptr;
  uint32_t pix_n = 0;
  int pix_key[3] = {-1, -1, -1};
  std::vector<int> pix_host;
  std::unique_ptr<DevBufs> pix_bufs;
  // Sampler "random": per render tile {first pixel in pixlist, pixel count, tile number in the whole film}
  std::vector<int> tiles_host;
  int* tiles = nullptr;
  ~pt_scene() {
    for (int c = 0; c < CL_COUNT; c++) {
      if (cls_stream[c]) cudaStreamDestroy(cls_stream[c]);
      if (ev_join[c]) cudaEventDestroy(ev_join[c]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (int k = 0; k < 2; k++) {
      if (ev_iter[k]) cudaEventDestroy(ev_iter[k]);
      for (int j = 0; j < 4; j++) if (ev_time[k][j]) cudaEventDestroy(ev_time[k][j]);
    }
    if (counts_host) cudaFreeHost(counts_host);
    if (stream) cudaStreamDestroy(stream);
    if (d) cudaFree(d);
  }
};

namespace {

// core/rng.rs PCG32 with the default state: compute_radical_inverse_permutations (radical_inverse.rs:93-109)
struct Pcg {
  uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
  uint32_t next() {
    uint64_t old = state;
    state = old * 0x5851f42d4c957f2dULL + inc;
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
  }
  uint32_t bounded(uint32_t b) {
    uint32_t threshold = (~b + 1u) % b;
    for (;;) {
      uint32_t r = next();
      if (r >= threshold) return r % b;
    }
  }
};

float roughness_to_alpha(float roughness) {  // trowbridge_reitz.rs:113-121 (host libm ln, like the reference)
  roughness = fmaxf(roughness, 1e-3f);
  float x = logf(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

uint32_t round_up_pow2(uint32_t v) { v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1; }
uint32_t log2int(uint32_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }
uint64_t math_mod_h(int64_t a, int64_t b) { int64_t r = a - (a / b) * b; return r < 0 ? (uint64_t)(r + b) : (uint64_t)r; }
void ext_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {
  if (b == 0) { *x = 1; *y = 0; return; }
  int64_t d = (int64_t)(a / b), xp, yp;
  ext_gcd(b, a % b, &xp, &yp);
  *x = yp;
  *y = xp - (d * yp);
}
uint64_t mult_inverse(int64_t a, int64_t n) { int64_t x, y; ext_gcd((uint64_t)a, (uint64_t)n, &x, &y); return math_mod_h(x, n); }

int grid_for(const pt_scene* s, size_t n, int block, int per_sm) {
  size_t need = (n + (size_t)block - 1) / (size_t)block;
  size_t cap = (size_t)s->n_sm * (size_t)per_sm;
  size_t g = std::min(need, cap);
  return (int)std::max<size_t>(g, 1);
}

template <int NC>
int precompute_light_distribution(pt_scene* s, uint32_t n_vox, float* func, float* cdf, float* func_int) {
  size_t total = (size_t)n_vox * s->h.n_lights;
#ifdef PT_HOSTCHECK
  {  // tests/hostcheck: the serial stand-in for a launch would spend ~10 s per Cornell scene here; slices on host threads
    hostcheck::begin_launch();
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), 16));
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) {
      const size_t a = total * t / nt, b = total * (t + 1) / nt;
      th.emplace_back([=]() { k_voxel_contrib<NC>(s->h, func, n_vox, a, b - a); });
    }
    for (std::thread& t : th) t.join();
  }
#else
  PT_LAUNCH(grid_for(s, total, 128, 16), 128, s->stream, k_voxel_contrib<NC>)(s->h, func, n_vox, (size_t)0, total);
  CK(cudaGetLastError());
#endif
  PT_LAUNCH(grid_for(s, n_vox, 128, 16), 128, s->stream, k_voxel_cdf)(s->h.n_lights, n_vox, func, cdf, func_int);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

// copies 0 and 1 of the triangle records from the identity copy (dscene.cuh: recs_perm)
int permute_recs(pt_scene* s) {
#if PT_TRI_PERM
  const uint32_t n = s->h.rec_stride;
  if (n == 0) return 0;
  PT_LAUNCH(grid_for(s, n, 256, 64), 256, s->stream, k_permute_recs)(const_cast<f4*>(s->h.recs_perm), n);
  CK(cudaGetLastError());
#else
  (void)s;
#endif
  return 0;
}

int build_scene(pt_scene* s, const pt_scene_desc* D) {
  DScene& H = s->h;
  memset(&H, 0, sizeof(H));
  s->bufs.keep = true;
  const uint32_t NC = D->n_channels;
  H.n_channels = NC;
  H.n_tris = D->n_tris;
  H.n_lights = D->n_lights;
  H.n_materials = D->n_materials;
  H.n_objects = D->n_objects;
  H.n_instances = D->n_instances;
  H.has_instances = D->n_instances > 0 ? 1 : 0;
  memcpy(H.xyz_w, D->xyz_weights, sizeof(H.xyz_w));
  H.xyz_scale = D->xyz_scale;
  H.camera = D->camera;
  H.film = D->film;
  H.integrator = D->integrator;

  // ---- accelerators: instanced objects first, then the top level ------------------------------------------
  {
    std::string err = pth::build_objects(D, &s->objects);
    if (!err.empty()) { pt_set_error(err); return 1; }
  }
  // stack depth: the worst-case bound 3*depth+1 is far from typical use; the kernels flag a real overflow instead
  std::vector<pth::GpuNodeH> nodes;
  std::vector<pth::LNodeH> lnodes;
  const bool lbvh = D->accel.type == PT_ACCEL_LBVH;
  H.accel_type = lbvh ? PT_ACCEL_LBVH : PT_ACCEL_QBVH;
  std::vector<f4> recs;
  std::vector<ObjectRec> orecs(D->n_objects);
  for (uint32_t k = 0; k < D->n_objects; k++) {
    ObjectHost& O = s->objects[k];
    ObjectRec& R = orecs[k];
    memset(&R, 0, sizeof(R));
    R.node_base = (uint32_t)(lbvh ? lnodes.size() : nodes.size());
    R.rec_base = (uint32_t)(recs.size() / 3);
    if (lbvh) {
      const size_t before = lnodes.size();
      uint32_t depth = pth::flatten_lbvh(O.bvh, &lnodes);
      if (depth + 2 > (uint32_t)PT_MAX_STACK / 2) { pt_set_error("BVH too deep for the LBVH traversal stack"); return 1; }
      R.n_nodes = (uint32_t)(lnodes.size() - before);
    } else
    R.n_nodes = (uint32_t)O.bvh.nodes.size();
    R.n_prims = (uint32_t)O.ordered.size();
    R.single = O.single ? 1 : 0;
    for (int a = 0; a < 3; a++) { R.bmin[a] = O.bounds.mn[a]; R.bmax[a] = O.bounds.mx[a]; }
    if (!lbvh) nodes.insert(nodes.end(), O.bvh.nodes.begin(), O.bvh.nodes.end());
    for (int32_t p : O.ordered) {
      if (p >= 0) {
        const float* q = D->tri_p + (size_t)p * 9;
        recs.push_back(mk4(q[0], q[1], q[2], bitsf((uint32_t)p)));
        recs.push_back(mk4(q[3], q[4], q[5], bitsf(D->tri_flags[p] & 0xffu)));
        recs.push_back(mk4(q[6], q[7], q[8], 0.0f));
      } else {
        // instance record: v0.xyz / v2.xyz = a CONSERVATIVE world-space box of the instance for the cull test of
        // trav_leaf (dtrav2.cuh): TransformedPrimitive::world_bound grown by 1e-3 * (1 + |coordinate|), i.e. four
        // orders of magnitude above the rounding of the ray transform it stands in for.  Objects with a single
        // primitive have no bounds test in the reference (the triangle is always tested and counted): infinite box.
        const pt_instance& I = D->instances[(uint32_t)~p];
        const ObjectHost& IO = s->objects[(size_t)I.object];
        float lo[3], hi[3];
        if (IO.single || IO.ordered.empty()) {
          for (int a = 0; a < 3; a++) { lo[a] = -INFINITY; hi[a] = INFINITY; }
        } else {
          pth::BBox wb = pth::instance_world_bounds(I.m, IO.bounds);
          for (int a = 0; a < 3; a++) {
            float g = 1e-3f * (1.0f + fmaxf(fabsf(wb.mn[a]), fabsf(wb.mx[a])));
            lo[a] = wb.mn[a] - g; hi[a] = wb.mx[a] + g;
            if (!(lo[a] <= hi[a])) { lo[a] = -INFINITY; hi[a] = INFINITY; }  // NaN / inverted bounds: never cull
          }
        }
        recs.push_back(mk4(lo[0], lo[1], lo[2], bitsf((uint32_t)~p)));
        recs.push_back(mk4(0, 0, 0, bitsf(PT_REC_INSTANCE)));
        recs.push_back(mk4(hi[0], hi[1], hi[2], 0));
      }
    }
  }
  CK(s->bufs.upload((const QNode*)nodes.data(), nodes.size(), &H.nodes));
  static_assert(sizeof(pth::LNodeH) == 2 * sizeof(f4), "LNodeH = 2 x f4");
  CK(s->bufs.upload((const f4*)lnodes.data(), lnodes.size() * 2, &H.lnodes));
#if PT_TRI_PERM
  {  // three copies on the device, one on the wire: the identity copy is uploaded, k_permute_recs derives the other two
    const size_t n = recs.size();
    if (n > 0xffffffffull / 3 || (D->n_instances > 0 && n / 3 >= (1ull << 27))) { pt_set_error("too many triangle records"); return 1; }
    f4* all = nullptr;
    CK(s->bufs.alloc(3 * n, &all));
    if (n > 0) CK(s->bufs.upload_into(all + 2 * n, recs.data(), n));
    H.recs_perm = all;
    H.recs = all ? all + 2 * n : nullptr;
    H.rec_stride = (uint32_t)n;
    if (permute_recs(s)) return 1;
  }
#else
  CK(s->bufs.upload(recs.data(), recs.size(), &H.recs));
  H.recs_perm = nullptr;
  H.rec_stride = 0;
#endif
  CK(s->bufs.upload(orecs.data(), orecs.size(), &H.objects));
  {
    std::vector<InstanceRec> ir(D->n_instances);
    for (uint32_t i = 0; i < D->n_instances; i++) {
      memset(&ir[i], 0, sizeof(InstanceRec));
      memcpy(ir[i].m, D->instances[i].m, sizeof(ir[i].m));
      memcpy(ir[i].minv, D->instances[i].minv, sizeof(ir[i].minv));
      ir[i].object = D->instances[i].object;
    }
    CK(s->bufs.upload(ir.data(), ir.size(), &H.instances));
  }
  {
    std::vector<f4> tv((size_t)D->n_tris * 3);
    for (size_t t = 0; t < D->n_tris; t++)
      for (int v = 0; v < 3; v++) tv[t * 3 + v] = mk4(D->tri_p[t * 9 + v * 3], D->tri_p[t * 9 + v * 3 + 1], D->tri_p[t * 9 + v * 3 + 2], 0.0f);
    CK(s->bufs.upload(tv.data(), tv.size(), &H.tri_v));
  }
  if (D->tri_n) CK(s->bufs.upload(D->tri_n, (size_t)D->n_tris * 9, &H.tri_n));
  if (D->tri_uv) CK(s->bufs.upload(D->tri_uv, (size_t)D->n_tris * 6, &H.tri_uv));
  if (D->tri_s) CK(s->bufs.upload(D->tri_s, (size_t)D->n_tris * 9, &H.tri_s));
  CK(s->bufs.upload(D->tri_flags, D->n_tris, &H.tri_flags));
  CK(s->bufs.upload(D->tri_material, D->n_tris, &H.tri_material));
  CK(s->bufs.upload(D->tri_light, D->n_tris, &H.tri_light));

  // ---- materials: clamp_zero'd spectra become extra pool entries; roughness -> alpha on the host ---------------
  std::vector<float> spectra(D->spectra, D->spectra + (size_t)D->n_spectra * NC);
  std::vector<pt_material> mats(D->materials, D->materials + D->n_materials);
  std::vector<float> alpha((size_t)D->n_materials * 2, 0.0f);
  auto clamped = [&](int idx) {
    if (idx < 0) return idx;
    bool neg = false;
    for (uint32_t c = 0; c < NC; c++) if (spectra[(size_t)idx * NC + c] < 0.0f) neg = true;
    if (!neg) return idx;
    size_t n = spectra.size() / NC;
    for (uint32_t c = 0; c < NC; c++) { float v = spectra[(size_t)idx * NC + c]; spectra.push_back(v < 0.0f ? 0.0f : v); }
    return (int)n;
  };
  for (uint32_t m = 0; m < D->n_materials; m++) {
    pt_material& M = mats[m];
    if (M.type == PT_MAT_PLASTIC) { M.kd = clamped(M.kd); M.ks = clamped(M.ks); }
    if (M.type == PT_MAT_MIRROR) M.kr = clamped(M.kr);
    float ur = M.uroughness, vr = M.vroughness;
    if (M.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
    alpha[2 * m] = ur;
    alpha[2 * m + 1] = vr;
  }
  // uber / substrate / translucent (materials/{uber,substrate,translucent}.rs compute_scattering_functions): every parameter
  // is a constant, so the lobes are listed here once; products of spectra become pool entries (RGB only, front end)
  std::vector<MatLobes> lobes(D->n_materials);
  {
    auto get = [&](int idx, float* c) { for (uint32_t k = 0; k < NC; k++) c[k] = idx >= 0 ? spectra[(size_t)idx * NC + k] : 0.0f; };
    auto clamp0 = [&](float* c) { for (uint32_t k = 0; k < NC; k++) c[k] = c[k] < 0.0f ? 0.0f : c[k]; };
    auto black = [&](const float* c) { for (uint32_t k = 0; k < NC; k++) if (c[k] != 0.0f) return false; return true; };
    auto add = [&](const float* c) { size_t n = spectra.size() / NC; for (uint32_t k = 0; k < NC; k++) spectra.push_back(c[k]); return (int)n; };
    // the lobes of material m appended to T (every lobe scaled by spectrum `scale`, -1 = none); false = no BSDF at all
    const bool multi = D->integrator.type == PT_INTEGRATOR_PATH;  // allow_multiple_lobes: path.rs passes true, the tree integrators false
    std::function<bool(uint32_t, MatLobes&, int)> list_lobes = [&](uint32_t m, MatLobes& T, int scale) -> bool {
      const pt_material& M = mats[m];
      const float ax = std::max(0.001f, alpha[2 * m]), ay = std::max(0.001f, alpha[2 * m + 1]);  // TrowbridgeReitzDistribution::new
      bool overflow = false;
      auto lobe = [&](int kind, int sp0, int sp1, int fres, float a, float b, float ei, float et) {
        if (T.n >= 5) { overflow = true; return; }
        const int k = T.n++;
        T.kind[k] = kind; T.sp0[k] = sp0; T.sp1[k] = sp1; T.fres[k] = fres; T.a[k] = a; T.b[k] = b; T.ei[k] = ei; T.et[k] = et; T.scale[k] = scale;
      };
      float kd[60], ks[60], kr[60], kt[60], op[60], t[60], x[60];
      get(M.kd, kd); get(M.ks, ks); get(M.kr, kr); get(M.kt, kt); get(M.opacity, op);
      if (M.type == PT_MAT_MATTE) {  // matte.rs:37-56 (Kd is not clamped)
        if (!black(kd)) {
          const float sig = std::min(std::max(M.sigma, 0.0f), 90.0f);
          if (sig == 0.0f) lobe(LB_LAMBERT, M.kd, -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
          else {
            const float sg = sig * (PT_PI / 180.0f), s2 = sg * sg;
            lobe(LB_OREN, M.kd, -1, FR_NOOP, 1.0f - (s2 / (2.0f * (s2 + 0.33f))), 0.45f * s2 / (s2 + 0.09f), 1.0f, 1.0f);
          }
        }
        return !overflow;
      }
      if (M.type == PT_MAT_PLASTIC) {  // plastic.rs:43-70 (Kd, Ks clamped above)
        if (!black(kd)) lobe(LB_LAMBERT, M.kd, -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
        if (!black(ks)) lobe(LB_MF_REFL, M.ks, -1, FR_DIEL, ax, ax, 1.5f, 1.0f);
        return !overflow;
      }
      if (M.type == PT_MAT_MIRROR) {
        if (!black(kr)) lobe(LB_MIRROR, M.kr, -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
        return !overflow;
      }
      if (M.type == PT_MAT_METAL) {
        lobe(LB_MF_REFL, M.kr, M.ks, FR_COND, ax, ay, 1.0f, 1.0f);
        return !overflow;
      }
      if (M.type == PT_MAT_GLASS) {  // glass.rs:47-110 (Kr = kr, Kt = ks)
        const bool rb = black(kr), tb = black(ks);
        if (rb && tb) return false;
        T.eta = M.eta;
        if (M.uroughness == 0.0f && M.vroughness == 0.0f) {
          if (multi) lobe(LB_FRESNEL_SPEC, M.kr, M.ks, FR_NOOP, 0.0f, 0.0f, 1.0f, M.eta);
          else {
            if (!rb) lobe(LB_SPEC_REFL_D, M.kr, -1, FR_DIEL, 0.0f, 0.0f, 1.0f, M.eta);
            if (!tb) lobe(LB_SPEC_TRANS, M.ks, -1, FR_NOOP, 0.0f, 0.0f, 1.0f, M.eta);
          }
        } else {
          if (!rb) lobe(LB_MF_REFL, M.kr, -1, FR_DIEL, ax, ay, 1.0f, M.eta);
          if (!tb) lobe(LB_MF_TRANS, M.ks, -1, FR_NOOP, ax, ay, 1.0f, M.eta);
        }
        return !overflow;
      }
      clamp0(kd); clamp0(ks);
      if (M.type == PT_MAT_UBER) {  // uber.rs:85-125
        clamp0(kr); clamp0(kt);
        for (uint32_t k = 0; k < NC; k++) t[k] = 1.0f - op[k];
        clamp0(t);
        T.eta = !black(t) ? 1.0f : M.eta;
        if (!black(t)) lobe(LB_SPEC_TRANS, add(t), -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
        for (uint32_t k = 0; k < NC; k++) x[k] = op[k] * kd[k];
        if (!black(x)) lobe(LB_LAMBERT, add(x), -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
        for (uint32_t k = 0; k < NC; k++) x[k] = op[k] * ks[k];
        if (!black(x)) lobe(LB_MF_REFL, add(x), -1, FR_DIEL, ax, ay, 1.0f, M.eta);
        for (uint32_t k = 0; k < NC; k++) x[k] = op[k] * kr[k];
        if (!black(x)) lobe(LB_SPEC_REFL_D, add(x), -1, FR_DIEL, 0.0f, 0.0f, 1.0f, M.eta);
        for (uint32_t k = 0; k < NC; k++) x[k] = op[k] * kt[k];
        if (!black(x)) lobe(LB_SPEC_TRANS, add(x), -1, FR_NOOP, 0.0f, 0.0f, 1.0f, M.eta);
      } else if (M.type == PT_MAT_SUBSTRATE) {  // substrate.rs:44-59
        if (!black(kd) && !black(ks)) lobe(LB_FRESNEL_BLEND, add(kd), add(ks), FR_NOOP, ax, ay, 1.0f, 1.0f);
      } else if (M.type == PT_MAT_TRANSLUCENT) {  // translucent.rs:48-98 (reflect = kr, transmit = kt)
        clamp0(kr); clamp0(kt);
        T.eta = 1.5f;
        if (black(kr) && black(kt)) return false;
        if (!black(kd)) {
          for (uint32_t k = 0; k < NC; k++) x[k] = kr[k] * kd[k];
          if (!black(kr)) lobe(LB_LAMBERT, add(x), -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
          for (uint32_t k = 0; k < NC; k++) x[k] = kt[k] * kd[k];
          if (!black(kt)) lobe(LB_LAMBERT_T, add(x), -1, FR_NOOP, 0.0f, 0.0f, 1.0f, 1.0f);
        }
        if (!black(ks)) {
          for (uint32_t k = 0; k < NC; k++) x[k] = kr[k] * ks[k];
          if (!black(kr)) lobe(LB_MF_REFL, add(x), -1, FR_DIEL, ax, ay, 1.0f, 1.5f);
          for (uint32_t k = 0; k < NC; k++) x[k] = kt[k] * ks[k];
          if (!black(kt)) lobe(LB_MF_TRANS, add(x), -1, FR_NOOP, ax, ay, 1.0f, 1.5f);
        }
      }
      return !overflow;
    };
    for (uint32_t m = 0; m < D->n_materials; m++) {
      const pt_material& M = mats[m];
      MatLobes& T = lobes[m];
      memset(&T, 0, sizeof(T));
      T.eta = 1.0f;
      for (int k = 0; k < 5; k++) T.scale[k] = -1;
      if (M.type < PT_MAT_UBER || M.type > PT_MAT_MIX) continue;
      if (NC != 3) { pt_set_error("uber / substrate / translucent / mix materials are only supported with RGB spectra"); return 1; }
      if (M.type != PT_MAT_MIX) {
        if (!list_lobes(m, T, -1)) T.n = -1;
        continue;
      }
      // MixMaterial (mix.rs:60-93): the lobes of both materials as ScaledBxDFs, s1 = clamp(amount), s2 = clamp(1 - s1); eta of the first
      if (M.kd < 0 || M.ks < 0 || (uint32_t)M.kd >= D->n_materials || (uint32_t)M.ks >= D->n_materials || mats[M.kd].type == PT_MAT_MIX || mats[M.ks].type == PT_MAT_MIX) {
        pt_set_error("mix material: sub-material index out of range, or a nested mix (not supported)"); return 1;
      }
      float s1[60], s2[60];
      get(M.kr, s1); clamp0(s1);
      for (uint32_t k = 0; k < NC; k++) s2[k] = 1.0f - s1[k];
      clamp0(s2);
      MatLobes A, B2;
      memset(&A, 0, sizeof(A)); memset(&B2, 0, sizeof(B2));
      A.eta = B2.eta = 1.0f;
      const bool ok1 = list_lobes((uint32_t)M.kd, A, add(s1)), ok2 = list_lobes((uint32_t)M.ks, B2, add(s2));
      if (!ok1 || !ok2 || A.n + B2.n > 5) { pt_set_error("mix material: a sub-material without a BSDF, or more than five lobes in total (not supported on the GPU path)"); return 1; }
      T = A;
      for (int k = 0; k < B2.n; k++) {
        const int q = T.n++;
        T.kind[q] = B2.kind[k]; T.sp0[q] = B2.sp0[k]; T.sp1[q] = B2.sp1[k]; T.fres[q] = B2.fres[k]; T.a[q] = B2.a[k]; T.b[q] = B2.b[k]; T.ei[q] = B2.ei[k]; T.et[q] = B2.et[k];
        T.scale[q] = B2.scale[k];
      }
    }
  }
  s->class_mask = 1u;
  for (uint32_t m = 0; m < D->n_materials; m++) {
    if (mats[m].type >= 0 && mats[m].type <= 4) s->class_mask |= 1u << (mats[m].type + 1);
    else if (mats[m].type >= PT_MAT_UBER && mats[m].type <= PT_MAT_MIX) s->class_mask |= 1u << CL_EXTRA;
  }
  CK(s->bufs.upload(spectra.data(), spectra.size(), &H.spectra));
  CK(s->bufs.upload(lobes.data(), lobes.size(), &H.mat_lobes));
  CK(s->bufs.upload(mats.data(), mats.size(), &H.materials));
  CK(s->bufs.upload(alpha.data(), alpha.size(), &H.alpha));
  CK(s->bufs.upload(D->lights, (size_t)D->n_lights + D->n_emitters_only, &H.lights));
  CK(s->bufs.upload(D->spots, D->n_spots, &H.spots));
  // ---- infinite lights (lights/infinite.rs): level 0 of the map + Distribution2D tables; the pyramid stays on the host
  H.envs = nullptr; H.inf_lights = nullptr; H.n_inf = 0; H.has_map_lights = 0;
  std::vector<pth::EnvMapHost> env_host(D->n_envmaps);
  {
    std::vector<EnvDev> envs(D->n_envmaps);
    for (uint32_t k = 0; k < D->n_envmaps; k++) {
      const pt_envmap& M = D->envmaps[k];
      if (M.width < 1 || M.height < 1 || M.width > (1 << 14) || M.height > (1 << 14) || !M.texels) { pt_set_error("pt_envmap: width / height must be in 1 .. 16384"); return 1; }
      if (NC != 3) { pt_set_error("infinite lights are only supported with RGB spectra"); return 1; }
      pth::EnvMapHost& E = env_host[k];
      E.build_pyramid(M.width, M.height, M.texels);
      E.build_distribution();
      EnvDev& V = envs[k];
      memset(&V, 0, sizeof(V));
      memcpy(V.l2w, M.l2w, sizeof(V.l2w));
      memcpy(V.w2l, M.w2l, sizeof(V.w2l));
      const pth::EnvMapHost::Level& L0 = E.pyramid[0];  // the map as the MIPMap holds it (resampled to powers of two)
      V.w = L0.w; V.h = L0.h; V.nu = E.nu; V.nv = E.nv; V.m_int = E.m_int;
      V.hitcher = M.hitcher;
      memcpy(V.xf, M.xf, sizeof(V.xf));
      memcpy(V.screen, M.screen_bounds, sizeof(V.screen));
      std::vector<f4> tex((size_t)L0.w * L0.h);
      for (size_t i = 0; i < tex.size(); i++) tex[i] = mk4(L0.rgb[3 * i], L0.rgb[3 * i + 1], L0.rgb[3 * i + 2], 0.0f);
      CK(s->bufs.upload(tex.data(), tex.size(), &V.texels));
      CK(s->bufs.upload(E.func.data(), E.func.size(), &V.func));
      CK(s->bufs.upload(E.cdf.data(), E.cdf.size(), &V.cdf));
      CK(s->bufs.upload(E.row_int.data(), E.row_int.size(), &V.row_int));
      CK(s->bufs.upload(E.mcdf.data(), E.mcdf.size(), &V.mcdf));
    }
    std::vector<int32_t> inf;
    for (uint32_t i = 0; i < D->n_lights; i++) {
      const int lt = D->lights[i].type;
      if (lt != PT_LIGHT_INFINITE && lt != PT_LIGHT_PROJECTION && lt != PT_LIGHT_GONIOMETRIC) continue;
      if (D->lights[i].tri < 0 || (uint32_t)D->lights[i].tri >= D->n_envmaps) { pt_set_error("infinite / projection / goniometric light: envmap index out of range"); return 1; }
      if (lt == PT_LIGHT_INFINITE) inf.push_back((int32_t)i);
      else H.has_map_lights = 1;
    }
    if (!envs.empty()) CK(s->bufs.upload(envs.data(), envs.size(), &H.envs));
    if (!inf.empty()) CK(s->bufs.upload(inf.data(), inf.size(), &H.inf_lights));
    H.n_inf = (uint32_t)inf.size();
  }
  H.light_sample_prefix = nullptr;
  if (D->integrator.type == PT_INTEGRATOR_DIRECT || D->integrator.type == PT_INTEGRATOR_WHITTED) {
    if (D->integrator.max_depth > PT_RI_MAX_DEPTH) { pt_set_error("directlighting / whitted: maxdepth above 32 is not supported on the GPU path"); return 1; }
    std::vector<uint32_t> prefix(D->n_lights + 1, 0u);  // directlighting "all": lighting round -> (light, sample)
    for (uint32_t i = 0; i < D->n_lights; i++) prefix[i + 1] = prefix[i] + (uint32_t)std::max(D->lights[i].n_samples, 1);
    if (prefix[D->n_lights] > 0x7fffu) { pt_set_error("directlighting: more than 32767 light samples per vertex"); return 1; }
    CK(s->bufs.upload(prefix.data(), prefix.size(), &H.light_sample_prefix));
  }
  H.tile_arrays = nullptr;
  H.tile_arr_stride = 0u;
  H.tile_tab = 0u;
  H.tile_sample = nullptr;
  if (D->sampler.type >= PT_SAMPLER_RANDOM) {  // floats per tile of the random / stratified samplers' tables and 2D sample arrays (build_pixlist allocates them)
    uint64_t elems = 0;
    if (D->integrator.type == PT_INTEGRATOR_AO) elems = (uint64_t)std::max(D->integrator.ao_samples, 0);
    if (D->integrator.type == PT_INTEGRATOR_DIRECT && D->integrator.direct_strategy == PT_DIRECT_ALL) {
      for (uint32_t i = 0; i < D->n_lights; i++) elems += 2ull * (uint64_t)std::max(D->lights[i].n_samples, 1);
      elems *= (uint64_t)std::max(D->integrator.max_depth, 0);
    }
    const uint64_t spp_eff = D->sampler.type == PT_SAMPLER_ZEROTWO ? round_up_pow2((uint32_t)std::max(D->sampler.samples_per_pixel, 1)) : (uint64_t)std::max(D->sampler.samples_per_pixel, 0);
    elems *= 2ull * spp_eff;
    if (D->sampler.type >= PT_SAMPLER_STRATIFIED) {
      H.tile_tab = 3u * (uint32_t)std::max(D->sampler.n_sampled_dimensions, 0) * (uint32_t)spp_eff;
      elems += H.tile_tab;
    }
    if (elems > (1ull << 24)) { pt_set_error("Sampler \"random\": the integrator's sample arrays exceed 16M values per pixel"); return 1; }
    H.tile_arr_stride = (uint32_t)elems;
  }
  for (uint32_t i = 0; i < D->n_lights; i++)
    if (D->lights[i].type == PT_LIGHT_SPOT && (D->lights[i].tri < 0 || (uint32_t)D->lights[i].tri >= D->n_spots)) { pt_set_error("spot light index out of range"); return 1; }

  // ---- sampler (halton.rs:57-101, sobol.rs:18-31) ---------------------------------------------------------------
  SamplerDev& SD = H.sampler;
  memset(&SD, 0, sizeof(SD));
  SD.type = D->sampler.type;
  SD.spp = D->sampler.samples_per_pixel;
  SD.at_center = D->sampler.sample_at_pixel_center;
  SD.nd = D->sampler.n_sampled_dimensions; SD.strata = D->sampler.strata; SD.jitter = D->sampler.jitter;
  SD.inc2d = 2u; SD.x2d = 0u; SD.y2d = 1u;
  if (SD.type == PT_SAMPLER_ZEROTWO) {  // ZeroTwoSequenceSampler::new rounds up to a power of two (zerotwosequence.rs:13-21)
    SD.spp = (int)round_up_pow2((uint32_t)std::max(SD.spp, 1));
    if (SD.nd < 0 || SD.nd > 64) { pt_set_error("pt_sampler: 0 <= n_sampled_dimensions <= 64"); return 1; }
    SD.inc2d = 0x10000u; SD.x2d = 0x4000u; SD.y2d = 0x8000u;
  }
  if (SD.type == PT_SAMPLER_STRATIFIED) {
    if (SD.strata < 1 || SD.strata * SD.strata != SD.spp || SD.nd < 0 || SD.nd > 64) { pt_set_error("pt_sampler: stratified needs strata^2 = samples_per_pixel, 0 <= n_sampled_dimensions <= 64"); return 1; }
    SD.inc2d = 0x10000u; SD.x2d = 0x4000u; SD.y2d = 0x8000u;
  }
  for (int i = 0; i < 4; i++) SD.sb[i] = D->film.sample_bounds[i];
  if (SD.sb[2] - SD.sb[0] >= 65536 || SD.sb[3] - SD.sb[1] >= 32768) { pt_set_error("film too large for packed pixel ids"); return 1; }
  {
    int res[2] = {SD.sb[2] - SD.sb[0], SD.sb[3] - SD.sb[1]};
    const int BASES[2] = {2, 3};
    for (int i = 0; i < 2; i++) {
      int scale = 1, e = 0;
      while (scale < std::min(res[i], 128)) { scale *= BASES[i]; e++; }
      SD.base_scales[i] = scale;
      SD.base_exponents[i] = e;
    }
    SD.sample_stride = SD.base_scales[0] * SD.base_scales[1];
    SD.mult_inverse[0] = (int)mult_inverse(SD.base_scales[1], SD.base_scales[0]);
    SD.mult_inverse[1] = (int)mult_inverse(SD.base_scales[0], SD.base_scales[1]);
    if (SD.type == PT_SAMPLER_SOBOL) {
      SD.spp = (int)round_up_pow2((uint32_t)SD.spp);
      SD.resolution = round_up_pow2((uint32_t)std::max(res[0], res[1]));
      SD.log2_resolution = log2int(SD.resolution);
    }
  }
  {
    std::vector<uint32_t> primes, sums;
    for (uint32_t n = 2; primes.size() < 1024; n++) {
      bool is = true;
      for (uint32_t p : primes) { if ((uint64_t)p * p > n) break; if (n % p == 0) { is = false; break; } }
      if (is) primes.push_back(n);
    }
    uint32_t acc = 0;
    for (uint32_t p : primes) { sums.push_back(acc); acc += p; }
    std::vector<uint16_t> perms(acc);
    Pcg rng;
    size_t off = 0;
    for (uint32_t p : primes) {
      for (uint32_t j = 0; j < p; j++) perms[off + j] = (uint16_t)j;
      for (uint32_t j = 0; j < p; j++) {  // shuffle_array (sampling.rs:4-15)
        uint32_t other = j + rng.bounded(p - j);
        std::swap(perms[off + j], perms[off + other]);
      }
      off += p;
    }
    std::vector<PrimeRec> recs1024(primes.size());
    for (size_t i = 0; i < primes.size(); i++) {
      recs1024[i].prime = primes[i];
      recs1024[i].perm_offset = sums[i];
      recs1024[i].magic = (~0ull / primes[i]) + 1ull;  // ceil(2^64 / p)
    }
    CK(s->bufs.upload(recs1024.data(), recs1024.size(), &H.prime_recs));
    CK(s->bufs.upload(perms.data(), perms.size(), &H.perms));
  }
  if (SD.type == PT_SAMPLER_SOBOL) {
    try {
      std::string dir = pth::data_dir();
      std::vector<uint32_t> m32(1024 * 52);
      std::vector<uint64_t> vdc((25 + 26) * 52);
      FILE* f = fopen((dir + "/sobol32.bin").c_str(), "rb");
      if (!f || fread(m32.data(), 4, m32.size(), f) != m32.size()) throw std::runtime_error("cannot read " + dir + "/sobol32.bin");
      fclose(f);
      f = fopen((dir + "/vdc.bin").c_str(), "rb");
      if (!f || fread(vdc.data(), 8, vdc.size(), f) != vdc.size()) throw std::runtime_error("cannot read " + dir + "/vdc.bin");
      fclose(f);
      CK(s->bufs.upload(m32.data(), m32.size(), &H.sobol32));
      CK(s->bufs.upload(vdc.data(), vdc.size(), &H.vdc));
    } catch (const std::exception& e) { pt_set_error(e.what()); return 1; }
  }

  // ---- world bound, light distribution -------------------------------------------------------------------------------
  const pth::BBox& wb = s->objects[0].bounds;
  {
    float dx = wb.mx[0] - wb.mn[0], dy = wb.mx[1] - wb.mn[1], dz = wb.mx[2] - wb.mn[2];
    H.world_radius = sqrtf(dx * dx + dy * dy + dz * dz) * 0.5f;  // Bounds3f::bounding_sphere (bounds3.rs:184-188)
  }
  LightDistDev& LD = H.ld;
  memset(&LD, 0, sizeof(LD));
  int strat = D->integrator.light_strategy;
  if (strat == PT_LIGHTS_UNIFORM && D->n_lights != 1) strat = PT_LIGHTS_SPATIAL;
  if (D->integrator.type != PT_INTEGRATOR_PATH) strat = PT_LIGHTS_UNIFORM;  // only the path integrator reads a light distribution: no table to build
  LD.strategy = strat;
  for (int a = 0; a < 3; a++) { LD.wb_min[a] = wb.mn[a]; LD.wb_max[a] = wb.mx[a]; LD.voxels[a] = 1; }
  uint32_t n_vox = 1;
  if (strat == PT_LIGHTS_SPATIAL && D->n_lights > 0) {
    float diag[3] = {wb.mx[0] - wb.mn[0], wb.mx[1] - wb.mn[1], wb.mx[2] - wb.mn[2]};
    int me = (diag[0] > diag[1] && diag[0] > diag[2]) ? 0 : (diag[1] > diag[2] ? 1 : 2);
    for (int a = 0; a < 3; a++) {
      float v = ceilf(diag[a] / diag[me] * 64.0f);
      uint32_t u = (v != v || v <= 0.0f) ? 0u : (v >= 4294967296.0f ? 0xffffffffu : (uint32_t)v);
      LD.voxels[a] = std::min(std::max(u, 1u), 64u);
    }
    n_vox = LD.voxels[0] * LD.voxels[1] * LD.voxels[2];
  }
  float *func = nullptr, *cdf = nullptr, *fint = nullptr;
  if (D->n_lights > 0) {
    size_t tot = (size_t)n_vox * D->n_lights;
    if (tot > (size_t)3 << 30) { pt_set_error("light distribution table too large"); return 1; }
    CK(s->bufs.alloc(tot, &func));
    CK(s->bufs.alloc((size_t)n_vox * (D->n_lights + 1), &cdf));
    CK(s->bufs.alloc(n_vox, &fint));
    LD.func = func; LD.cdf = cdf; LD.func_int = fint;
  }
  CK(cudaMalloc(&s->d, sizeof(DScene)));
  CK(cudaMemcpy(s->d, &H, sizeof(DScene), cudaMemcpyHostToDevice));
  if (D->n_lights > 0) {
    if (strat == PT_LIGHTS_SPATIAL) {
      int rc = NC == 3 ? precompute_light_distribution<3>(s, n_vox, func, cdf, fint) : precompute_light_distribution<PT_NCS>(s, n_vox, func, cdf, fint);
      if (rc) return rc;
    } else {
      // uniform / power (lightdistrib/{uniform,power}.rs): one distribution, built on the host
      std::vector<float> f(D->n_lights, 1.0f);
      if (strat == PT_LIGHTS_POWER) {
        for (uint32_t i = 0; i < D->n_lights; i++) {
          const pt_light& L = D->lights[i];
          const float* sp = D->spectra + (size_t)L.spectrum * NC;
          float env_rgb[3];
          float k;
          if (L.type == PT_LIGHT_INFINITE) {  // infinite.rs:103-108: lookup((0.5, 0.5), 0.5) * pi r^2
            env_host[(size_t)L.tri].lookup(0.5f, 0.5f, 0.5f, env_rgb);
            sp = env_rgb;
            k = PT_PI * H.world_radius * H.world_radius;
          } else if (L.type == PT_LIGHT_AREA) {
            const float* p = D->tri_p + (size_t)L.tri * 9;
            f3 a = mk3(p[3] - p[0], p[4] - p[1], p[5] - p[2]), b = mk3(p[6] - p[0], p[7] - p[1], p[8] - p[2]);
            float area = 0.5f * len3(cross3(a, b));
            k = (L.two_sided ? 2.0f : 1.0f) * area * PT_PI;
          } else if (L.type == PT_LIGHT_PROJECTION) {  // projection.rs:113-117: map((0.5, 0.5), 0.5) * I * 2 pi (1 - cos_total_width)
            env_host[(size_t)L.tri].lookup(0.5f, 0.5f, 0.5f, env_rgb);
            for (int c = 0; c < 3; c++) env_rgb[c] = env_rgb[c] * sp[c];
            sp = env_rgb;
            k = 2.0f * PT_PI * (1.0f - D->envmaps[L.tri].cos_total_width);
          } else if (L.type == PT_LIGHT_GONIOMETRIC) {  // goniometric.rs:64-68: 4 pi * L * map((0.5, 0.5), 0.5)
            float c3[3];
            env_host[(size_t)L.tri].lookup(0.5f, 0.5f, 0.5f, c3);
            for (int c = 0; c < 3; c++) env_rgb[c] = ((4.0f * PT_PI) * sp[c]) * c3[c];
            sp = env_rgb;
            k = 1.0f;
          } else if (L.type == PT_LIGHT_POINT) k = 4.0f * PT_PI;
          else if (L.type == PT_LIGHT_SPOT) k = 2.0f * PT_PI * (1.0f - 0.5f * (D->spots[L.tri].cos_falloff_start - D->spots[L.tri].cos_total_width));  // spot.rs:82-85
          else k = PT_PI * H.world_radius * H.world_radius;
          const float* w = D->xyz_weights + 60;
          float acc = w[0] * (sp[0] * k);
          for (uint32_t c = 1; c < NC; c++) acc += w[c] * (sp[c] * k);
          f[i] = acc * D->xyz_scale;
        }
      }
      uint32_t nl = D->n_lights;
      std::vector<float> c(nl + 1, 0.0f);
      for (uint32_t j = 1; j < nl + 1; j++) c[j] = c[j - 1] + f[j - 1] / (float)nl;
      float fi = c[nl];
      if (fi == 0.0f) for (uint32_t j = 1; j < nl + 1; j++) c[j] = (float)j / (float)nl;
      else for (uint32_t j = 1; j < nl + 1; j++) c[j] /= fi;
      CK(cudaMemcpy(func, f.data(), nl * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(cdf, c.data(), (nl + 1) * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(fint, &fi, 4, cudaMemcpyHostToDevice));
    }
  }
  // film
  const int* cb = D->film.cropped_pixel_bounds;
  s->film_pixels = (uint32_t)(std::max(cb[2] - cb[0], 0) * std::max(cb[3] - cb[1], 0));
  CK(s->bufs.alloc(std::max(s->film_pixels, 1u), &s->film));
  CK(s->bufs.alloc((size_t)std::max(s->film_pixels, 1u) * 3, &s->rgb_dev));
  CK(cudaMemset(s->film, 0, (size_t)std::max(s->film_pixels, 1u) * sizeof(f4)));
  CK(s->bufs.alloc(4, &s->counters));
  CK(s->bufs.alloc(1, &s->overflow));
  CK(s->bufs.alloc(1, &s->fetch));
  CK(cudaMemset(s->counters, 0, 4 * sizeof(unsigned long long)));
  CK(cudaMemset(s->overflow, 0, sizeof(int)));
  return 0;
}

// A render that failed half way leaves queue counters, bins and pixel -> slot maps dirty: drop the buffers so that the
// next render on this scene allocates and initialises them again (ensure_wave).
void invalidate_wave_state(pt_scene* s) {
  cudaStreamSynchronize(s->stream);
  cudaGetLastError();
  s->wave_bufs.reset();
  s->slot_bufs.reset();
  s->wave_cap = 0;
  s->slot_segs = 0;
}

// Path state for `nseg` waves in flight of up to `seg_cap` paths each (one wave: nseg = 1).  Buffers only grow.
int ensure_wave(pt_scene* s, uint32_t seg_cap, uint32_t nseg = 1) {
  seg_cap = std::max(seg_cap, 1u);
  nseg = std::min(std::max(nseg, 1u), (uint32_t)PT_MAX_SEGMENTS);
  const uint64_t need = (uint64_t)seg_cap * nseg;
  if (need > 0x7fffffffull) { pt_set_error("wave state too large"); return 1; }
  const int* sb = s->h.film.sample_bounds;
  const size_t sbn = (size_t)std::max(sb[2] - sb[0], 1) * std::max(sb[3] - sb[1], 1);
  if (!s->slot_bufs || s->slot_segs < nseg) {  // one pixel -> slot map per wave in flight
    s->slot_bufs.reset(new DevBufs());
    CK(s->slot_bufs->alloc(sbn * nseg, &s->slot_of));
    CK(cudaMemset(s->slot_of, 0xff, sbn * nseg * sizeof(int)));
    s->slot_segs = nseg;
  }
  s->W.seg_cap = seg_cap;
  s->W.nseg = nseg;
  if (s->wave_bufs && s->wave_cap >= need) return 0;
  const uint32_t cap = (uint32_t)need;
  s->wave_bufs.reset(new DevBufs());
  DevBufs& B = *s->wave_bufs;
  WaveState& W = s->W;
  const uint32_t NC = s->h.n_channels;
  W.cap = cap;
  CK(B.alloc(cap, &W.ray_o)); CK(B.alloc(cap, &W.ray_d)); CK(B.alloc(cap, &W.hit));
  W.hit_inst = nullptr;
  if (s->h.has_instances) CK(B.alloc(cap, &W.hit_inst));
  const uint32_t SS = st_stride(NC);  // RGB: padded to 16-byte records
  CK(B.alloc((size_t)cap * SS, &W.beta)); CK(B.alloc((size_t)cap * SS, &W.L));
  CK(B.alloc(cap, &W.eta_scale)); CK(B.alloc(cap, &W.flags)); CK(B.alloc(cap, &W.dim)); CK(B.alloc(cap, &W.sindex));
  CK(B.alloc(cap, &W.pix)); CK(B.alloc(cap, &W.p_film));
  CK(B.alloc(cap, &W.sh_o)); CK(B.alloc(cap, &W.sh_d)); CK(B.alloc((size_t)cap * SS, &W.pendA));
  CK(B.alloc(cap, &W.mis_o)); CK(B.alloc(cap, &W.mis_d)); CK(B.alloc((size_t)cap * SS, &W.pendB)); CK(B.alloc(cap, &W.mis_light));
  Queues& Q = s->Q;
  CK(B.alloc(cap, &Q.ext)); CK(B.alloc(cap, &Q.ext_next)); CK(B.alloc(cap, &Q.shadow)); CK(B.alloc(cap, &Q.mis)); CK(B.alloc(32, &Q.counts));
  CK(cudaMemset(Q.counts, 0, 32 * sizeof(uint32_t)));
  CK(B.alloc((size_t)cap * CL_COUNT, &Q.bins));
  const int itype = s->h.integrator.type;
  const bool tree = itype == PT_INTEGRATOR_DIRECT || itype == PT_INTEGRATOR_WHITTED;
  W.sample_num = nullptr; W.arr_off = nullptr;
  W.n_frames = 0; W.fr_ro = W.fr_rd = W.fr_hit = W.fr_beta = nullptr; W.fr_inst = nullptr;
  Q.wait = Q.wait_next = nullptr;
  if (itype == PT_INTEGRATOR_AO || tree) { CK(B.alloc(cap, &W.sample_num)); CK(B.alloc(cap, &W.arr_off)); }
  if (tree) {  // the pending specular_transmit branches: one frame per depth below max_depth - 1 (ri_body)
    W.n_frames = (uint32_t)std::max(s->h.integrator.max_depth - 1, 1);
    const size_t nf = (size_t)cap * W.n_frames;
    CK(B.alloc(nf, &W.fr_ro)); CK(B.alloc(nf, &W.fr_rd)); CK(B.alloc(nf, &W.fr_hit)); CK(B.alloc(nf, &W.fr_beta));
    if (s->h.has_instances) CK(B.alloc(nf, &W.fr_inst));
    CK(B.alloc(cap, &Q.wait)); CK(B.alloc(cap, &Q.wait_next));
  }
  s->wave_cap = cap;
  return 0;
}

// k_shade for every shading class of the scene over the bins the last k_trace filled: forked onto the class streams
// after everything queued on s->stream so far, joined back before anything queued later.  Returns the launch count.
int shade_classes(pt_scene* s, int NC, int grid, unsigned long long* vc) {
  int n = 0;
#if PT_SHADE_STREAMS
  int live = 0;
  for (int c = 0; c < CL_COUNT; c++) live += (s->class_mask >> c) & 1;
  const bool fork = live > 1 && s->ev_fork != nullptr;
  if (fork) cudaEventRecord(s->ev_fork, s->stream);
  for (int c = 0; c < CL_COUNT; c++)
    if (s->class_mask & (1u << c)) {
      cudaStream_t q = fork ? s->cls_stream[c] : s->stream;
      if (fork) cudaStreamWaitEvent(q, s->ev_fork, 0);
      launch_shade(NC, c, grid, q, s->h, s->W, s->Q_cur, vc);
      if (fork) { cudaEventRecord(s->ev_join[c], q); cudaStreamWaitEvent(s->stream, s->ev_join[c], 0); }
      n++;
    }
#else
  for (int c = 0; c < CL_COUNT; c++)
    if (s->class_mask & (1u << c)) { launch_shade(NC, c, grid, s->stream, s->h, s->W, s->Q_cur, vc); n++; }
#endif
  return n;
}

// k_ri in the variant the scene needs: whitted or directlighting, two-lobe BSDFs unless it has uber / translucent materials,
// the infinite-light code only with an infinite light (the generic kernel is 10-20 % slower with either compiled in)
template <int NC>
void launch_ri(pt_scene* s, int grid, const WaveState& W, const Queues& Q, uint32_t n_ext, uint32_t n_wait, const uint32_t* prev, unsigned long long* vc) {
  if constexpr (NC == 3) {
    const bool wide = (s->class_mask & (1u << CL_EXTRA)) != 0, env = s->h.n_inf != 0 || s->h.has_map_lights;
#define PT_RI(ML, ENV, WH) PT_LAUNCH(grid, PT_SHADE_BLOCK, s->stream, k_ri<NC, ML, ENV, WH>)(s->h, W, Q, Q.wait, n_ext, n_wait, prev, vc)
#define PT_RI2(ML, ENV) do { if (wh) PT_RI(ML, ENV, true); else PT_RI(ML, ENV, false); } while (0)
    const bool wh = s->h.integrator.type == PT_INTEGRATOR_WHITTED;
    if (wide) { if (env) PT_RI2(PT_LOBES_EXTRA, true); else PT_RI2(PT_LOBES_EXTRA, false); }
    else { if (env) PT_RI2(PT_MAX_LOBES, true); else PT_RI2(PT_MAX_LOBES, false); }
#undef PT_RI2
#undef PT_RI
  }
}

struct WaveTotals {
  uint64_t regular = 0, shadow = 0, launches = 0, trace_launches = 0, trace_rays = 0;
  double trace_ms = 0, shade_ms = 0;
};

// runs the bounce loop for n paths whose camera rays are in the state and whose ids are in Q.ext[0..n)
template <int NC>
int run_bounces(pt_scene* s, uint32_t n, bool count, bool timed, WaveTotals* T) {
  Queues Q = s->Q;
  WaveState& W = s->W;
  uint32_t counts[4] = {n, 0, 0, 0};
  uint32_t n_wait = 0;  // directlighting / whitted: paths between two lighting rounds (no ray of their own to extend)
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
  const bool inst = s->h.has_instances != 0;
  int guard = 0;
  while (counts[0] + counts[2] + counts[3] + n_wait > 0) {
    if (++guard > 100000) { pt_set_error("bounce loop did not terminate"); return 1; }
    uint32_t total = counts[0] + counts[3] + counts[2];
    int g = grid_for(s, total, PT_TRACE_BLOCK, 8);
    CK(cudaMemsetAsync(Q.counts + 8, 0, 8 * sizeof(uint32_t), s->stream));  // shading-class bin sizes
    if (timed) cudaEventRecord(e0, s->stream);
#define LAUNCH_TRACE(K, I, C) PT_LAUNCH(g, PT_TRACE_BLOCK, s->stream, K<NC, I, C>)(s->h, W, Q, counts[0], counts[3], counts[2], s->counters, s->overflow, 0)
    if (s->h.accel_type == PT_ACCEL_LBVH) {
      if (inst) { if (count) LAUNCH_TRACE(k_trace_lbvh, true, true); else LAUNCH_TRACE(k_trace_lbvh, true, false); }
      else { if (count) LAUNCH_TRACE(k_trace_lbvh, false, true); else LAUNCH_TRACE(k_trace_lbvh, false, false); }
    } else {
      if (inst) { if (count) LAUNCH_TRACE(k_trace, true, true); else LAUNCH_TRACE(k_trace, true, false); }
      else { if (count) LAUNCH_TRACE(k_trace, false, true); else LAUNCH_TRACE(k_trace, false, false); }
    }
#undef LAUNCH_TRACE
    CK(cudaGetLastError());
    if (timed) cudaEventRecord(e1, s->stream);
    T->regular += counts[0] + counts[3];
    T->shadow += counts[2];
    T->launches++;
    T->trace_launches++;
    T->trace_rays += total;
    uint32_t n_ext = counts[0];
    CK(cudaMemsetAsync(Q.counts, 0, 8 * sizeof(uint32_t), s->stream));  // queue counters + the trace fetch counter
    if (n_ext > 0 && s->h.integrator.type == PT_INTEGRATOR_AOV) {  // aov.rs: no bounces, nothing is pushed
      PT_LAUNCH(grid_for(s, n_ext, 128, 8), 128, s->stream, k_aov<NC>)(s->h, W, Q.ext, n_ext);
      CK(cudaGetLastError());
      T->launches++;
    } else if (s->h.integrator.type == PT_INTEGRATOR_AO) {  // ao.rs: every path in flight makes its next occlusion ray
      if constexpr (NC == 3) {
        Queues Qo = Q;
        Qo.shadow = Q.mis;  // the queue this round fills; the one just traced is its input list
        PT_LAUNCH(grid_for(s, (size_t)counts[0] + counts[2], PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM), PT_SHADE_BLOCK, s->stream, k_ao<NC>)(s->h, W, Qo, Q.shadow, counts[0], counts[2], nullptr);
        CK(cudaGetLastError());
        std::swap(Q.shadow, Q.mis);
        T->launches++;
      }
    } else if (s->h.integrator.type == PT_INTEGRATOR_DIRECT || s->h.integrator.type == PT_INTEGRATOR_WHITTED) {
      if constexpr (NC == 3) {  // one step of the specular-tree walk of every path in flight
        const int gr = grid_for(s, (size_t)counts[0] + n_wait, PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM);
        unsigned long long* vc = count ? s->counters + 2 : nullptr;
        launch_ri<NC>(s, gr, W, Q, counts[0], n_wait, nullptr, vc);
        CK(cudaGetLastError());
        std::swap(Q.wait, Q.wait_next);
        T->launches++;
      }
    } else if (n_ext > 0) {
      int gs = grid_for(s, n_ext, PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM);
      unsigned long long* vc = count ? s->counters + 2 : nullptr;
      s->Q_cur = Q;
      T->launches += shade_classes(s, NC, gs, vc);
      CK(cudaGetLastError());
    }
    if (timed) cudaEventRecord(e2, s->stream);
    uint32_t c8[8];
    CK(cudaMemcpyAsync(c8, Q.counts, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < 4; k++) counts[k] = c8[k];
    n_wait = c8[5];
    if (timed) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, e0, e1);
      cudaEventElapsedTime(&b, e1, e2);
      T->trace_ms += a;
      T->shade_ms += b;
    }
    counts[0] = counts[1];
    counts[1] = 0;
    std::swap(Q.ext, Q.ext_next);
  }
  if (timed) { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); }
  return 0;
}

int finish_render(pt_scene* s, pt_stats* st, const WaveTotals& T, bool count) {
  CK(cudaStreamSynchronize(s->stream));
  int ovf = 0;
  CK(cudaMemcpy(&ovf, s->overflow, sizeof(int), cudaMemcpyDeviceToHost));
  if (ovf) { pt_set_error("traversal stack overflow"); return 1; }
  if (st) {
    st->regular_rays += T.regular;
    st->shadow_rays += T.shadow;
    st->kernel_launches += T.launches;
    st->extend_ms += T.trace_ms;
    st->shade_ms += T.shade_ms;
    st->extend_launches += T.trace_launches;
    st->extend_rays += T.trace_rays;
    if (count) {
      unsigned long long c[4];
      CK(cudaMemcpy(c, s->counters, sizeof(c), cudaMemcpyDeviceToHost));
      st->nodes_visited += c[0];
      st->tris_tested += c[1];
      st->path_vertices += c[2];
      // algorithmic bytes: one node per visit (128 B 4-wide / 32 B LinearBVHNode), 48 B per triangle test, 48 B ray + hit
      st->extend_bytes += (s->h.accel_type == PT_ACCEL_LBVH ? 32ull : 128ull) * c[0] + 48ull * c[1] + 48ull * T.trace_rays;
      CK(cudaMemset(s->counters, 0, 4 * sizeof(unsigned long long)));
    }
  }
  return 0;
}

int build_pixlist(pt_scene* s, int shard_index, int shard_count, int tile) {
  if (s->pix_key[0] == shard_index && s->pix_key[1] == shard_count && s->pix_key[2] == tile && s->pix_bufs) return 0;
  const int* sb = s->h.film.sample_bounds;
  int w = sb[2] - sb[0], h = sb[3] - sb[1];
  std::vector<int>& px = s->pix_host;
  px.clear();
  s->tiles_host.clear();
  int ntx = (w + tile - 1) / tile, nty = (h + tile - 1) / tile;
  for (int ty = 0; ty < nty; ty++)
    for (int tx = 0; tx < ntx; tx++) {
      int ti = ty * ntx + tx;
      if (shard_count > 1 && (ti % shard_count) != shard_index) continue;
      int first = (int)px.size();
      for (int y = ty * tile; y < std::min((ty + 1) * tile, h); y++)
        for (int x = tx * tile; x < std::min((tx + 1) * tile, w); x++) px.push_back((y << 16) | x);
      s->tiles_host.push_back(first);
      s->tiles_host.push_back((int)px.size() - first);
      s->tiles_host.push_back(ti);
    }
  s->pix_bufs.reset(new DevBufs());
  const int* dp = nullptr;
  CK(s->pix_bufs->upload(px.data(), px.size(), &dp));
  s->tiles = nullptr;
  s->h.tile_rng = nullptr;
  s->h.tile_sample = nullptr;
  if (s->h.sampler.type >= PT_SAMPLER_RANDOM && !s->tiles_host.empty()) {
    const int* dt = nullptr;
    CK(s->pix_bufs->upload(s->tiles_host.data(), s->tiles_host.size(), &dt));
    s->tiles = (int*)dt;
    CK(s->pix_bufs->alloc(s->tiles_host.size() / 3, &s->h.tile_rng));
    s->h.tile_arrays = nullptr;
    if (s->h.tile_arr_stride) CK(s->pix_bufs->alloc(s->tiles_host.size() / 3 * (size_t)s->h.tile_arr_stride, &s->h.tile_arrays));
    if (s->h.sampler.type >= PT_SAMPLER_STRATIFIED) CK(s->pix_bufs->alloc(s->tiles_host.size() / 3, &s->h.tile_sample));
  }
  s->pixlist = (int*)dp;
  s->pix_n = (uint32_t)px.size();
  s->pix_key[0] = shard_index; s->pix_key[1] = shard_count; s->pix_key[2] = tile;
  return 0;
}

// One wave = `ns` consecutive sample numbers of a contiguous chunk of the pixel list.
struct WaveJob {
  uint32_t c0, npix;
  int s0;
  uint32_t ns;
  int fx0, fy0, fx1, fy1;  // film footprint of the pixel chunk
};
struct WaveSeg {
  enum { IDLE = 0, ACTIVE = 1, DRAINING = 2, TRACED = 3 };
  int state = IDLE;
  int job = -1;
  int start_iter = 0;
  uint32_t map_c0 = 0xffffffffu, map_npix = 0;  // pixel chunk currently in this segment's slot_of map
};

// Integrator::render for the index-based samplers (halton, sobol): the bounce loop of li(), software-pipelined over
// `wave_segments` waves.  A wave's late bounces carry few rays (a 3 M-path wave of the 1 M-triangle scene ends with five
// iterations of < 2 % of its rays) and such launches are bound by the latency of single rays, not by throughput; so the
// next waves start in their own segments of the path state while the previous ones thin out, and every launch of
// k_trace / k_shade works on the union of their queues.  Paths never interact, and the film gathers the finished waves
// strictly in wave order, so the image does not depend on the pipelining.
template <int NC>
int render_waves(pt_scene* s, const pt_render_opts* o, pt_stats* st) {
  const int spp = s->h.sampler.spp;
  int s_begin = 0, s_end = spp;
  if (o && !(o->spp_begin == 0 && o->spp_end == 0)) { s_begin = std::max(o->spp_begin, 0); s_end = std::min(o->spp_end, spp); }
  int shard_i = o ? o->shard_index : 0, shard_n = (o && o->shard_count > 0) ? o->shard_count : 1;
  int tile = (o && o->tile_size > 0) ? o->tile_size : 16;
  bool count = o && o->collect_counters;
  const bool timed = st != nullptr;
  if (build_pixlist(s, shard_i, shard_n, tile)) return 1;
  uint32_t cap = (o && o->wave_paths > 0) ? (uint32_t)o->wave_paths : (NC == 3 ? PT_WAVE_PATHS_RGB : (1u << 19));
  if (!(o && o->wave_paths > 0) && (s->h.integrator.type == PT_INTEGRATOR_DIRECT || s->h.integrator.type == PT_INTEGRATOR_WHITTED)) {
    // the tree walk keeps max_depth - 1 frames of 68 bytes per path slot (ensure_wave): bound them to ~8 GB over the waves in flight
    const uint64_t frames = (uint64_t)std::max(s->h.integrator.max_depth - 1, 1);
    cap = (uint32_t)std::min<uint64_t>(cap, std::max<uint64_t>(1u << 18, (8ull << 30) / (frames * 68ull * 6ull)));
  }
  uint32_t npix_total = s->pix_n;
  int ns_total = std::max(s_end - s_begin, 0);
  const bool is_path = s->h.integrator.type == PT_INTEGRATOR_PATH;
  if (npix_total == 0 || ns_total == 0 || (is_path && s->h.n_lights == 0)) return 0;  // no lights: PathIntegrator::li returns zero (path.rs:71-75)
  uint32_t pix_chunk, ns_wave;
  if ((uint64_t)npix_total * (uint64_t)ns_total <= cap) { pix_chunk = npix_total; ns_wave = (uint32_t)ns_total; }
  else if (npix_total <= cap) { pix_chunk = npix_total; ns_wave = std::max(1u, cap / npix_total); }
  else {  // more sample pixels than path slots (4K films): equal pixel chunks, one sample number per wave
    const uint32_t n_chunks = (npix_total + cap - 1) / cap;
    pix_chunk = (npix_total + n_chunks - 1) / n_chunks;
    ns_wave = 1;
  }
  const int* sb = s->h.film.sample_bounds;
  const int* cb = s->h.film.cropped_pixel_bounds;
  const int sbw = sb[2] - sb[0];
  const size_t sbn = (size_t)std::max(sb[2] - sb[0], 1) * std::max(sb[3] - sb[1], 1);
  const float rx = s->h.film.filter_radius[0], ry = s->h.film.filter_radius[1];
  std::vector<WaveJob> jobs;
  for (uint32_t c0 = 0; c0 < npix_total; c0 += pix_chunk) {
    WaveJob J;
    J.c0 = c0;
    J.npix = std::min(pix_chunk, npix_total - c0);
    int minx = 1 << 30, miny = 1 << 30, maxx = -(1 << 30), maxy = -(1 << 30);
    for (uint32_t i = c0; i < c0 + J.npix; i++) {
      int pk = s->pix_host[i], x = (pk & 0xffff) + sb[0], y = (pk >> 16) + sb[1];
      minx = std::min(minx, x); maxx = std::max(maxx, x); miny = std::min(miny, y); maxy = std::max(maxy, y);
    }
    J.fx0 = std::max((int)floorf((float)minx - rx), cb[0]); J.fy0 = std::max((int)floorf((float)miny - ry), cb[1]);
    J.fx1 = std::min((int)ceilf((float)(maxx + 1) + rx), cb[2]); J.fy1 = std::min((int)ceilf((float)(maxy + 1) + ry), cb[3]);
    for (int s0 = s_begin; s0 < s_end; s0 += (int)ns_wave) {
      J.s0 = s0;
      J.ns = (uint32_t)std::min<int>((int)ns_wave, s_end - s0);
      jobs.push_back(J);
    }
  }
  const uint32_t seg_cap = (uint32_t)std::min<uint64_t>((uint64_t)pix_chunk * ns_wave, cap);
  uint32_t nseg = (o && o->wave_segments > 0) ? (uint32_t)o->wave_segments : 6u;  // measured: profiles/r1_wave_pipelining.md
  nseg = (uint32_t)std::min<size_t>(std::min(nseg, (uint32_t)PT_MAX_SEGMENTS), jobs.size());
  if (ensure_wave(s, seg_cap, nseg)) return 1;
  Queues Q = s->Q;
  WaveState& W = s->W;
  const bool inst = s->h.has_instances != 0;
  const bool is_aov = s->h.integrator.type == PT_INTEGRATOR_AOV;
  const bool is_ao = s->h.integrator.type == PT_INTEGRATOR_AO;  // paths live on as occlusion rays: the shadow queue counts as in flight
  const bool is_ri = s->h.integrator.type == PT_INTEGRATOR_DIRECT || s->h.integrator.type == PT_INTEGRATOR_WHITTED;  // ... or on the wait list
  WaveTotals T;
  std::vector<WaveSeg> segs(nseg);
  std::vector<int> seg_of_job(jobs.size(), -1);
  size_t next_job = 0, next_film = 0;
  int iter = 0, last_start = -(1 << 30), wave_iters = is_ao ? s->h.integrator.ao_samples + 2 : s->h.integrator.max_depth + 3;
  int rc = 0;
  // The host never waits for the bounce it has just queued.  Queue sizes stay on the device: k_raygen appends the
  // camera rays behind the continuation rays of the last shade stage, k_trace reads the three queue sizes, k_shade its
  // bin sizes.  The counters of iteration t are copied to pinned memory behind shade(t) and looked at only after iteration
  // t + 1 has been queued, so the GPU always has the next bounce waiting.  Until they arrive the host works with upper
  // bounds (a path pushes at most one continuation, one shadow and one MIS ray per vertex) for the launch shapes, and a
  // wave is retired one iteration later than it could be - which changes when k_finish / k_film run, not what they add.
  struct Pending { bool any = false; int iter = 0, slot = 0; uint32_t n_new = 0; bool timed = false; } pend;
  uint32_t e_last = 0;  // continuation + camera rays traced by the last queued iteration (exact once its predecessor is resolved)
  // k_finish + k_film for the waves whose rays are all traced, in wave order
  auto gather_ready = [&]() -> int {
    while (next_film < jobs.size() && seg_of_job[next_film] >= 0 && segs[seg_of_job[next_film]].state == WaveSeg::TRACED) {
      const int k = seg_of_job[next_film];
      const WaveJob& J = jobs[next_film];
      const uint32_t n = J.npix * J.ns, base = (uint32_t)k * seg_cap;
      PT_LAUNCH(grid_for(s, n, 256, PT_STREAM_GRID_PER_SM), 256, s->stream, k_finish<NC>)(s->h, W, base, n);
      CK(cudaGetLastError());
      T.launches++;
      if (J.fx1 > J.fx0 && J.fy1 > J.fy0) {
        PT_LAUNCH(grid_for(s, (size_t)(J.fx1 - J.fx0) * (J.fy1 - J.fy0), 128, PT_STREAM_GRID_PER_SM), 128, s->stream, k_film<NC>)(s->h, W, s->slot_of + (size_t)k * sbn, base, J.npix, J.ns, J.fx0,
                                                                                                             J.fy0, J.fx1, J.fy1, s->film);
        CK(cudaGetLastError());
        T.launches++;
      }
      wave_iters = std::max(iter - segs[k].start_iter, 1);
      segs[k].state = WaveSeg::IDLE;
      segs[k].job = -1;
      next_film++;
    }
    return 0;
  };
  // the counters of the pending iteration have arrived: ray totals of the iteration queued after it, wave retirement
  auto resolve = [&](uint32_t n_new_after) -> int {
    if (!pend.any) return 0;
    if (cudaEventSynchronize(s->ev_iter[pend.slot]) != cudaSuccess) { pt_set_error(std::string("bounce loop: ") + cudaGetErrorString(cudaGetLastError())); return 1; }
    const uint32_t* c = s->counts_host + 32 * pend.slot;
    if (pend.timed) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, s->ev_time[pend.slot][0], s->ev_time[pend.slot][1]);
      cudaEventElapsedTime(&b, s->ev_time[pend.slot][2], s->ev_time[pend.slot][3]);
      T.trace_ms += a;
      T.shade_ms += b;
    }
    // what shade(pend.iter) pushed is what the next iteration traces (it is queued already, or nothing was pushed)
    T.regular += c[1] + c[3];
    T.shadow += c[2];
    T.trace_rays += c[1] + c[2] + c[3];
    for (uint32_t k = 0; k < nseg; k++)
      if (segs[k].state == WaveSeg::ACTIVE && segs[k].start_iter <= pend.iter && (nseg > 1 ? c[16 + k] : c[1] + (is_ao ? c[2] : 0u) + (is_ri ? c[5] : 0u)) == 0) segs[k].state = WaveSeg::TRACED;
    e_last = c[1] + (is_ao ? c[2] : 0u) + (is_ri ? c[5] : 0u) + n_new_after;
    pend.any = false;
    return gather_ready();
  };
  for (;;) {
    if (++iter > 10000000) { pt_set_error("bounce loop did not terminate"); rc = 1; break; }
    // (1) start the next wave in an idle segment; the starts are spread over a wave's life (wave_iters iterations)
    bool any_busy = false;
    for (auto& g : segs) any_busy = any_busy || g.state != WaveSeg::IDLE;
    const int stagger = std::max(1, wave_iters / (int)nseg);
    uint32_t n_new = 0;
    if (next_job < jobs.size() && (!any_busy || iter - last_start >= stagger)) {
      for (uint32_t k = 0; k < nseg; k++) {
        WaveSeg& g = segs[k];
        if (g.state != WaveSeg::IDLE) continue;
        const WaveJob& J = jobs[next_job];
        const uint32_t n = J.npix * J.ns, base = k * seg_cap;
        int* map = s->slot_of + (size_t)k * sbn;
        if (g.map_c0 != J.c0 || g.map_npix != J.npix) {
          if (g.map_npix > 0) {
            PT_LAUNCH(grid_for(s, g.map_npix, 256, 8), 256, s->stream, k_set_slots)(s->pixlist, g.map_c0, g.map_npix, sbw, map, 1);
            T.launches++;
          }
          PT_LAUNCH(grid_for(s, J.npix, 256, 8), 256, s->stream, k_set_slots)(s->pixlist, J.c0, J.npix, sbw, map, 0);
          if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_set_slots launch failed"); break; }
          T.launches++;
          g.map_c0 = J.c0; g.map_npix = J.npix;
        }
        PT_LAUNCH(grid_for(s, n, PT_SHADE_BLOCK, PT_STREAM_GRID_PER_SM), PT_SHADE_BLOCK, s->stream, k_raygen<NC>)(s->h, W, Q, s->pixlist, J.c0, J.npix, J.s0, J.ns, base, 0u, Q.counts + 1);
        if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_raygen launch failed"); break; }
        T.launches++;
        if (st) st->camera_rays += n;
        n_new = n;
        g.state = WaveSeg::ACTIVE;
        g.job = (int)next_job;
        g.start_iter = iter;
        seg_of_job[next_job] = (int)k;
        next_job++;
        last_start = iter;
        break;
      }
      if (rc) break;
    }
    // upper bounds of the three queues of this iteration (e_last = the vertices the pending shade stage had at most)
    const uint32_t ub_ext = (pend.any ? e_last : 0u) + n_new, ub_side = pend.any ? e_last : 0u;
    if (ub_ext == 0 && ub_side == 0) {
      // nothing can be in flight: every started wave has been retired by the last resolve
      if ((rc = resolve(0u))) break;
      bool busy = false;
      for (auto& g : segs) busy = busy || g.state != WaveSeg::IDLE;
      if (!busy && next_job >= jobs.size()) break;
      if (busy) { pt_set_error("bounce loop: a wave is active without rays in flight"); rc = 1; break; }
      continue;
    }
    // (2) one bounce of every wave in flight: trace the three queues, shade the continuation hits
    const int slot = iter & 1;
    int g = grid_for(s, (size_t)ub_ext + 2 * (size_t)ub_side, PT_TRACE_BLOCK, 8);
    if ((rc = (cudaMemsetAsync(Q.counts + 8, 0, 8 * sizeof(uint32_t), s->stream) != cudaSuccess))) { pt_set_error("cudaMemsetAsync failed"); break; }  // shading-class bin sizes
    if (timed) cudaEventRecord(s->ev_time[slot][0], s->stream);
#define LAUNCH_TRACE(K, I, C) PT_LAUNCH(g, PT_TRACE_BLOCK, s->stream, K<NC, I, C>)(s->h, W, Q, n_new, 0u, 0u, s->counters, s->overflow, 1)
    if (s->h.accel_type == PT_ACCEL_LBVH) {
      if (inst) { if (count) LAUNCH_TRACE(k_trace_lbvh, true, true); else LAUNCH_TRACE(k_trace_lbvh, true, false); }
      else { if (count) LAUNCH_TRACE(k_trace_lbvh, false, true); else LAUNCH_TRACE(k_trace_lbvh, false, false); }
    } else {
      if (inst) { if (count) LAUNCH_TRACE(k_trace, true, true); else LAUNCH_TRACE(k_trace, true, false); }
      else { if (count) LAUNCH_TRACE(k_trace, false, true); else LAUNCH_TRACE(k_trace, false, false); }
    }
#undef LAUNCH_TRACE
    if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_trace launch failed"); break; }
    if (timed) { cudaEventRecord(s->ev_time[slot][1], s->stream); cudaEventRecord(s->ev_time[slot][2], s->stream); }
    T.regular += n_new;
    T.trace_rays += n_new;
    T.launches++;
    T.trace_launches++;
    if (is_ao || is_ri) cudaMemcpyAsync(Q.counts + 24, Q.counts + 1, 5 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s->stream);  // k_ao's / k_ri's input list sizes
    cudaMemsetAsync(Q.counts, 0, 8 * sizeof(uint32_t), s->stream);  // queue counters + the trace fetch counter
    if ((rc = (cudaMemsetAsync(Q.counts + 16, 0, PT_MAX_SEGMENTS * sizeof(uint32_t), s->stream) != cudaSuccess))) { pt_set_error("cudaMemsetAsync failed"); break; }
    if (is_aov) {  // aov.rs: no bounces, nothing is pushed: only the camera rays of this iteration are in the queue
      if (n_new > 0) {
        PT_LAUNCH(grid_for(s, n_new, 128, 8), 128, s->stream, k_aov<NC>)(s->h, W, Q.ext, n_new);
        if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_aov launch failed"); break; }
        T.launches++;
      }
    } else if (is_ao) {  // ao.rs: every path in flight makes its next occlusion ray
      if constexpr (NC == 3) {
        Queues Qo = Q;
        Qo.shadow = Q.mis;  // the queue this round fills; the one just traced is its input list
        PT_LAUNCH(grid_for(s, (size_t)ub_ext + ub_side, PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM), PT_SHADE_BLOCK, s->stream, k_ao<NC>)(s->h, W, Qo, Q.shadow, n_new, 0u, Q.counts + 24);
        if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_ao launch failed"); break; }
        std::swap(Q.shadow, Q.mis);
        T.launches++;
      }
    } else if (is_ri) {  // directlighting / whitted: one step of the specular-tree walk of every path in flight
      if constexpr (NC == 3) {
        const int gr = grid_for(s, (size_t)ub_ext + ub_side, PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM);
        unsigned long long* vc = count ? s->counters + 2 : nullptr;
        launch_ri<NC>(s, gr, W, Q, n_new, 0u, Q.counts + 24, vc);
        if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_ri launch failed"); break; }
        std::swap(Q.wait, Q.wait_next);
        T.launches++;
      }
    } else if (ub_ext > 0) {
      int gs = grid_for(s, ub_ext, PT_SHADE_BLOCK, PT_SHADE_GRID_PER_SM);
      unsigned long long* vc = count ? s->counters + 2 : nullptr;
      s->Q_cur = Q;
      T.launches += shade_classes(s, NC, gs, vc);
      if ((rc = (cudaGetLastError() != cudaSuccess))) { pt_set_error("k_shade launch failed"); break; }
    }
    if (timed) cudaEventRecord(s->ev_time[slot][3], s->stream);
    if ((rc = (cudaMemcpyAsync(s->counts_host + 32 * slot, Q.counts, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess))) { pt_set_error("cudaMemcpyAsync failed"); break; }
    if ((rc = (cudaEventRecord(s->ev_iter[slot], s->stream) != cudaSuccess))) { pt_set_error("cudaEventRecord failed"); break; }
    std::swap(Q.ext, Q.ext_next);
    // (3) the iteration queued BEFORE this one: its counters say what this one traces and which waves it completes
    const bool had = pend.any;
    if ((rc = resolve(n_new))) break;
    if (!had) e_last = n_new;
    pend.any = true; pend.iter = iter; pend.slot = slot; pend.n_new = n_new; pend.timed = timed;
  }
  if (!rc && pend.any) rc = resolve(0u);
  if (rc) { invalidate_wave_state(s); return 1; }
  // leave the pixel -> slot maps empty for the next render
  for (uint32_t k = 0; k < nseg; k++)
    if (segs[k].map_npix > 0) {
      PT_LAUNCH(grid_for(s, segs[k].map_npix, 256, 8), 256, s->stream, k_set_slots)(s->pixlist, segs[k].map_c0, segs[k].map_npix, sbw, s->slot_of + (size_t)k * sbn, 1);
      CK(cudaGetLastError());
      T.launches++;
    }
  return finish_render(s, st, T, count);
}

// Sampler "random" (samplers/random.rs): the reference gives every 16x16 render tile its own PCG32 sequence
// (sampler.rs:285, seed = tile number) and runs the tile's pixels and samples one after the other, each path drawing a
// data-dependent count of numbers.  Reproducing the streams therefore serialises the paths of a tile: wave (q, n) carries
// sample n of pixel q of every tile, all tiles in parallel; the film is gathered once per tile group as usual.
template <int NC>
int render_waves_random(pt_scene* s, const pt_render_opts* o, pt_stats* st) {
  const int spp = s->h.sampler.spp;
  if (o && !(o->spp_begin == 0 && o->spp_end == 0) && !(o->spp_begin <= 0 && o->spp_end >= spp)) {
    pt_set_error("Sampler \"random\": a sample-number range cannot be rendered on its own (sequential per-tile streams); shard by tiles");
    return 1;
  }
  int shard_i = o ? o->shard_index : 0, shard_n = (o && o->shard_count > 0) ? o->shard_count : 1;
  bool count = o && o->collect_counters;
  if (build_pixlist(s, shard_i, shard_n, 16)) return 1;  // render tiles are 16x16 (sampler.rs:267)
  const uint32_t n_tiles = (uint32_t)(s->tiles_host.size() / 3);
  if (n_tiles == 0 || spp <= 0 || (s->h.integrator.type == PT_INTEGRATOR_PATH && s->h.n_lights == 0)) return 0;
  {  // RNG::new_sequence(tile number) (core/rng.rs:21-33)
    std::vector<uint64_t> init(n_tiles);
    for (uint32_t t = 0; t < n_tiles; t++) {
      Pcg r;
      r.state = 0;
      r.inc = ((uint64_t)(uint32_t)s->tiles_host[3 * t + 2] << 1) | 1ull;
      r.next();
      r.state += 0x853c49e6748fea9bULL;
      r.next();
      init[t] = r.state;
    }
    CK(cudaMemcpy(s->h.tile_rng, init.data(), n_tiles * sizeof(uint64_t), cudaMemcpyHostToDevice));
  }
  uint32_t cap = (o && o->wave_paths > 0) ? (uint32_t)o->wave_paths : (NC == 3 ? PT_WAVE_PATHS_RGB : (1u << 19));
  cap = std::max<uint32_t>(cap, 256u * (uint32_t)spp);
  const int* sb = s->h.film.sample_bounds;
  const int* cb = s->h.film.cropped_pixel_bounds;
  const int sbw = sb[2] - sb[0];
  const float rx = s->h.film.filter_radius[0], ry = s->h.film.filter_radius[1];
  WaveTotals T;
  for (uint32_t t0 = 0; t0 < n_tiles;) {
    uint32_t t1 = t0, npix = 0, maxq = 0;
    while (t1 < n_tiles && (uint64_t)(npix + (uint32_t)s->tiles_host[3 * t1 + 1]) * (uint64_t)spp <= cap) {
      npix += (uint32_t)s->tiles_host[3 * t1 + 1];
      maxq = std::max(maxq, (uint32_t)s->tiles_host[3 * t1 + 1]);
      t1++;
    }
    const uint32_t c0 = (uint32_t)s->tiles_host[3 * t0];
    const uint32_t n = npix * (uint32_t)spp;
    if (ensure_wave(s, n)) return 1;
    PT_LAUNCH(grid_for(s, npix, 256, 8), 256, s->stream, k_set_slots)(s->pixlist, c0, npix, sbw, s->slot_of, 0);
    CK(cudaGetLastError());
    T.launches++;
    int minx = 1 << 30, miny = 1 << 30, maxx = -(1 << 30), maxy = -(1 << 30);
    for (uint32_t i = c0; i < c0 + npix; i++) {
      int pk = s->pix_host[i], x = (pk & 0xffff) + sb[0], y = (pk >> 16) + sb[1];
      minx = std::min(minx, x); maxx = std::max(maxx, x); miny = std::min(miny, y); maxy = std::max(maxy, y);
    }
    int fx0 = std::max((int)floorf((float)minx - rx), cb[0]), fy0 = std::max((int)floorf((float)miny - ry), cb[1]);
    int fx1 = std::min((int)ceilf((float)(maxx + 1) + rx), cb[2]), fy1 = std::min((int)ceilf((float)(maxy + 1) + ry), cb[3]);
    for (uint32_t q = 0; q < maxq; q++) {
      uint32_t active = 0;
      for (uint32_t t = t0; t < t1; t++) active += (uint32_t)s->tiles_host[3 * t + 1] > q ? 1u : 0u;
      for (int smp = 0; smp < spp; smp++) {
        CK(cudaMemsetAsync(s->Q.counts + 5, 0, sizeof(uint32_t), s->stream));
        PT_LAUNCH(grid_for(s, t1 - t0, 128, 8), 128, s->stream, k_raygen_random<NC>)(s->h, s->W, s->Q, s->pixlist, s->tiles, t0, t1 - t0, c0, npix, q, smp);
        CK(cudaGetLastError());
        T.launches++;
        if (st) st->camera_rays += active;
        if (run_bounces<NC>(s, active, count, st != nullptr, &T)) return 1;
      }
    }
    PT_LAUNCH(grid_for(s, n, 256, 8), 256, s->stream, k_finish<NC>)(s->h, s->W, 0u, n);
    CK(cudaGetLastError());
    T.launches++;
    if (fx1 > fx0 && fy1 > fy0) {
      PT_LAUNCH(grid_for(s, (size_t)(fx1 - fx0) * (fy1 - fy0), 128, 8), 128, s->stream, k_film<NC>)(s->h, s->W, s->slot_of, 0u, npix, (uint32_t)spp, fx0, fy0, fx1, fy1, s->film);
      CK(cudaGetLastError());
      T.launches++;
    }
    PT_LAUNCH(grid_for(s, npix, 256, 8), 256, s->stream, k_set_slots)(s->pixlist, c0, npix, sbw, s->slot_of, 1);
    CK(cudaGetLastError());
    T.launches++;
    t0 = t1;
  }
  return finish_render(s, st, T, count);
}

int launch_intersect(pt_scene* s, const f4* ro, const f4* rd, uint32_t n, f4* hits, float* tt, unsigned char* occ, bool count) {
  CK(cudaMemsetAsync(s->fetch, 0, sizeof(uint32_t), s->stream));
  int g = grid_for(s, n, PT_TRACE_BLOCK, 8);
#define LAUNCH_ISECT(K, I, C) PT_LAUNCH(g, PT_TRACE_BLOCK, s->stream, K<I, C>)(s->h, ro, rd, n, hits, tt, occ, s->fetch, s->counters, s->overflow)
  if (s->h.accel_type == PT_ACCEL_LBVH) {
    if (s->h.has_instances) { if (count) LAUNCH_ISECT(k_intersect_lbvh, true, true); else LAUNCH_ISECT(k_intersect_lbvh, true, false); }
    else { if (count) LAUNCH_ISECT(k_intersect_lbvh, false, true); else LAUNCH_ISECT(k_intersect_lbvh, false, false); }
  } else {
    if (s->h.has_instances) { if (count) LAUNCH_ISECT(k_intersect, true, true); else LAUNCH_ISECT(k_intersect, true, false); }
    else { if (count) LAUNCH_ISECT(k_intersect, false, true); else LAUNCH_ISECT(k_intersect, false, false); }
  }
#undef LAUNCH_ISECT
  CK(cudaGetLastError());
  return 0;
}

int upload_rays(pt_scene* s, const float* rays, uint32_t n, DevBufs& B, f4** ro, f4** rd) {
  std::vector<f4> o(n), d(n);
  for (uint32_t i = 0; i < n; i++) {
    const float* q = rays + (size_t)i * 8;
    o[i] = mk4(q[0], q[1], q[2], q[3]);
    d[i] = mk4(q[4], q[5], q[6], q[7]);
  }
  const f4 *po = nullptr, *pd = nullptr;
  CK(B.upload(o.data(), n, &po));
  CK(B.upload(d.data(), n, &pd));
  *ro = (f4*)po; *rd = (f4*)pd;
  (void)s;
  return 0;
}

}  // namespace

extern "C" {

int pt_scene_create(const pt_scene_desc* desc, int device, pt_scene** out) {
  if (!desc || !out) { pt_set_error("pt_scene_create: null argument"); return 1; }
  *out = nullptr;
  if (desc->abi_version != PT_ABI_VERSION) { pt_set_error("pt_scene_create: ABI version mismatch"); return 1; }
  if (desc->n_channels != 3 && desc->n_channels != PT_NCS) { pt_set_error("pt_scene_create: n_channels must be 3 or 60"); return 1; }
  if (desc->n_objects < 1) { pt_set_error("pt_scene_create: need at least the top-level object"); return 1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { pt_set_error(std::string("pt_scene_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback"); return 1; }
  CK(cudaSetDevice(device));
  std::unique_ptr<pt_scene> s(new pt_scene());
  s->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  s->n_sm = prop.multiProcessorCount;
  CK(cudaStreamCreate(&s->stream));
  CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  for (int c = 0; c < CL_COUNT; c++) {
    CK(cudaStreamCreateWithFlags(&s->cls_stream[c], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s->ev_join[c], cudaEventDisableTiming));
  }
  CK(cudaMallocHost((void**)&s->counts_host, 2 * 32 * sizeof(uint32_t)));
  for (int k = 0; k < 2; k++) {
    CK(cudaEventCreateWithFlags(&s->ev_iter[k], cudaEventDisableTiming));
    for (int j = 0; j < 4; j++) CK(cudaEventCreate(&s->ev_time[k][j]));
  }
  if (build_scene(s.get(), desc)) return 1;
  *out = s.release();
  return 0;
}

void pt_scene_destroy(pt_scene* scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  delete scene;
}

int pt_scene_reupload(pt_scene* s, uint64_t* bytes) {
  if (!s) { pt_set_error("null scene"); return 1; }
  CK(cudaSetDevice(s->device));
  uint64_t total = 0;
  for (auto& st : s->bufs.staged) {
    CK(cudaMemcpyAsync(st.dev, st.host, st.bytes, cudaMemcpyHostToDevice, s->stream));
    total += st.bytes;
  }
  CK(cudaMemcpyAsync(s->d, &s->h, sizeof(DScene), cudaMemcpyHostToDevice, s->stream));
  total += sizeof(DScene);
  if (permute_recs(s)) return 1;
  CK(cudaStreamSynchronize(s->stream));
  if (bytes) *bytes = total;
  return 0;
}

int pt_film_clear(pt_scene* s) {
  if (!s) { pt_set_error("null scene"); return 1; }
  CK(cudaSetDevice(s->device));
  CK(cudaMemsetAsync(s->film, 0, (size_t)std::max(s->film_pixels, 1u) * sizeof(f4), s->stream));
  return 0;
}

int pt_render_partial(pt_scene* s, const pt_render_opts* opts, float** dev_film_xyzw, pt_stats* stats) {
  if (!s) { pt_set_error("null scene"); return 1; }
  CK(cudaSetDevice(s->device));
  if (stats) memset(stats, 0, sizeof(*stats));
  if (pt_film_clear(s)) return 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s->stream));
  int rc;
  if (s->h.sampler.type >= PT_SAMPLER_RANDOM) rc = s->h.n_channels == 3 ? render_waves_random<3>(s, opts, stats) : render_waves_random<PT_NCS>(s, opts, stats);
  else rc = s->h.n_channels == 3 ? render_waves<3>(s, opts, stats) : render_waves<PT_NCS>(s, opts, stats);
  if (rc) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    invalidate_wave_state(s);
    return rc;
  }
  CK(cudaEventRecord(e1, s->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (stats) stats->device_ms = ms;
  if (dev_film_xyzw) *dev_film_xyzw = (float*)s->film;
  return 0;
}

int pt_film_resolve(pt_scene* s, const float* dev_film_xyzw, float* out_rgb_host) {
  if (!s || !out_rgb_host) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  const f4* film = dev_film_xyzw ? (const f4*)dev_film_xyzw : s->film;
  if (s->film_pixels == 0) return 0;
  PT_LAUNCH(grid_for(s, s->film_pixels, 256, 8), 256, s->stream, k_film_resolve)(s->h, film, s->rgb_dev, s->film_pixels);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_rgb_host, s->rgb_dev, (size_t)s->film_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int pt_render(pt_scene* s, const pt_render_opts* opts, float* out_rgb, pt_stats* stats) {
  float* film = nullptr;
  if (pt_render_partial(s, opts, &film, stats)) return 1;
  if (stats) stats->kernel_launches += 1;
  return pt_film_resolve(s, film, out_rgb);
}

int pt_intersect(pt_scene* s, const float* rays, uint32_t n, int32_t* prim, float* t, float* bary, uint64_t* nodes_visited, uint64_t* tris_tested) {
  if (!s || !rays) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  if (n == 0) return 0;
  DevBufs B;
  f4 *ro, *rd, *hits;
  float* tt;
  if (upload_rays(s, rays, n, B, &ro, &rd)) return 1;
  CK(B.alloc(n, &hits));
  CK(B.alloc(n, &tt));
  bool count = nodes_visited || tris_tested;
  CK(cudaMemsetAsync(s->counters, 0, 4 * sizeof(unsigned long long), s->stream));
  int g = grid_for(s, n, PT_TRACE_BLOCK, 8);
  if (launch_intersect(s, ro, rd, n, hits, tt, nullptr, count)) return 1;
  CK(cudaGetLastError());
  std::vector<f4> hh(n);
  std::vector<float> th(n);
  CK(cudaMemcpyAsync(hh.data(), hits, n * sizeof(f4), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaMemcpyAsync(th.data(), tt, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  unsigned long long c[4] = {0, 0, 0, 0};
  int ovf = 0;
  CK(cudaMemcpyAsync(c, s->counters, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaMemcpyAsync(&ovf, s->overflow, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (ovf) { pt_set_error("traversal stack overflow"); return 1; }
  for (uint32_t i = 0; i < n; i++) {
    int tri = (int)fbits(hh[i].w);
    if (prim) prim[i] = tri;
    if (t) t[i] = th[i];
    if (bary) { bary[3 * i] = tri >= 0 ? hh[i].x : 0.0f; bary[3 * i + 1] = tri >= 0 ? hh[i].y : 0.0f; bary[3 * i + 2] = tri >= 0 ? hh[i].z : 0.0f; }
  }
  if (nodes_visited) *nodes_visited = c[0];
  if (tris_tested) *tris_tested = c[1];
  CK(cudaMemset(s->counters, 0, 4 * sizeof(unsigned long long)));
  return 0;
}

int pt_intersect_p(pt_scene* s, const float* rays, uint32_t n, uint8_t* occluded) {
  if (!s || !rays || !occluded) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  if (n == 0) return 0;
  DevBufs B;
  f4 *ro, *rd;
  unsigned char* occ;
  if (upload_rays(s, rays, n, B, &ro, &rd)) return 1;
  CK(B.alloc(n, &occ));
  int g = grid_for(s, n, PT_TRACE_BLOCK, 8);
  if (launch_intersect(s, ro, rd, n, nullptr, nullptr, occ, false)) return 1;
  CK(cudaMemcpyAsync(occluded, occ, n, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int pt_intersect_device(pt_scene* s, const void* dev_o_tmax, const void* dev_d_time, uint32_t n, void* dev_hits, int repeat, float* ms_per_launch) {
  if (!s || !dev_o_tmax || !dev_d_time || !dev_hits) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int g = grid_for(s, n, PT_TRACE_BLOCK, 8);
  CK(cudaEventRecord(e0, s->stream));
  for (int r = 0; r < std::max(repeat, 1); r++) {
    if (launch_intersect(s, (const f4*)dev_o_tmax, (const f4*)dev_d_time, n, (f4*)dev_hits, nullptr, nullptr, false)) return 1;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(e1, s->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_per_launch) *ms_per_launch = ms / (float)std::max(repeat, 1);
  return 0;
}

int pt_sampler_values(pt_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, uint32_t dim0, uint32_t ndims, float* out) {
  if (!s || !pixel_xy || !sample_num || !out) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  if (s->h.sampler.type >= PT_SAMPLER_RANDOM) { pt_set_error("Samplers \"random\" / \"stratified\" have no index-based stream: only pt_render reproduces them (per-tile sequential PCG32)"); return 1; }
  if (n == 0 || ndims == 0) return 0;
  DevBufs B;
  const int *pxy, *sn;
  float* o;
  CK(B.upload(pixel_xy, (size_t)n * 2, &pxy));
  CK(B.upload(sample_num, n, &sn));
  CK(B.alloc((size_t)n * ndims, &o));
  PT_LAUNCH(grid_for(s, (size_t)n * ndims, 128, 8), 128, s->stream, k_sampler_values)(s->h, pxy, sn, n, dim0, ndims, o);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, o, (size_t)n * ndims * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

static int run_list(pt_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, bool bounce) {
  if (ensure_wave(s, std::max(n, 1u))) return 1;
  DevBufs B;
  const int *pxy, *sn;
  CK(B.upload(pixel_xy, (size_t)n * 2, &pxy));
  CK(B.upload(sample_num, n, &sn));
  if (s->h.n_channels == 3) PT_LAUNCH(grid_for(s, n, 128, 8), 128, s->stream, k_raygen_list<3>)(s->h, s->W, s->Q, pxy, sn, n);
  else PT_LAUNCH(grid_for(s, n, 128, 8), 128, s->stream, k_raygen_list<PT_NCS>)(s->h, s->W, s->Q, pxy, sn, n);
  CK(cudaGetLastError());
  if (bounce && (s->h.n_lights > 0 || s->h.integrator.type != PT_INTEGRATOR_PATH)) {
    WaveTotals T;
    int rc = s->h.n_channels == 3 ? run_bounces<3>(s, n, false, false, &T) : run_bounces<PT_NCS>(s, n, false, false, &T);
    if (rc) { invalidate_wave_state(s); return rc; }
    if (s->h.n_channels == 3) PT_LAUNCH(grid_for(s, n, 256, 8), 256, s->stream, k_finish<3>)(s->h, s->W, 0u, n);
    else PT_LAUNCH(grid_for(s, n, 256, 8), 256, s->stream, k_finish<PT_NCS>)(s->h, s->W, 0u, n);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int pt_camera_rays(pt_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* rays, float* p_film) {
  if (!s || !pixel_xy || !sample_num || !rays) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  if (s->h.sampler.type >= PT_SAMPLER_RANDOM) { pt_set_error("Samplers \"random\" / \"stratified\" have no index-based stream: only pt_render reproduces them (per-tile sequential PCG32)"); return 1; }
  if (n == 0) return 0;
  if (run_list(s, pixel_xy, sample_num, n, false)) return 1;
  std::vector<f4> o(n), d(n);
  std::vector<f2> pf(n);
  CK(cudaMemcpy(o.data(), s->W.ray_o, n * sizeof(f4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(d.data(), s->W.ray_d, n * sizeof(f4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pf.data(), s->W.p_film, n * sizeof(f2), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n; i++) {
    float* q = rays + (size_t)i * 8;
    q[0] = o[i].x; q[1] = o[i].y; q[2] = o[i].z; q[3] = o[i].w; q[4] = d[i].x; q[5] = d[i].y; q[6] = d[i].z; q[7] = d[i].w;
    if (p_film) { p_film[2 * i] = pf[i].x; p_film[2 * i + 1] = pf[i].y; }
  }
  return 0;
}

int pt_li(pt_scene* s, const int32_t* pixel_xy, const int32_t* sample_num, uint32_t n, float* out_l) {
  if (!s || !pixel_xy || !sample_num || !out_l) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  if (s->h.sampler.type >= PT_SAMPLER_RANDOM) { pt_set_error("Samplers \"random\" / \"stratified\" have no index-based stream: only pt_render reproduces them (per-tile sequential PCG32)"); return 1; }
  if (n == 0) return 0;
  if (run_list(s, pixel_xy, sample_num, n, true)) return 1;
  const uint32_t NC = s->h.n_channels, cap = s->W.cap;
  std::vector<float> L((size_t)cap * st_stride(NC));
  CK(cudaMemcpy(L.data(), s->W.L, L.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n; i++)
    for (uint32_t c = 0; c < NC; c++) out_l[(size_t)i * NC + c] = L[(size_t)i * st_stride(NC) + c];  // path-major records
  int ovf = 0;
  CK(cudaMemcpy(&ovf, s->overflow, sizeof(int), cudaMemcpyDeviceToHost));
  if (ovf) { pt_set_error("traversal stack overflow"); return 1; }
  return 0;
}

int pt_bvh_info(pt_scene* s, uint32_t object, uint32_t* n_nodes, uint32_t* n_prims, int32_t* ordered) {
  if (!s || object >= s->objects.size()) { pt_set_error("bad object"); return 1; }
  const ObjectHost& O = s->objects[object];
  std::vector<pth::RefNodeH> ref;
  if (n_nodes) { pth::flatten_reference(O.bvh, &ref); *n_nodes = (uint32_t)ref.size(); }
  if (n_prims) *n_prims = (uint32_t)O.ordered.size();
  if (ordered) for (size_t i = 0; i < O.ordered.size(); i++) ordered[i] = O.ordered[i];
  return 0;
}

int pt_bvh_node(pt_scene* s, uint32_t object, uint32_t node, float* boxes24, int64_t* children4, int32_t* axes3, int32_t* is_leaf) {
  if (!s || object >= s->objects.size()) { pt_set_error("bad object"); return 1; }
  static thread_local uint64_t cached_uid = 0;
  static thread_local uint32_t cached_obj = 0;
  static thread_local std::vector<pth::RefNodeH> ref;
  if (cached_uid != s->uid || cached_obj != object || ref.empty()) {
    pth::flatten_reference(s->objects[object].bvh, &ref);
    cached_uid = s->uid;
    cached_obj = object;
  }
  if (node >= ref.size()) { pt_set_error("bad node"); return 1; }
  const pth::RefNodeH& q = ref[node];
  memcpy(boxes24, q.bboxes, sizeof(q.bboxes));
  for (int k = 0; k < 4; k++) children4[k] = q.children[k];
  for (int k = 0; k < 3; k++) axes3[k] = q.axes[k];
  *is_leaf = q.is_leaf;
  return 0;
}

int pt_light_distribution(pt_scene* s, const float* points, uint32_t n, float* out_func) {
  if (!s || !points || !out_func) { pt_set_error("null argument"); return 1; }
  CK(cudaSetDevice(s->device));
  const uint32_t nl = s->h.n_lights;
  if (nl == 0 || n == 0) return 0;
  const LightDistDev& D = s->h.ld;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t vox = D.strategy == PT_LIGHTS_SPATIAL ? ld_voxel(D, mk3(points[3 * i], points[3 * i + 1], points[3 * i + 2])) : 0u;
    CK(cudaMemcpy(out_func + (size_t)i * nl, D.func + (size_t)vox * nl, nl * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return 0;
}

}  // extern "C"

#if defined(PT_HOSTCHECK)
#include "shade_tu.cu"  // tests/hostcheck builds one translation unit
#endif
