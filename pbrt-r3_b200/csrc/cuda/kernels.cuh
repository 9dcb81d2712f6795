// kernels.cuh — __global__ wrappers around the PT_HD stage bodies (dshade.cuh, dtrav2.cuh, dfilm.cuh).
// Launch shapes: k_trace = persistent warps with dynamic ray fetch (PT_TRACE_MINBLOCKS x 148 SMs on B200); the
// streaming kernels (raygen, shade, finish, film) get about one item per thread (launch.cuh).
#pragma once
#include "launch.cuh"
#include "dfilm.cuh"
#include "dtrav2.cuh"
#include "dlbvh.cuh"

namespace ptd {

#define PT_TRACE_BLOCK 128
#ifndef PT_TRACE_MINBLOCKS
#define PT_TRACE_MINBLOCKS 7
#endif

// Traversal stack: the first PT_SMEM_STACK entries of every thread live in shared memory, strided so that
// consecutive lanes hit consecutive banks ([entry][thread]); deeper entries spill to a small local array.
#ifndef PT_SMEM_STACK
#define PT_SMEM_STACK 24
#endif
// On the device the column is addressed with explicit st.shared / ld.shared on a 32-bit shared-space address (the
// generic-pointer form cost an LDL of the pointer + a 64-bit IMAD.WIDE + a generic ST per push); the spill array is a
// separate local array so that `n` and the address stay in registers.
#ifndef PT_OPT_STS
#define PT_OPT_STS 1
#endif
#if defined(__CUDA_ARCH__) && PT_OPT_STS
#define PT_STS_ASM 1
#else
#define PT_STS_ASM 0
#endif
// Words per thread of the instance save area: the world-space ray and everything trav_set_ray derives from it, parked
// in shared memory while the lane walks an instanced object (instead of 8 live registers + recomputing 6 IEEE divisions
// on the way out).
#define PT_SAVE_WORDS 15
struct SharedStack {
#if PT_STS_ASM
  uint32_t sa;       // shared-space byte address of smem[threadIdx.x]; entry stride = PT_TRACE_BLOCK * 4 bytes
#else
  int* sm;           // tests/hostcheck: &smem[0]
#endif
  int n;
  int* spill;        // PT_MAX_STACK - PT_SMEM_STACK entries of local memory
  float* sv;         // save area column: sv[k * PT_TRACE_BLOCK], k < PT_SAVE_WORDS (null when the scene has no instances)
  __device__ __forceinline__ void attach(int* smem_col, int* spill_arr, float* save_col = nullptr) {
#if PT_STS_ASM
    sa = (uint32_t)__cvta_generic_to_shared(smem_col);
#else
    sm = smem_col;
#endif
    spill = spill_arr;
    sv = save_col;
    n = 0;
  }
  __device__ __forceinline__ void save(int k, float v) { sv[k * PT_TRACE_BLOCK] = v; }
  __device__ __forceinline__ float saved(int k) const { return sv[k * PT_TRACE_BLOCK]; }
  // fast path of the node phase: k entries fit into the shared part -> put(pos, x) at n + pos, then advance(k)
  __device__ __forceinline__ bool room(int k) const { return n + k <= PT_SMEM_STACK; }
  __device__ __forceinline__ void put(int pos, int x) {
#if PT_STS_ASM
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa + (uint32_t)(n + pos) * (PT_TRACE_BLOCK * 4u)), "r"(x) : "memory");
#else
    sm[(n + pos) * PT_TRACE_BLOCK] = x;
#endif
  }
  __device__ __forceinline__ void advance(int k) { n += k; }
  __device__ __forceinline__ bool push(int x) {
    if (n < PT_SMEM_STACK) {
#if PT_STS_ASM
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa + (uint32_t)n * (PT_TRACE_BLOCK * 4u)), "r"(x) : "memory");
#else
      sm[n * PT_TRACE_BLOCK] = x;
#endif
    } else if (n < PT_MAX_STACK) spill[n - PT_SMEM_STACK] = x;
    else return false;
    n++;
    return true;
  }
  __device__ __forceinline__ bool pop(int* x) {
    if (n == 0) return false;
    n--;
    if (n < PT_SMEM_STACK) {
#if PT_STS_ASM
      int v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sa + (uint32_t)n * (PT_TRACE_BLOCK * 4u)) : "memory");
      *x = v;
#else
      *x = sm[n * PT_TRACE_BLOCK];
#endif
    } else *x = spill[n - PT_SMEM_STACK];
    return true;
  }
};

// Persistent-warp traversal driver.  Every lane owns one ray at a time; when PT_REFILL_MIN or more lanes of a warp are
// idle the warp claims that many new rays with ONE atomicAdd on the launch's fetch counter (ballot / popc ranks),
// so lanes do not wait for the slowest ray of a fixed 32-ray batch.  IO supplies the rays and consumes the results:
//   void load(uint32_t k, Trav& T)   -> trav_begin(...)        void store(uint32_t k, const Trav& T)
#ifndef PT_REFILL_MIN
#define PT_REFILL_MIN 16
#endif
// two-level scenes: rays take longer and differ more in length, refilling earlier pays (profiles/r1_instanced_trace.md)
#ifndef PT_REFILL_MIN_INST
#define PT_REFILL_MIN_INST 8
#endif
#ifndef PT_NODE_EXIT_MIN
#define PT_NODE_EXIT_MIN 8
#endif
#ifndef PT_VOTE_AGE
#define PT_VOTE_AGE 2
#endif
// majority-vote state machine for two-level scenes: measured SLOWER than while-while (profiles/r1_instanced_trace.md), off
#ifndef PT_INST_VOTE
#define PT_INST_VOTE 0
#endif
template <bool INST, bool COUNT, typename IO>
__device__ __forceinline__ void trace_persistent(const DScene& S, IO& io, uint32_t total, uint32_t* fetch, SharedStack& st, int* ovf, TravCounters* tc) {
  Trav T;
  T.done = true;
#if defined(__CUDA_ARCH__)
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  constexpr int REFILL = INST ? PT_REFILL_MIN_INST : PT_REFILL_MIN;
  bool have = false, exhausted = false;
  uint32_t myk = 0;
  for (;;) {
    unsigned idle = __ballot_sync(FULL, !have);
    int nidle = __popc(idle);
    if (!exhausted && nidle >= REFILL) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(fetch, (uint32_t)nidle);
      base = __shfl_sync(FULL, base, 0);
      if (!have) {
        uint32_t k = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if (k < total) {
          myk = k;
          st.n = 0;
          io.load(k, T);
          have = true;
          if (T.done) { io.store(k, T); have = false; }
        }
      }
      if (base + (uint32_t)nidle >= total) exhausted = true;
    }
    unsigned active = __ballot_sync(FULL, have);
    if (active == 0) {
      if (exhausted) break;
      continue;
    }
    if (INST && PT_INST_VOTE) {
      // Two-level scenes: a lane is walking nodes (0), has a triangle leaf of an instanced object (1), a top-level leaf
      // whose records are mostly instances to enter - ray transform, object-bounds test (2) - or the barrier that ends an
      // instanced object (3).  Running "the leaf phase" for all of them at once serialises three different code paths at
      // 3-5 active lanes each (ncu, instanced 20 M-triangle scene: 8.3 active lanes per instruction overall).  So every
      // step the warp votes and runs only the state most lanes are in; the others wait for company.
      // and a state that was passed over gains PT_VOTE_AGE votes per step, so that no lane waits for long.
      int w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      for (;;) {
        const int stt = (!have || T.done) ? 4 : (T.lc == 0 ? 0 : (T.lc < 0 ? 3 : (T.inst >= 0 ? 1 : 2)));
        const unsigned m0 = __ballot_sync(FULL, stt == 0), m1 = __ballot_sync(FULL, stt == 1), m2 = __ballot_sync(FULL, stt == 2), m3 = __ballot_sync(FULL, stt == 3);
        const int c0 = __popc(m0), c1 = __popc(m1), c2 = __popc(m2), c3 = __popc(m3);
        if ((c0 | c1 | c2 | c3) == 0) break;
        const int v0 = c0 ? c0 + w0 : 0, v1 = c1 ? c1 + w1 : 0, v2 = c2 ? c2 + w2 : 0, v3 = c3 ? c3 + w3 : 0;
        int pick = 0, best = v0;
        if (v1 > best) { pick = 1; best = v1; }
        if (v2 > best) { pick = 2; best = v2; }
        if (v3 > best) { pick = 3; best = v3; }
        w0 = (pick == 0 || !c0) ? 0 : w0 + PT_VOTE_AGE;
        w1 = (pick == 1 || !c1) ? 0 : w1 + PT_VOTE_AGE;
        w2 = (pick == 2 || !c2) ? 0 : w2 + PT_VOTE_AGE;
        w3 = (pick == 3 || !c3) ? 0 : w3 + PT_VOTE_AGE;
        if (pick == 0) {
          if (stt == 0) trav_node_step<INST, COUNT>(S, T, st, ovf, tc);
        } else if (stt == pick) {
          trav_leaf<INST, COUNT>(S, T, st, ovf, tc);
          T.lc = 0;
        }
        if (have && T.done) { io.store(myk, T); have = false; }
        unsigned a = __ballot_sync(FULL, have);
        if (a == 0) break;
        if (!exhausted && __popc(a) <= 32 - REFILL) break;
      }
      continue;
    }
    for (;;) {
      // while-while: all lanes run slab tests until each has a leaf (or is finished), then all run triangle tests.
      // The node loop is warp-synchronous (one vote per step): the lanes of a warp meet again after EVERY node, whatever
      // branches the step took (left to the compiler, lanes that reach the loop latch early run ahead and are never
      // merged back: measured 7 instead of 13 active lanes per instruction).
      {
        bool inn = have;
        for (;;) {
          unsigned m = __ballot_sync(FULL, inn);
          if (m == 0u) break;
#if PT_NODE_EXIT_MIN > 0
          // few lanes still walking nodes while others wait with a leaf: let the leaf phase run first
          if (__popc(m) < PT_NODE_EXIT_MIN && __any_sync(FULL, have && !T.done && T.lc != 0)) break;
#endif
          if (inn) inn = trav_node_step<INST, COUNT>(S, T, st, ovf, tc);
        }
      }
      if (have && !T.done && T.lc != 0) {
        trav_leaf<INST, COUNT>(S, T, st, ovf, tc);
        T.lc = 0;
      }
      if (have && T.done) { io.store(myk, T); have = false; }
      unsigned a = __ballot_sync(FULL, have);
      if (a == 0) break;
      if (!exhausted && __popc(a) <= 32 - REFILL) break;
    }
  }
#else
  (void)fetch;
  for (uint32_t k = 0; k < total; k++) {  // tests/hostcheck: serial
    st.n = 0;
    io.load(k, T);
    while (!T.done) {
      trav_nodes<INST, COUNT>(S, T, st, ovf, tc);
      if (!T.done) trav_leaf<INST, COUNT>(S, T, st, ovf, tc);
    }
    io.store(k, T);
  }
#endif
}

// rays and results of the wavefront: continuation (A) rays, then MIS (C) rays, then shadow (B) rays
template <int NC, bool INST>
struct WaveIO {
  const DScene& S;
  WaveState& W;
  Queues& Q;
  uint32_t n_ext, n_mis;
  // ray k of the launch: o, d, tmax, time; returns true for an any-hit (shadow) query
  __device__ __forceinline__ bool ray(uint32_t k, f3* o3, f3* d3, float* tmax, float* time) {
    if (k < n_ext + n_mis) {
      const bool is_ext = k < n_ext;
      uint32_t i = is_ext ? Q.ext[k] : Q.mis[k - n_ext];
      f4 o = is_ext ? W.ray_o[i] : W.mis_o[i], d = is_ext ? W.ray_d[i] : W.mis_d[i];
      *o3 = xyz(o); *d3 = xyz(d); *tmax = o.w; *time = d.w;
      return false;
    }
    uint32_t i = Q.shadow[k - n_ext - n_mis];
    f4 o = W.sh_o[i], d = W.sh_d[i];
    *o3 = xyz(o); *d3 = xyz(d); *tmax = d.w; *time = W.ray_d[i].w;
    return true;
  }
  __device__ __forceinline__ void load(uint32_t k, Trav& T) {
    f3 o, d;
    float tmax, time;
    const bool any = ray(k, &o, &d, &tmax, &time);
    trav_begin(S, T, o, d, tmax, time, any);
  }
  __device__ __forceinline__ void store(uint32_t k, const Trav& T) {
    if (k < n_ext) {
      uint32_t i = Q.ext[k];
      const int tri = T.found ? T.hit.tri : -1;
      W.hit[i] = mk4(T.hit.b0, T.hit.b1, T.hit.b2, bitsf((uint32_t)tri));
      if (INST) W.hit_inst[i] = T.found ? T.hit.instance : -1;
      if (T.found) W.ray_o[i].w = T.ray_tmax;
      // A ray that leaves the scene ends its path (without infinite lights path.rs:93-104 adds nothing), and the
      // pending light / MIS terms of its last vertex are folded into L by k_finish in the same order (L + a) + b that
      // the shade stage would use: nothing to shade.  With infinite lights the escaped ray goes to class CL_NONE for Le.
      if (tri < 0 && S.n_inf == 0u) return;
      // bin the path by shading class for the per-material shade kernels (warp-aggregated per class)
      const int cls = shade_class(S, tri, (int)(W.flags[i] >> 8));
      uint32_t slot;
#if defined(__CUDA_ARCH__)
      {
        unsigned m = __activemask();
        unsigned peers = __match_any_sync(m, cls);
        int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&Q.counts[8 + cls], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        slot = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      }
#else
      slot = Q.counts[8 + cls]++;
#endif
      Q.bins[(size_t)cls * W.cap + slot] = i;
    } else if (k < n_ext + n_mis) {
      uint32_t i = Q.mis[k - n_ext];
      mis_resolve<NC>(S, W, i, T.hit, T.found, xyz(W.mis_d[i]));
    } else {
      uint32_t i = Q.shadow[k - n_ext - n_mis];
      if (T.found) st_store<NC>(W.pendA, W.cap, i, sp_const<NC>(0.0f));
    }
  }
};

// Scene::intersect for continuation (A) and MIS (C) rays, Scene::intersect_p for shadow (B) rays, in one launch.
template <int NC, bool INST, bool COUNT>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, PT_TRACE_MINBLOCKS) k_trace(const __grid_constant__ DScene S, WaveState W, Queues Q, uint32_t n_ext, uint32_t n_mis,
                                                          uint32_t n_sh, unsigned long long* counters, int* overflow, int dev_counts) {
  // dev_counts: the queue sizes are what the last shade stage left in Q.counts (the host does not wait for them,
  // render_waves); n_ext then only holds the camera rays appended since
  if (dev_counts) { n_ext += Q.counts[1]; n_mis = Q.counts[3]; n_sh = Q.counts[2]; }
  __shared__ int smem[PT_SMEM_STACK * PT_TRACE_BLOCK];
  __shared__ float ssave[INST ? PT_SAVE_WORDS * PT_TRACE_BLOCK : 1];
  int spill[PT_MAX_STACK - PT_SMEM_STACK];
  SharedStack st;
  st.attach(smem + threadIdx.x, spill, INST ? ssave + threadIdx.x : nullptr);
  TravCounters tc;
  tc.nodes = 0; tc.tris = 0;
  int ovf = 0;
  WaveIO<NC, INST> io{S, W, Q, n_ext, n_mis};
  trace_persistent<INST, COUNT>(S, io, n_ext + n_mis + n_sh, Q.counts + 4, st, &ovf, &tc);
  if (COUNT) {
    atomicAdd(&counters[0], tc.nodes);
    atomicAdd(&counters[1], tc.tris);
  }
  if (ovf) *overflow = 1;
}

template <int NC>
__global__ void __launch_bounds__(PT_SHADE_BLOCK) k_raygen(const __grid_constant__ DScene S, WaveState W, Queues Q, const int* __restrict__ pixlist,
                                                           uint32_t c0, uint32_t npix, int s0, uint32_t ns, uint32_t base, uint32_t q0, const uint32_t* q0_dev) {
  // path slots base .. base + npix * ns (the wave's segment of the state), appended to the extend queue at q0 (+ the
  // number of continuation rays the last shade stage pushed, read on the device)
  if (q0_dev) q0 += *q0_dev;
  const uint32_t total = npix * ns;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t slot = i % npix, s = i / npix;
    int pk = pixlist[c0 + slot];
    int px = (pk & 0xffff) + S.sampler.sb[0], py = (pk >> 16) + S.sampler.sb[1];
    raygen_body<NC>(S, W, Q, base + i, px, py, s0 + (int)s, -1, 0u, q0 + i);
  }
}
// Sampler "random", wave w of a tile group: one path per tile = pixel number w / spp of the tile (scan order, as
// render_tile visits them), sample number w % spp.  tiles[t] = {first pixel in pixlist, pixel count, tile sequence number}.
template <int NC>
__global__ void k_raygen_random(const __grid_constant__ DScene S, WaveState W, Queues Q, const int* __restrict__ pixlist, const int* __restrict__ tiles,
                                uint32_t tile0, uint32_t n_tiles, uint32_t c0, uint32_t npix, uint32_t q, int sample) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
    const int* tr = tiles + 3 * (size_t)(tile0 + t);
    if (q >= (uint32_t)tr[1]) continue;
    uint32_t slot = (uint32_t)tr[0] + q - c0;
    int pk = pixlist[c0 + slot];
    int px = (pk & 0xffff) + S.sampler.sb[0], py = (pk >> 16) + S.sampler.sb[1];
    if (S.sampler.type >= PT_SAMPLER_STRATIFIED) {  // start_pixel: tables and Latin-hypercube arrays of the pixel; current_pixel_sample_index
      if (sample == 0) {
        const bool all = S.integrator.type == PT_INTEGRATOR_DIRECT && S.integrator.direct_strategy == PT_DIRECT_ALL;
        const uint32_t n_arr = S.integrator.type == PT_INTEGRATOR_AO ? 1u : (all ? 2u * S.n_lights * (uint32_t)S.integrator.max_depth : 0u);
        if (S.sampler.type == PT_SAMPLER_STRATIFIED) stratified_start_pixel(S, tile0 + t, (uint32_t)tr[2], n_arr, all ? S.light_sample_prefix : nullptr, (uint32_t)S.integrator.ao_samples);
        else zerotwo_start_pixel(S, tile0 + t, (uint32_t)tr[2], n_arr, all ? S.light_sample_prefix : nullptr, (uint32_t)S.integrator.ao_samples);
      }
      S.tile_sample[tile0 + t] = (uint32_t)sample;
    } else if (sample == 0 && S.tile_arr_stride) {  // start_pixel: the pixel's sample arrays come first in the tile's stream
      float* a = S.tile_arrays + (size_t)(tile0 + t) * S.tile_arr_stride;
      for (uint32_t k = 0; k < S.tile_arr_stride; k++) a[k] = random_draw(S, tile0 + t, (uint32_t)tr[2]);
    }
    uint32_t qpos = atomicAdd(&Q.counts[5], 1u);
    raygen_body<NC>(S, W, Q, (uint32_t)sample * npix + slot, px, py, sample, (int)(tile0 + t), (uint32_t)tr[2], qpos);
  }
}
// explicit (pixel, sample number) list: pt_li / pt_camera_rays
template <int NC>
__global__ void k_raygen_list(const __grid_constant__ DScene S, WaveState W, Queues Q, const int* __restrict__ pixel_xy, const int* __restrict__ sample_num,
                              uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    raygen_body<NC>(S, W, Q, i, pixel_xy[2 * i], pixel_xy[2 * i + 1], sample_num[i]);
}

// Integrator "aov": the whole li() after the camera ray's closest-hit query
template <int NC>
__global__ void k_aov(const __grid_constant__ DScene S, WaveState W, const uint32_t* __restrict__ ids, uint32_t n) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) aov_body<NC>(S, W, ids[k]);
}

// Integrator "ambientocclusion": one round of every path in flight (ao_body).  The paths are those whose camera /
// pass-through ray the last k_trace answered (Q.ext[0..n_ext)) and those whose occlusion ray it answered (sh_in[0..n_sh));
// Q.shadow is the queue this round fills.  prev = {n_ext, n_sh} as the last shade stage left them on the device (the host
// does not wait for them, render_waves), n_ext then only holds the camera rays appended since.
template <int NC>
__global__ void __launch_bounds__(PT_SHADE_BLOCK) k_ao(const __grid_constant__ DScene S, WaveState W, Queues Q, const uint32_t* __restrict__ sh_in, uint32_t n_ext, uint32_t n_sh,
                                                       const uint32_t* prev) {
  if (prev) { n_ext += prev[0]; n_sh = prev[1]; }
  const uint32_t total = n_ext + n_sh;
  // whole warps run the loop body: the queue appends are warp-aggregated over the active mask
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) ao_body<NC>(S, W, Q, k < n_ext ? Q.ext[k] : sh_in[k - n_ext]);
}

// Integrators "directlighting" and "whitted": one step of every path in flight (ri_body) - the vertices the last k_trace
// found (Q.ext[0..n_ext): continuation / camera rays) and the paths between two lighting rounds of a vertex
// (wait_in[0..n_wait)); Q.wait_next is the list this step fills.  prev = the copy of Q.counts[1..6) taken after the trace
// stage (render_waves: the host does not wait for the sizes), n_ext then only holds the camera rays appended since.
#ifndef PT_RI_MINBLOCKS
#define PT_RI_MINBLOCKS 3
#endif
template <int NC, int ML, bool ENV, bool WHITTED>
__global__ void __launch_bounds__(PT_SHADE_BLOCK, PT_RI_MINBLOCKS) k_ri(const __grid_constant__ DScene S, WaveState W, Queues Q, const uint32_t* __restrict__ wait_in, uint32_t n_ext, uint32_t n_wait,
                                                       const uint32_t* prev, unsigned long long* vertex_counter) {
  if (prev) { n_ext += prev[0]; n_wait = prev[4]; }
  const uint32_t total = n_ext + n_wait;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const bool fresh = k < n_ext;
    ri_body<NC, ML, ENV, WHITTED>(S, W, Q, fresh ? Q.ext[k] : wait_in[k - n_ext], fresh, vertex_counter);
  }
}

template <int NC>
__global__ void k_finish(const __grid_constant__ DScene S, WaveState W, uint32_t base, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) finish_body<NC>(S, W, base + i);
}

// all[0..n) = (y,z,x,w) and all[n..2n) = (z,x,y,w) of the uploaded records all[2n..3n) (dscene.cuh: recs_perm)
__global__ void k_permute_recs(f4* all, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const f4 v = all[2 * (size_t)n + i];
    // instance records hold a world-space box, read in x, y, z order whatever the copy (trav_leaf)
    const bool inst = (fbits(all[2 * (size_t)n + (i / 3u) * 3u + 1u].w) & PT_REC_INSTANCE) != 0u;
    all[i] = inst ? v : mk4(v.y, v.z, v.x, v.w);
    all[(size_t)n + i] = inst ? v : mk4(v.z, v.x, v.y, v.w);
  }
}

__global__ void k_set_slots(const int* __restrict__ pixlist, uint32_t c0, uint32_t npix, int sbw, int* slot_of, int clear) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
    int pk = pixlist[c0 + i];
    slot_of[(pk >> 16) * sbw + (pk & 0xffff)] = clear ? -1 : (int)i;
  }
}

template <int NC>
__global__ void k_film(const __grid_constant__ DScene S, WaveState W, const int* __restrict__ slot_of, uint32_t base, uint32_t npix, uint32_t ns, int x0, int y0,
                       int x1, int y1, f4* film) {
  int w = x1 - x0, h = y1 - y0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x)
    film_gather_body<NC>(S, W, slot_of, base, npix, ns, x0 + i % w, y0 + i / w, film);
}

__global__ void k_film_resolve(const __grid_constant__ DScene S, const f4* film, float* rgb, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) film_resolve_body(S.film, film[i], rgb + 3 * (size_t)i);
}

// ---- stage entry points ----------------------------------------------------------------------------------------
struct ArrayIO {  // pt_intersect / pt_intersect_p / pt_intersect_device: rays and results in plain arrays
  const DScene& S;
  const f4* ro;
  const f4* rd;
  f4* hits;
  float* t_out;
  unsigned char* occ;
  bool any;
  __device__ __forceinline__ bool ray(uint32_t k, f3* o3, f3* d3, float* tmax, float* time) {
    f4 o = ro[k], d = rd[k];
    *o3 = xyz(o); *d3 = xyz(d); *tmax = o.w; *time = d.w;
    return any;
  }
  __device__ __forceinline__ void load(uint32_t k, Trav& T) {
    f4 o = ro[k], d = rd[k];
    trav_begin(S, T, xyz(o), xyz(d), o.w, d.w, any);
  }
  __device__ __forceinline__ void store(uint32_t k, const Trav& T) {
    if (any) { occ[k] = T.found ? 1 : 0; return; }
    hits[k] = mk4(T.hit.b0, T.hit.b1, T.hit.b2, bitsf((uint32_t)(T.found ? T.hit.tri : -1)));
    if (t_out) t_out[k] = T.found ? T.hit.t : 0.0f;
  }
};
template <bool INST, bool COUNT>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, PT_TRACE_MINBLOCKS) k_intersect(const __grid_constant__ DScene S, const f4* __restrict__ ro, const f4* __restrict__ rd, uint32_t n,
                                                              f4* hits, float* t_out, unsigned char* occ, uint32_t* fetch, unsigned long long* counters,
                                                              int* overflow) {
  __shared__ int smem[PT_SMEM_STACK * PT_TRACE_BLOCK];
  __shared__ float ssave[INST ? PT_SAVE_WORDS * PT_TRACE_BLOCK : 1];
  int spill[PT_MAX_STACK - PT_SMEM_STACK];
  SharedStack st;
  st.attach(smem + threadIdx.x, spill, INST ? ssave + threadIdx.x : nullptr);
  TravCounters tc;
  tc.nodes = 0; tc.tris = 0;
  int ovf = 0;
  ArrayIO io{S, ro, rd, hits, t_out, occ, occ != nullptr};
  trace_persistent<INST, COUNT>(S, io, n, fetch, st, &ovf, &tc);
  if (COUNT) {
    atomicAdd(&counters[0], tc.nodes);
    atomicAdd(&counters[1], tc.tris);
  }
  if (ovf) *overflow = 1;
}

// ---- LinearBVHNode layout (dlbvh.cuh): one ray per lane, 32 rays per warp fetch ------------------------------------
template <bool INST, bool COUNT, typename IO>
__device__ __forceinline__ void trace_lbvh(const DScene& S, IO& io, uint32_t total, uint32_t* fetch, int* ovf, TravCounters* tc) {
  LStack st;
  st.n = 0;
#if defined(__CUDA_ARCH__)
  const int lane = threadIdx.x & 31;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(fetch, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= total) break;
    const uint32_t k = base + (uint32_t)lane;
    if (k >= total) continue;
#else
  (void)fetch;
  for (uint32_t k = 0; k < total; k++) {
#endif
    Trav T;
    f3 o, d;
    float tmax, time;
    const bool any = io.ray(k, &o, &d, &tmax, &time);
    T.hit.tri = -1; T.hit.instance = -1;
    T.hit.t = 0.0f; T.hit.b0 = T.hit.b1 = T.hit.b2 = 0.0f;
    T.ray_tmax = tmax;
    st.n = 0;
    if (any) T.found = lbvh_traverse<true, INST, COUNT, 0>(S, 0u, o, d, time, T.ray_tmax, &T.hit, st, ovf, tc);
    else T.found = lbvh_traverse<false, INST, COUNT, 0>(S, 0u, o, d, time, T.ray_tmax, &T.hit, st, ovf, tc);
    io.store(k, T);
  }
}
template <int NC, bool INST, bool COUNT>
__global__ void __launch_bounds__(PT_TRACE_BLOCK) k_trace_lbvh(const __grid_constant__ DScene S, WaveState W, Queues Q, uint32_t n_ext, uint32_t n_mis, uint32_t n_sh,
                                                               unsigned long long* counters, int* overflow, int dev_counts) {
  if (dev_counts) { n_ext += Q.counts[1]; n_mis = Q.counts[3]; n_sh = Q.counts[2]; }
  TravCounters tc;
  tc.nodes = 0; tc.tris = 0;
  int ovf = 0;
  WaveIO<NC, INST> io{S, W, Q, n_ext, n_mis};
  trace_lbvh<INST, COUNT>(S, io, n_ext + n_mis + n_sh, Q.counts + 4, &ovf, &tc);
  if (COUNT) {
    atomicAdd(&counters[0], tc.nodes);
    atomicAdd(&counters[1], tc.tris);
  }
  if (ovf) *overflow = 1;
}
template <bool INST, bool COUNT>
__global__ void __launch_bounds__(PT_TRACE_BLOCK) k_intersect_lbvh(const __grid_constant__ DScene S, const f4* __restrict__ ro, const f4* __restrict__ rd, uint32_t n, f4* hits,
                                                                   float* t_out, unsigned char* occ, uint32_t* fetch, unsigned long long* counters, int* overflow) {
  TravCounters tc;
  tc.nodes = 0; tc.tris = 0;
  int ovf = 0;
  ArrayIO io{S, ro, rd, hits, t_out, occ, occ != nullptr};
  trace_lbvh<INST, COUNT>(S, io, n, fetch, &ovf, &tc);
  if (COUNT) {
    atomicAdd(&counters[0], tc.nodes);
    atomicAdd(&counters[1], tc.tris);
  }
  if (ovf) *overflow = 1;
}

__global__ void k_sampler_values(const __grid_constant__ DScene S, const int* pixel_xy, const int* sample_num, uint32_t n, uint32_t dim0, uint32_t ndims, float* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * ndims; i += gridDim.x * blockDim.x) {
    uint32_t s = i / ndims, k = i % ndims;
    int px = pixel_xy[2 * s], py = pixel_xy[2 * s + 1];
    long long idx = sampler_index(S, px, py, sample_num[s]);
    out[i] = sampler_dim(S, idx, dim0 + k, px, py);
  }
}

// ---- SpatialLightDistribution::compute_distribution for every voxel (spatial.rs:113-197) ---------------------------
template <int NC>
__global__ void k_voxel_contrib(const __grid_constant__ DScene S, float* func, uint32_t n_vox, size_t first, size_t count) {
  // (voxel, light) pairs first .. first + count of the n_vox x n_lights table (the whole table on the device; the CPU test
  // build gives each host thread a slice)
  const uint32_t nl = S.n_lights;
  const LightDistDev& D = S.ld;
  for (size_t id = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x; id < first + count; id += (size_t)gridDim.x * blockDim.x) {
    uint32_t vox = (uint32_t)(id / nl), j = (uint32_t)(id % nl);
    uint32_t pz = vox % D.voxels[2], py = (vox / D.voxels[2]) % D.voxels[1], px = vox / (D.voxels[2] * D.voxels[1]);
    const uint32_t pi[3] = {px, py, pz};
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      float t0 = (float)pi[a] / (float)D.voxels[a], t1 = (float)(pi[a] + 1) / (float)D.voxels[a];
      float a0 = lerpf(t0, D.wb_min[a], D.wb_max[a]), a1 = lerpf(t1, D.wb_min[a], D.wb_max[a]);
      lo[a] = a0 <= a1 ? a0 : a1;  // Bounds3f::new
      hi[a] = a0 >= a1 ? a0 : a1;
    }
    float contrib = 0.0f;
    for (uint32_t i = 0; i < 128; i++) {
      float tx = radical_inverse2(i), ty = radical_inverse_base(3u, i), tz = radical_inverse_base(5u, i);
      RefPt it;
      it.p = mk3(lerpf(tx, lo[0], hi[0]), lerpf(ty, lo[1], hi[1]), lerpf(tz, lo[2], hi[2]));
      it.perr = mk3(0.0f, 0.0f, 0.0f);
      it.n = mk3(0.0f, 0.0f, 0.0f);
      f2 u = mk2(radical_inverse_base(7u, i), radical_inverse_base(11u, i));
      LightSampleD<NC> ls;
      if (light_sample<NC, true>(S, (int)j, it, u, &ls))
        if (ls.pdf > 0.0f) contrib += sp_y(S, ls.li) / ls.pdf;
    }
    func[id] = contrib;
  }
}
// per voxel: floor at 0.1 % of the mean, then Distribution1D::new (distribution.rs:35-62)
__global__ void k_voxel_cdf(uint32_t nl, uint32_t n_vox, float* func, float* cdf, float* func_int) {
  for (uint32_t vox = blockIdx.x * blockDim.x + threadIdx.x; vox < n_vox; vox += gridDim.x * blockDim.x) {
    float* f = func + (size_t)vox * nl;
    float* c = cdf + (size_t)vox * (nl + 1);
    float sum = 0.0f;
    for (uint32_t j = 0; j < nl; j++) sum += f[j];
    float avg = sum / (float)(128u * nl);
    float mn = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (uint32_t j = 0; j < nl; j++) f[j] = fmaxf(mn, f[j]);
    c[0] = 0.0f;
    for (uint32_t j = 1; j < nl + 1; j++) c[j] = c[j - 1] + f[j - 1] / (float)nl;
    float fi = c[nl];
    if (fi == 0.0f) { for (uint32_t j = 1; j < nl + 1; j++) c[j] = (float)j / (float)nl; }
    else { for (uint32_t j = 1; j < nl + 1; j++) c[j] /= fi; }
    func_int[vox] = fi;
  }
}

}  // namespace ptd
