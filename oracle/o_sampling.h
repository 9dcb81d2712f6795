// ORACLE — TEST INFRASTRUCTURE ONLY (see o_math.h header).
// o_sampling.h — PCG32, low-discrepancy sequences, Halton/Sobol/Random samplers, warping
// functions, Distribution1D.
// Follows reference src/core/rng.rs, src/core/lowdiscrepancy/{radical_inverse,primes}.rs,
// src/core/lowdiscrepancy/sobol/sobol.rs, src/samplers/{halton,sobol,random}.rs,
// src/core/sampler/{sampler,global_sampler}.rs, src/core/sampling/{sampling,distribution}.rs.
#pragma once
#include "o_math.h"
#include <vector>
#include <cassert>

namespace orc {

// ---- core/rng.rs:3-66 ------------------------------------------------------------------
struct RNG {
  uint64_t state, inc;
  RNG() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
  explicit RNG(uint64_t seq) : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) { set_sequence(seq); }
  void set_sequence(uint64_t initseq) {
    state = 0;
    inc = (initseq << 1) | 1;
    uniform_uint32();
    state += 0x853c49e6748fea9bULL;
    uniform_uint32();
  }
  uint32_t uniform_uint32() {
    uint64_t oldstate = state;
    state = oldstate * 0x5851f42d4c957f2dULL + inc;
    uint32_t xorshifted = (uint32_t)(((oldstate >> 18) ^ oldstate) >> 27);
    uint32_t rot = (uint32_t)(oldstate >> 59);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
  }
  uint32_t uniform_uint32(uint32_t b) {   // uniform_uint32_threshold
    uint32_t threshold = (~b + 1u) % b;
    for (;;) {
      uint32_t r = uniform_uint32();
      if (r >= threshold) return r % b;
    }
  }
  float uniform_float() {
    float f = (float)uniform_uint32() * 2.3283064365386963e-10f;
    return fminf(ONE_MINUS_EPSILON, f);
  }
};

// ---- core/lowdiscrepancy/primes.rs: first 1024 primes and their prefix sums -----------------
struct Tables {
  std::vector<uint64_t> primes;      // PRIMES[1024]
  std::vector<uint64_t> prime_sums;  // PRIME_SUMS (prefix sums, [0] = 0)
  std::vector<uint16_t> perms;       // compute_radical_inverse_permutations with RNG::new()
  const uint32_t* sobol32 = nullptr; // SOBOL_MATRICES_32 [1024*52]   (data blob)
  const uint64_t* vdc = nullptr;     // VDC_SOBOL_MATRICES [25][52] ++ VDC_SOBOL_MATRICES_INV [26][52]
  Tables() {
    primes.reserve(1024);
    for (uint64_t n = 2; primes.size() < 1024; n++) {
      bool is = true;
      for (uint64_t p : primes) { if (p * p > n) break; if (n % p == 0) { is = false; break; } }
      if (is) primes.push_back(n);
    }
    prime_sums.resize(1024);
    uint64_t s = 0;
    for (int i = 0; i < 1024; i++) { prime_sums[i] = s; s += primes[i]; }
    // radical_inverse.rs:93-109 + sampling.rs:4-15 (shuffle_array, dim = 1)
    perms.resize(s);
    RNG rng;
    size_t off = 0;
    for (int i = 0; i < 1024; i++) {
      size_t len = primes[i];
      for (size_t j = 0; j < len; j++) perms[off + j] = (uint16_t)j;
      for (size_t j = 0; j < len; j++) {
        size_t other = j + rng.uniform_uint32((uint32_t)(len - j));
        std::swap(perms[off + j], perms[off + other]);
      }
      off += len;
    }
  }
};
inline Tables& tables() { static Tables t; return t; }

// ---- core/lowdiscrepancy/radical_inverse.rs -----------------------------------------------------
inline uint32_t reverse_bits32(uint32_t n) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
  n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
  n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
  n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
  return n;
}
inline uint64_t reverse_bits64(uint64_t n) {
  uint64_t n0 = reverse_bits32((uint32_t)n);
  uint64_t n1 = reverse_bits32((uint32_t)(n >> 32));
  return (n0 << 32) | n1;
}
inline Float radical_inverse_specialized(uint64_t base, uint64_t a) {   // :38-50
  Float inv_base = 1.0f / (Float)base;
  uint64_t reversed_digits = 0;
  Float inv_base_n = 1.0f;
  while (a != 0) {
    uint64_t next = a / base;
    uint64_t digit = a - next * base;
    reversed_digits = reversed_digits * base + digit;
    inv_base_n *= inv_base;
    a = next;
  }
  return rmin((Float)reversed_digits * inv_base_n, ONE_MINUS_EPSILON);
}
inline Float radical_inverse(uint32_t base_index, uint64_t a) {            // :52-58
  if (base_index == 0) return (Float)reverse_bits64(a) * 5.4210108624275222e-20f;
  return radical_inverse_specialized(tables().primes[base_index], a);
}
inline uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, size_t ndigits) {  // :60-69
  uint64_t index = 0;
  for (size_t i = 0; i < ndigits; i++) {
    uint64_t digit = inverse % base;
    inverse /= base;
    index = index * base + digit;
  }
  return index;
}
inline Float scrambled_radical_inverse(uint32_t base_index, uint64_t a, const uint16_t* perm) {  // :71-91
  uint64_t base = tables().primes[base_index];
  Float inv_base = 1.0f / (Float)base;
  uint64_t reverse_digits = 0;
  Float inv_base_n = 1.0f;
  while (a != 0) {
    uint64_t next = a / base;
    uint64_t digit = a - next * base;
    reverse_digits = reverse_digits * base + perm[digit];
    inv_base_n *= inv_base;
    a = next;
  }
  return rmin(inv_base_n * ((Float)reverse_digits + inv_base * (Float)perm[0] / (1.0f - inv_base)), ONE_MINUS_EPSILON);
}

// ---- core/lowdiscrepancy/sobol/sobol.rs -----------------------------------------------------------
inline uint64_t sobol_interval_to_index(uint32_t m, uint64_t frame, int32_t px, int32_t py) {  // :5-34
  if (m == 0) return 0;
  const uint64_t* vdc = tables().vdc;
  const uint64_t* vdc_inv = vdc + 25 * 52;
  uint32_t m2 = m << 1;
  uint64_t index = frame << (m2 & 63);
  uint64_t delta = 0;
  for (int c = 0; frame != 0; c++, frame >>= 1)
    if (frame & 1) delta ^= vdc[(m - 1) * 52 + c];
  uint64_t b = ((((uint64_t)(uint32_t)px) << (m & 63)) | (uint64_t)(int64_t)py) ^ delta;
  for (int c = 0; b != 0; c++, b >>= 1)
    if (b & 1) index ^= vdc_inv[(m - 1) * 52 + c];
  return index;
}
inline float sobol_sample_float(int64_t a, uint32_t dimension, uint32_t scramble) {       // :41-56
  const uint32_t* M = tables().sobol32;
  const size_t LEN = 1024 * 52;
  uint32_t v = scramble;
  size_t i = std::min((size_t)dimension * 52, LEN - 1);
  while (a != 0) {
    if (a & 1) v ^= M[i];
    a >>= 1;
    i += 1;
    i %= LEN;
  }
  float fv = (float)((double)v * 2.3283064365386963e-10);
  return fminf(fv, ONE_MINUS_EPSILON);
}

// ---- samplers ---------------------------------------------------------------------------------
struct SamplerConfig {
  int type;            // pt_sampler_type
  int spp;             // effective samples per pixel
  int sb[4];           // sample bounds x0,y0,x1,y1
  bool at_center;
  // halton (samplers/halton.rs:57-101)
  int base_scales[2], base_exponents[2], sample_stride, mult_inverse[2];
  // sobol (samplers/sobol.rs:18-31)
  uint32_t resolution, log2_resolution;
  // sample arrays the integrator requested (base_sampler.rs:32-44): ambientocclusion one 2D array, directlighting two per light
  uint32_t n_arrays_1d, n_arrays_2d;
  std::vector<uint32_t> array_sizes_2d;  // n of each requested 2D array, in request order (random sampler: its own values per array)
  // stratified (samplers/stratified.rs:15-30): xsamples = ysamples strata per axis, jitter, n_sampled_dimensions tables
  int strat_n = 0, strat_dims = 0;
  bool jitter = true;
};

inline uint64_t math_mod(int64_t a, int64_t b) {  // halton.rs:22-29
  int64_t result = a - (a / b) * b;
  return result < 0 ? (uint64_t)(result + b) : (uint64_t)result;
}
inline void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {  // halton.rs:31-38
  if (b == 0) { *x = 1; *y = 0; return; }
  int64_t d = (int64_t)(a / b);
  int64_t xp, yp;
  extended_gcd(b, a % b, &xp, &yp);
  *x = yp;
  *y = xp - (d * yp);
}
inline uint64_t multiplicative_inverse(int64_t a, int64_t n) {  // halton.rs:40-43
  int64_t x, y;
  extended_gcd((uint64_t)a, (uint64_t)n, &x, &y);
  return math_mod(x, n);
}
inline uint32_t round_up_pow2(uint32_t v) { v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1; }
inline uint32_t log2int(uint32_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }

inline SamplerConfig make_sampler_config(int type, int spp, const int sb[4], bool at_center) {
  SamplerConfig c{};
  c.type = type; c.spp = spp; c.at_center = at_center;
  for (int i = 0; i < 4; i++) c.sb[i] = sb[i];
  int res[2] = {sb[2] - sb[0], sb[3] - sb[1]};
  const int BASES[2] = {2, 3};
  for (int i = 0; i < 2; i++) {
    int scale = 1, exp = 0;
    while (scale < std::min(res[i], 128)) { scale *= BASES[i]; exp += 1; }
    c.base_scales[i] = scale; c.base_exponents[i] = exp;
  }
  c.sample_stride = c.base_scales[0] * c.base_scales[1];
  c.mult_inverse[0] = (int)multiplicative_inverse(c.base_scales[1], c.base_scales[0]);
  c.mult_inverse[1] = (int)multiplicative_inverse(c.base_scales[0], c.base_scales[1]);
  if (type == 1) {  // sobol.rs:19-30
    c.spp = (int)round_up_pow2((uint32_t)spp);
    c.resolution = round_up_pow2((uint32_t)std::max(res[0], res[1]));
    c.log2_resolution = log2int(c.resolution);
  }
  return c;
}

// One sampler instance = the state the reference keeps per tile clone.  The Halton pixel-offset
// cache is implemented with its INTENDED semantics (a pure function of the pixel; SURVEY §5 notes
// the reference's shared-cache race).
struct Sampler {
  const SamplerConfig* cfg;
  int px, py;           // current_pixel
  int64_t sample_num;   // current_pixel_sample_index
  int64_t interval_index;
  uint32_t dimension;
  RNG rng;              // random sampler: per-tile stream
  int64_t halton_offset;
  uint32_t array2d_offset = 0;  // get_2d_array calls so far in this sample (base_sampler.rs:59-70)
  std::vector<std::vector<Vec2>> random_arrays;  // RandomSampler: sample_array2d, refilled by every start_pixel (random.rs:46-60)
  // PixelSampler tables (core/sampler/pixel_sampler.rs:10-16): [dimension][sample of the pixel]
  std::vector<std::vector<Float>> samples1d;
  std::vector<std::vector<Vec2>> samples2d;
  uint32_t current1d = 0, current2d = 0;
  template <class T> void shuffle(T* v, size_t count) {  // shuffle_array with one dimension (sampling.rs:4-15)
    for (size_t i = 0; i < count; i++) { size_t other = i + rng.uniform_uint32((uint32_t)(count - i)); std::swap(v[i], v[other]); }
  }
  // ---- ZeroTwoSequenceSampler (samplers/zerotwosequence.rs:30-63; core/lowdiscrepancy/van_der_corput.rs:59-145) ----
  static const uint32_t* sobol_matrix(int which) {  // CVAN_DER_CORPUT = CSOBOL[0] = bit reversal; CSOBOL[1][i] = CSOBOL[1][i-1] ^ (CSOBOL[1][i-1] >> 1)
    static uint32_t m[2][32];
    static bool ready = false;
    if (!ready) {
      for (int i = 0; i < 32; i++) m[0][i] = 1u << (31 - i);
      m[1][0] = 1u << 31;
      for (int i = 1; i < 32; i++) m[1][i] = m[1][i - 1] ^ (m[1][i - 1] >> 1);
      ready = true;
    }
    return m[which];
  }
  static Float u32_to_unit(uint32_t v) { return std::min((Float)v * 2.3283064365386963e-10f, ONE_MINUS_EPSILON); }
  template <class T> void shuffle_blocks(T* v, size_t count, size_t dim) {  // shuffle_array (sampling.rs:4-15)
    for (size_t i = 0; i < count; i++) {
      size_t other = i + rng.uniform_uint32((uint32_t)(count - i));
      for (size_t j = 0; j < dim; j++) std::swap(v[dim * i + j], v[dim * other + j]);
    }
  }
  void van_der_corput(size_t n_per, size_t n_pixel, Float* out) {  // van_der_corput.rs:94-115
    uint32_t v = rng.uniform_uint32();
    const uint32_t* c = sobol_matrix(0);
    for (size_t i = 0; i < n_per * n_pixel; i++) { out[i] = u32_to_unit(v); v ^= c[__builtin_ctz((uint32_t)(i + 1))]; }
    for (size_t i = 0; i < n_pixel; i++) shuffle_blocks(out + i * n_per, n_per, 1);
    shuffle_blocks(out, n_pixel, n_per);
  }
  void sobol_2d(size_t n_per, size_t n_pixel, Vec2* out) {  // van_der_corput.rs:118-145
    uint32_t v0 = rng.uniform_uint32();
    uint32_t v1 = rng.uniform_uint32();
    const uint32_t *c0 = sobol_matrix(0), *c1 = sobol_matrix(1);
    for (size_t i = 0; i < n_per * n_pixel; i++) {
      out[i] = Vec2(u32_to_unit(v0), u32_to_unit(v1));
      int z = __builtin_ctz((uint32_t)(i + 1));
      v0 ^= c0[z]; v1 ^= c1[z];
    }
    for (size_t i = 0; i < n_pixel; i++) shuffle_blocks(out + i * n_per, n_per, 1);
    shuffle_blocks(out, n_pixel, n_per);
  }
  void zerotwo_start_pixel() {
    const size_t spp = (size_t)cfg->spp;
    samples1d.assign((size_t)cfg->strat_dims, std::vector<Float>(spp));
    samples2d.assign((size_t)cfg->strat_dims, std::vector<Vec2>(spp));
    for (auto& t : samples1d) van_der_corput(1, spp, t.data());
    for (auto& t : samples2d) sobol_2d(1, spp, t.data());
    random_arrays.resize(cfg->array_sizes_2d.size());
    for (size_t a = 0; a < random_arrays.size(); a++) {
      random_arrays[a].assign((size_t)cfg->array_sizes_2d[a] * spp, Vec2(0, 0));
      sobol_2d(cfg->array_sizes_2d[a], spp, random_arrays[a].data());
    }
  }
  void stratified_start_pixel() {  // stratified.rs:33-88
    const size_t spp = (size_t)cfg->spp;
    samples1d.assign((size_t)cfg->strat_dims, std::vector<Float>(spp));
    samples2d.assign((size_t)cfg->strat_dims, std::vector<Vec2>(spp));
    for (auto& t : samples1d) {  // stratified_sample_1d (sampling.rs:17-23) + shuffle
      Float inv = 1.0f / (Float)spp;
      for (size_t i = 0; i < spp; i++) { Float delta = cfg->jitter ? rng.uniform_float() : 0.5f; t[i] = std::min(((Float)i + delta) * inv, ONE_MINUS_EPSILON); }
      shuffle(t.data(), spp);
    }
    for (auto& t : samples2d) {  // stratified_sample_2d (sampling.rs:25-44; xsamples == ysamples) + shuffle
      const int n = cfg->strat_n;
      Float dx = 1.0f / (Float)n, dy = 1.0f / (Float)n;
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
          Float jx = cfg->jitter ? rng.uniform_float() : 0.5f;
          Float jy = cfg->jitter ? rng.uniform_float() : 0.5f;
          t[(size_t)(y * n + x)] = Vec2(std::min(((Float)x + jx) * dx, ONE_MINUS_EPSILON), std::min(((Float)y + jy) * dy, ONE_MINUS_EPSILON));
        }
      shuffle(t.data(), spp);
    }
    random_arrays.resize(cfg->array_sizes_2d.size());  // the 2D arrays: one Latin hypercube per sample (stratified.rs:72-82, sampling.rs:46-81)
    for (size_t a = 0; a < random_arrays.size(); a++) {
      const size_t n = cfg->array_sizes_2d[a];
      random_arrays[a].assign(n * spp, Vec2(0, 0));
      for (size_t j = 0; j < spp; j++) {
        std::vector<Float> v(2 * n);
        Float inv = 1.0f / (Float)n;
        for (size_t i = 0; i < n; i++)
          for (size_t k = 0; k < 2; k++) { Float sj = ((Float)i + rng.uniform_float()) * inv; v[2 * i + k] = std::min(sj, ONE_MINUS_EPSILON); }
        for (size_t k = 0; k < 2; k++)
          for (size_t i = 0; i < n; i++) { size_t other = i + rng.uniform_uint32((uint32_t)(n - i)); std::swap(v[2 * i + k], v[2 * other + k]); }
        for (size_t i = 0; i < n; i++) random_arrays[a][j * n + i] = Vec2(v[2 * i], v[2 * i + 1]);
      }
    }
  }

  Sampler(const SamplerConfig* c, uint32_t tile_seed) : cfg(c), px(0), py(0), sample_num(0), interval_index(0), dimension(0), rng((uint64_t)tile_seed), halton_offset(0) {}

  int64_t index_for_sample(int64_t n) const {
    if (cfg->type == 0) return halton_offset + n * (int64_t)cfg->sample_stride;     // halton.rs:146
    if (cfg->type == 1) return (int64_t)sobol_interval_to_index(cfg->log2_resolution, (uint64_t)n, px - cfg->sb[0], py - cfg->sb[1]);  // sobol.rs:169-172
    return 0;
  }
  void start_pixel(int x, int y) {
    px = x; py = y; sample_num = 0; dimension = 0; array2d_offset = 0;
    current1d = current2d = 0;
    if (cfg->type == 3) { stratified_start_pixel(); return; }
    if (cfg->type == 4) { zerotwo_start_pixel(); return; }
    if (cfg->type == 2 && !cfg->array_sizes_2d.empty()) {  // random.rs:46-60: every element of every array, in order, before the first sample
      random_arrays.resize(cfg->array_sizes_2d.size());
      for (size_t a = 0; a < random_arrays.size(); a++) {
        random_arrays[a].resize((size_t)cfg->array_sizes_2d[a] * (size_t)cfg->spp);
        for (Vec2& v : random_arrays[a]) { Float x0 = rng.uniform_float(); Float y0 = rng.uniform_float(); v = Vec2(x0, y0); }
      }
    }
    if (cfg->type == 0) {   // halton.rs:115-147
      halton_offset = 0;
      int64_t stride = cfg->sample_stride;
      if (stride > 1) {
        uint64_t pm[2] = {math_mod(px, 128), math_mod(py, 128)};
        const uint64_t BASES[2] = {2, 3};
        for (int i = 0; i < 2; i++) {
          uint64_t dim_offset = inverse_radical_inverse(BASES[i], pm[i], (size_t)cfg->base_exponents[i]);
          int64_t addition = (int64_t)(dim_offset * (uint64_t)(int64_t)((cfg->sample_stride / cfg->base_scales[i]) * cfg->mult_inverse[i]));
          halton_offset = halton_offset + addition;
        }
        halton_offset = halton_offset % stride;
      }
    }
    interval_index = index_for_sample(0);
  }
  bool set_sample_number(int64_t n) {
    sample_num = n; dimension = 0; array2d_offset = 0;
    current1d = current2d = 0;
    interval_index = index_for_sample(n);
    return n < cfg->spp;
  }
  bool start_next_sample() { return set_sample_number(sample_num + 1); }

  Float sample_dimension(int64_t index, uint32_t dim) const {
    if (cfg->type == 0) {   // halton.rs:149-162
      if (cfg->at_center && (dim == 0 || dim == 1)) return 0.5f;
      if (dim == 0) return radical_inverse(0, (uint64_t)(index >> cfg->base_exponents[0]));
      if (dim == 1) return radical_inverse(1, (uint64_t)(index / (int64_t)cfg->base_scales[1]));
      assert(dim < 1024);
      return scrambled_radical_inverse(dim, (uint64_t)index, tables().perms.data() + tables().prime_sums[dim]);
    }
    // sobol.rs:174-185
    Float s = sobol_sample_float(index, dim, 0);
    if (dim == 0 || dim == 1) {
      int cp = dim == 0 ? px : py;
      s = s * (Float)cfg->resolution + (Float)cfg->sb[dim];
      s = rclamp(s - (Float)cp, 0.0f, ONE_MINUS_EPSILON);
      s = s - truncf(s);
    }
    return s;
  }
  // GlobalSampler: dimensions [array_start_dim = 5, array_end_dim) belong to the sample arrays (global_sampler.rs:22,
  // halton.rs:176-178 / sobol.rs:43-45) and are skipped by get_1d / get_2d (halton.rs:227-231, 241-245)
  uint32_t array_end_dim() const { return 5u + cfg->n_arrays_1d + 2u * cfg->n_arrays_2d; }
  // element i of get_2d_array(n) for the current sample: start_pixel fills array[j] from the index of sample NUMBER j,
  // j < n * spp, and get_2d_array hands out [n * s, n * (s + 1)) (halton.rs:193-207, base_sampler.rs:59-70).  Every 2D array
  // reads the same two dimensions (`dim` is not advanced per array in the reference).
  // get_2d_array: None once every requested array has been handed out (base_sampler.rs:63-65)
  bool next_2d_array() {
    if (array2d_offset == cfg->n_arrays_2d) return false;
    array2d_offset++;
    return true;
  }
  // element i of the array handed out by the k-th get_2d_array call of this sample (k = 0, 1, ..)
  Vec2 array_2d_elem_of(uint32_t k, uint32_t n, uint32_t i) const {
    if (cfg->type >= 2) return random_arrays[k][(size_t)n * (size_t)sample_num + i];
    return array_2d_elem(n, i);
  }
  Vec2 array_2d_elem(uint32_t n, uint32_t i) const {
    int64_t index = index_for_sample((int64_t)n * sample_num + (int64_t)i);
    uint32_t dim = 5u + cfg->n_arrays_1d;
    return Vec2(sample_dimension(index, dim), sample_dimension(index, dim + 1));
  }
  Float get_1d() {
    if (cfg->type >= 3) {  // BasePixelSampler::get_1d (pixel_sampler.rs:53-64)
      if (current1d < samples1d.size()) return samples1d[current1d++][(size_t)sample_num];
      return rng.uniform_float();
    }
    if (cfg->type == 2) return rng.uniform_float();
    if (dimension >= 5u && dimension < array_end_dim()) dimension = array_end_dim();
    Float x = sample_dimension(interval_index, dimension);
    dimension += 1;
    return x;
  }
  Vec2 get_2d() {
    if (cfg->type >= 3 && current2d < samples2d.size()) return samples2d[current2d++][(size_t)sample_num];  // pixel_sampler.rs:66-77
    if (cfg->type >= 2) { Float a = rng.uniform_float(); Float b = rng.uniform_float(); return Vec2(a, b); }
    if (dimension + 1u >= 5u && dimension < array_end_dim()) dimension = array_end_dim();
    Float x = sample_dimension(interval_index, dimension);
    Float y = sample_dimension(interval_index, dimension + 1);
    dimension += 2;
    return Vec2(x, y);
  }
};

// ---- core/sampling/sampling.rs ---------------------------------------------------------------
inline Vec2 uniform_sample_triangle(Vec2 u) { Float su0 = sqrtf(u.x); return Vec2(1.0f - su0, u.y * su0); }  // :101-104
inline Vec2 concentric_sample_disk(Vec2 u) {                                                                    // :106-122
  Vec2 uo = u * 2.0f - Vec2(1.0f, 1.0f);
  if (uo.x == 0.0f && uo.y == 0.0f) return Vec2(0, 0);
  if (fabsf(uo.x) > fabsf(uo.y)) {
    Float r = uo.x;
    Float theta = PI_OVER_4 * (uo.y / uo.x);
    return Vec2(r * cosf(theta), r * sinf(theta));
  } else {
    Float r = uo.y;
    Float theta = PI_OVER_2 - PI_OVER_4 * (uo.x / uo.y);
    return Vec2(r * cosf(theta), r * sinf(theta));
  }
}
inline Vec3 uniform_sample_hemisphere(Vec2 u) {                                                                 // :85-90
  Float z = u.x;
  Float r = sqrtf(rmax(0.0f, 1.0f - z * z));
  Float phi = 2.0f * PI * u.y;
  return Vec3(r * cosf(phi), r * sinf(phi), z);
}
inline Vec3 cosine_sample_hemisphere(Vec2 u) {                                                                  // :141-145
  Vec2 d = concentric_sample_disk(u);
  Float z = sqrtf(rmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
  return Vec3(d.x, d.y, z);
}
inline Float power_heuristic(int nf, Float f_pdf, int ng, Float g_pdf) {                                        // :168-172
  Float f = (Float)nf * f_pdf, g = (Float)ng * g_pdf;
  return (f * f) / (f * f + g * g);
}

// ---- core/sampling/distribution.rs:12-106 --------------------------------------------------------
struct Distribution1D {
  std::vector<Float> func, cdf;
  Float func_int, inv_count;
  Distribution1D() : func_int(0), inv_count(0) {}
  explicit Distribution1D(const std::vector<Float>& f) : func(f) {
    size_t n = f.size();
    cdf.assign(n + 1, 0.0f);
    for (size_t i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] / (Float)n;
    func_int = cdf[n];
    if (func_int == 0.0f) for (size_t i = 1; i < n + 1; i++) cdf[i] = (Float)i / (Float)n;
    else for (size_t i = 1; i < n + 1; i++) cdf[i] /= func_int;
    inv_count = 1.0f / (Float)n;
  }
  size_t find_interval(Float u) const {   // :13-32
    size_t first = 0, len = cdf.size();
    while (len > 0) {
      size_t half = len >> 1, middle = first + half;
      if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
      else len = half;
    }
    size_t idx = first == 0 ? 0 : first - 1;
    if (idx > cdf.size() - 2) return cdf.size() - 2;
    return idx;
  }
  // :86-101 ; returns offset, pdf
  size_t sample_discrete(Float u, Float* pdf) const {
    size_t offset = find_interval(u);
    *pdf = func_int > 0.0f ? func[offset] * inv_count / func_int : 0.0f;
    return offset;
  }
  // :64-83 ; returns value, pdf, offset
  Float sample_continuous(Float u, Float* pdf, size_t* off) const {
    size_t offset = find_interval(u);
    Float cdf0 = cdf[offset], cdf1 = cdf[offset + 1];
    Float du = u - cdf0;
    Float span = cdf1 - cdf0;
    if (span > 0.0f) du /= span;
    *pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
    *off = offset;
    return ((Float)offset + du) * inv_count;
  }
};

}  // namespace orc
