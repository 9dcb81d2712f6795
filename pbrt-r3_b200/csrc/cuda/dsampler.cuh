// dsampler.cuh — Halton and Sobol sample streams on the device, reproducing the reference's values bit for bit.
// Replaces reference src/samplers/halton.rs:115-162 (pixel offset, sample_dimension), src/samplers/sobol.rs:168-185,
// src/core/lowdiscrepancy/radical_inverse.rs:38-91, src/core/lowdiscrepancy/sobol/sobol.rs:5-56,
// src/core/sampler/sampler.rs:26-36 (camera sample = dims 0-1 film, 2-3 lens, 4 time).
// A sample is a pure function of (pixel, sample number, dimension): the wavefront keeps only an i64 index
// and a dimension counter per path.
#pragma once
#include "dscene.cuh"

namespace ptd {

PT_HD uint32_t rev32(uint32_t n) {
#if defined(__CUDA_ARCH__)
  return __brev(n);
#else
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
  n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
  return n;
#endif
}
PT_HD uint64_t rev64(uint64_t n) { return ((uint64_t)rev32((uint32_t)n) << 32) | (uint64_t)rev32((uint32_t)(n >> 32)); }

// radical_inverse(base_index, a): base 2 by bit reversal, otherwise the digit loop (radical_inverse.rs:38-58)
PT_HD float radical_inverse_base(uint32_t base, uint64_t a) {
  float inv_base = div_rn(1.0f, (float)base);
  uint64_t rev = 0;
  float inv_base_n = 1.0f;
  if (a < 0x100000000ull) {  // 32-bit fast path: same digits, same float sequence
    uint32_t a32 = (uint32_t)a;
    while (a32 != 0) {
      uint32_t next = a32 / base;
      uint32_t digit = a32 - next * base;
      rev = rev * base + digit;
      inv_base_n *= inv_base;
      a32 = next;
    }
  } else {
    while (a != 0) {
      uint64_t next = a / base;
      uint64_t digit = a - next * base;
      rev = rev * base + digit;
      inv_base_n *= inv_base;
      a = next;
    }
  }
  return fminf((float)rev * inv_base_n, PT_ONE_MINUS_EPS);
}
PT_HD float radical_inverse2(uint64_t a) { return (float)rev64(a) * 5.4210108624275222e-20f; }

// floor(a / b) for a < 2^32 as the high half of a * ceil(2^64 / b): exact because a * b < 2^64
PT_HD uint32_t div_magic(uint32_t a, uint64_t magic) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__umul64hi((unsigned long long)a, (unsigned long long)magic);
#else
  return (uint32_t)(((unsigned __int128)a * (unsigned __int128)magic) >> 64);
#endif
}

// high half of a 32 x 32 bit product
PT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

PT_HD float scrambled_radical_inverse(uint32_t base, uint64_t magic, uint64_t a, const uint16_t* perm) {  // radical_inverse.rs:71-91
  float inv_base = div_rn(1.0f, (float)base);
  float inv_base_n = 1.0f;
  float frev;
  if (a * (uint64_t)base < 0x100000000ull) {
    // Everything fits 32 bits (the usual case: index < 2^21 at 64 spp): the reversed number has as many digits as a, so
    // rev < a * base < 2^32; and with m = ceil(2^32 / base) = (2^32 + e) / base, e <= base, floor(a * m / 2^32) = floor(a / base)
    // exactly because a * e <= a * base < 2^32.  Same digits, same float sequence as the reference's u64 loop.
    const uint32_t m32 = (uint32_t)(magic >> 32) + 1u;  // base is not a power of two (dimensions >= 2: primes >= 5)
    uint32_t a32 = (uint32_t)a, rev = 0;
    while (a32 != 0) {
      uint32_t next = mulhi32(a32, m32);
      uint32_t digit = a32 - next * base;
      rev = rev * base + (uint32_t)ldg(perm + digit);
      inv_base_n *= inv_base;
      a32 = next;
    }
    frev = (float)rev;
  } else {
    uint64_t rev = 0;
    if (a < 0x100000000ull) {
      uint32_t a32 = (uint32_t)a;
      while (a32 != 0) {
        uint32_t next = div_magic(a32, magic);
        uint32_t digit = a32 - next * base;
        rev = rev * base + (uint64_t)ldg(perm + digit);
        inv_base_n *= inv_base;
        a32 = next;
      }
    } else {
      while (a != 0) {
        uint64_t next = a / base;
        uint64_t digit = a - next * base;
        rev = rev * base + (uint64_t)ldg(perm + digit);
        inv_base_n *= inv_base;
        a = next;
      }
    }
    frev = (float)rev;
  }
  return fminf(inv_base_n * (frev + div_rn(inv_base * (float)ldg(perm), 1.0f - inv_base)), PT_ONE_MINUS_EPS);
}

PT_HD uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, int ndigits) {  // radical_inverse.rs:60-69
  uint64_t index = 0;
  for (int i = 0; i < ndigits; i++) {
    uint64_t digit = inverse % base;
    inverse /= base;
    index = index * base + digit;
  }
  return index;
}
PT_HD uint64_t math_mod(int64_t a, int64_t b) {  // halton.rs:22-29
  int64_t r = a - (a / b) * b;
  return r < 0 ? (uint64_t)(r + b) : (uint64_t)r;
}

PT_HD uint64_t sobol_interval_to_index(const uint64_t* vdc, uint32_t m, uint64_t frame, int32_t px, int32_t py) {  // sobol.rs:5-34
  if (m == 0) return 0;
  const uint64_t* vdc_inv = vdc + 25 * 52;
  uint32_t m2 = m << 1;
  uint64_t index = frame << (m2 & 63);
  uint64_t delta = 0;
  for (int c = 0; frame != 0; c++, frame >>= 1)
    if (frame & 1) delta ^= ldg(vdc + (m - 1) * 52 + c);
  uint64_t b = ((((uint64_t)(uint32_t)px) << (m & 63)) | (uint64_t)(int64_t)py) ^ delta;
  for (int c = 0; b != 0; c++, b >>= 1)
    if (b & 1) index ^= ldg(vdc_inv + (m - 1) * 52 + c);
  return index;
}
PT_HD float sobol_sample_float(const uint32_t* M, int64_t a, uint32_t dimension, uint32_t scramble) {  // sobol.rs:41-56
  const uint32_t LEN = 1024u * 52u;
  uint32_t v = scramble;
  uint32_t i = dimension * 52u;
  if (i > LEN - 1) i = LEN - 1;
  while (a != 0) {
    if (a & 1) v ^= ldg(M + i);
    a >>= 1;
    i += 1;
    if (i >= LEN) i -= LEN;
  }
  float fv = (float)((double)v * 2.3283064365386963e-10);
  return fminf(fv, PT_ONE_MINUS_EPS);
}

// index of sample `n` of pixel (px, py)  (halton.rs:115-147 with the INTENDED per-pixel offset; sobol.rs:168-172)
PT_HD int64_t sampler_index(const DScene& S, int px, int py, int64_t n) {
  const SamplerDev& c = S.sampler;
  if (c.type == PT_SAMPLER_HALTON) {
    int64_t offset = 0;
    int64_t stride = c.sample_stride;
    if (stride > 1) {
      uint64_t pm[2] = {math_mod(px, 128), math_mod(py, 128)};
      for (int i = 0; i < 2; i++) {
        uint64_t dim_offset = inverse_radical_inverse(i == 0 ? 2 : 3, pm[i], c.base_exponents[i]);
        int64_t addition = (int64_t)(dim_offset * (uint64_t)(int64_t)((c.sample_stride / c.base_scales[i]) * c.mult_inverse[i]));
        offset = offset + addition;
      }
      offset = offset % stride;
    }
    return offset + n * stride;
  }
  return (int64_t)sobol_interval_to_index(S.vdc, c.log2_resolution, (uint64_t)n, px - c.sb[0], py - c.sb[1]);
}

// Sampler "random" (samplers/random.rs:21-30, core/rng.rs:34-66): every draw advances the PCG32 stream of the path's
// render tile.  The reference runs the pixels and samples of a tile one after the other on one stream, so a wave holds at
// most one path per tile and the state lives in memory, not in the path.
PT_HD float random_draw(const DScene& S, uint32_t tile_slot, uint64_t tile_seq) {
  uint64_t old = S.tile_rng[tile_slot];
  S.tile_rng[tile_slot] = old * 0x5851f42d4c957f2dULL + ((tile_seq << 1) | 1ull);
  uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  uint32_t r = (xs >> rot) | (xs << ((~rot + 1u) & 31u));
  return fminf(PT_ONE_MINUS_EPS, (float)r * 2.3283064365386963e-10f);
}
PT_HD uint32_t random_u32(const DScene& S, uint32_t tile_slot, uint64_t tile_seq) {
  uint64_t old = S.tile_rng[tile_slot];
  S.tile_rng[tile_slot] = old * 0x5851f42d4c957f2dULL + ((tile_seq << 1) | 1ull);
  uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
}
PT_HD uint32_t random_bounded(const DScene& S, uint32_t tile_slot, uint64_t tile_seq, uint32_t b) {  // RNG::uniform_uint32_threshold (rng.rs:46-54)
  const uint32_t threshold = (~b + 1u) % b;
  for (;;) {
    uint32_t r = random_u32(S, tile_slot, tile_seq);
    if (r >= threshold) return r % b;
  }
}
// StratifiedSampler::start_pixel (stratified.rs:33-88) for the tile's next pixel: the 1D tables (stratified + shuffled), the
// 2D tables, then one Latin hypercube per sample for every requested 2D array (n_arr arrays of arr_n[a] elements per sample;
// arr_n = null: one array of arr_n0).  Layout: [nd][spp] | [nd][spp][2] | arrays in request order, [spp][n][2] each.
PT_HD void stratified_start_pixel(const DScene& S, uint32_t tile_slot, uint64_t seq, uint32_t n_arr, const uint32_t* arr_prefix, uint32_t arr_n0) {
  const SamplerDev& c = S.sampler;
  const uint32_t spp = (uint32_t)c.spp, nd = (uint32_t)c.nd;
  float* t = S.tile_arrays + (size_t)tile_slot * S.tile_arr_stride;
  const float inv = 1.0f / (float)spp;
  for (uint32_t d = 0; d < nd; d++, t += spp) {
    for (uint32_t i = 0; i < spp; i++) { float delta = c.jitter ? random_draw(S, tile_slot, seq) : 0.5f; t[i] = fminf(((float)i + delta) * inv, PT_ONE_MINUS_EPS); }
    for (uint32_t i = 0; i < spp; i++) { uint32_t o = i + random_bounded(S, tile_slot, seq, spp - i); float a = t[i]; t[i] = t[o]; t[o] = a; }
  }
  const uint32_t n = (uint32_t)c.strata;
  const float dx = 1.0f / (float)n;
  for (uint32_t d = 0; d < nd; d++, t += 2u * spp) {
    for (uint32_t y = 0; y < n; y++)
      for (uint32_t x = 0; x < n; x++) {
        float jx = c.jitter ? random_draw(S, tile_slot, seq) : 0.5f;
        float jy = c.jitter ? random_draw(S, tile_slot, seq) : 0.5f;
        t[2u * (y * n + x)] = fminf(((float)x + jx) * dx, PT_ONE_MINUS_EPS);
        t[2u * (y * n + x) + 1u] = fminf(((float)y + jy) * dx, PT_ONE_MINUS_EPS);
      }
    for (uint32_t i = 0; i < spp; i++) {
      uint32_t o = i + random_bounded(S, tile_slot, seq, spp - i);
      float a = t[2u * i], b = t[2u * i + 1u];
      t[2u * i] = t[2u * o]; t[2u * i + 1u] = t[2u * o + 1u];
      t[2u * o] = a; t[2u * o + 1u] = b;
    }
  }
  for (uint32_t a = 0; a < n_arr; a++) {  // latin_hypercube_2d per sample (sampling.rs:46-81)
    // arr_prefix (directlighting): prefix sums of the lights' n_samples; arrays come in pairs (u_light, u_scattering) per light
    const uint32_t nl = arr_prefix ? (n_arr / 2u) / (uint32_t)S.integrator.max_depth : 0u;
    const uint32_t j = arr_prefix ? (a / 2u) % nl : 0u;
    const uint32_t na = arr_prefix ? arr_prefix[j + 1u] - arr_prefix[j] : arr_n0;
    const float inva = 1.0f / (float)na;
    for (uint32_t s = 0; s < spp; s++, t += 2u * na) {
      for (uint32_t i = 0; i < na; i++)
        for (uint32_t k = 0; k < 2u; k++) t[2u * i + k] = fminf(((float)i + random_draw(S, tile_slot, seq)) * inva, PT_ONE_MINUS_EPS);
      for (uint32_t k = 0; k < 2u; k++)
        for (uint32_t i = 0; i < na; i++) {
          uint32_t o = i + random_bounded(S, tile_slot, seq, na - i);
          float v = t[2u * i + k]; t[2u * i + k] = t[2u * o + k]; t[2u * o + k] = v;
        }
    }
  }
}
// ZeroTwoSequenceSampler::start_pixel (zerotwosequence.rs:30-63; van_der_corput / sobol_2d, lowdiscrepancy/van_der_corput.rs:
// 59-145): every table is a Gray-code walk of a randomly scrambled generator matrix, then shuffled within each pixel sample's
// block and across the blocks.  Same layout and array bookkeeping as stratified_start_pixel.
PT_HD void zerotwo_blocks_shuffle(const DScene& S, uint32_t tile_slot, uint64_t seq, float* t, uint32_t n, uint32_t spp, uint32_t width) {
  // shuffle_array(&samples[i * n..], n, 1) for every pixel sample i, then shuffle_array(samples, spp, n) (sampling.rs:4-15);
  // width = floats per element
  for (uint32_t i = 0; i < spp; i++)
    for (uint32_t j = 0; j < n; j++) {
      uint32_t o = j + random_bounded(S, tile_slot, seq, n - j);
      for (uint32_t w = 0; w < width; w++) { float* a = t + width * (i * n + j) + w; float* b = t + width * (i * n + o) + w; float v = *a; *a = *b; *b = v; }
    }
  for (uint32_t i = 0; i < spp; i++) {
    uint32_t o = i + random_bounded(S, tile_slot, seq, spp - i);
    for (uint32_t w = 0; w < width * n; w++) { float* a = t + width * n * i + w; float* b = t + width * n * o + w; float v = *a; *a = *b; *b = v; }
  }
}
PT_HD void zerotwo_start_pixel(const DScene& S, uint32_t tile_slot, uint64_t seq, uint32_t n_arr, const uint32_t* arr_prefix, uint32_t arr_n0) {
  const SamplerDev& c = S.sampler;
  const uint32_t spp = (uint32_t)c.spp, nd = (uint32_t)c.nd;
  float* t = S.tile_arrays + (size_t)tile_slot * S.tile_arr_stride;
  uint32_t c1[32];  // second Sobol' generator matrix: c1[0] = 2^31, c1[i] = c1[i-1] ^ (c1[i-1] >> 1); the first is 2^31 >> i
  c1[0] = 0x80000000u;
  for (int i = 1; i < 32; i++) c1[i] = c1[i - 1] ^ (c1[i - 1] >> 1);
  for (uint32_t d = 0; d < nd; d++, t += spp) {
    uint32_t v = random_u32(S, tile_slot, seq);
    for (uint32_t i = 0; i < spp; i++) { t[i] = fminf((float)v * 2.3283064365386963e-10f, PT_ONE_MINUS_EPS); v ^= 0x80000000u >> (ffs_(i + 1u) - 1); }
    zerotwo_blocks_shuffle(S, tile_slot, seq, t, 1u, spp, 1u);
  }
  for (uint32_t d = 0; d < nd; d++, t += 2u * spp) {
    uint32_t v0 = random_u32(S, tile_slot, seq), v1 = random_u32(S, tile_slot, seq);
    for (uint32_t i = 0; i < spp; i++) {
      t[2u * i] = fminf((float)v0 * 2.3283064365386963e-10f, PT_ONE_MINUS_EPS);
      t[2u * i + 1u] = fminf((float)v1 * 2.3283064365386963e-10f, PT_ONE_MINUS_EPS);
      const int z = ffs_(i + 1u) - 1;
      v0 ^= 0x80000000u >> z; v1 ^= c1[z];
    }
    zerotwo_blocks_shuffle(S, tile_slot, seq, t, 1u, spp, 2u);
  }
  for (uint32_t a = 0; a < n_arr; a++) {
    const uint32_t nl = arr_prefix ? (n_arr / 2u) / (uint32_t)S.integrator.max_depth : 0u;
    const uint32_t j = arr_prefix ? (a / 2u) % nl : 0u;
    const uint32_t na = arr_prefix ? arr_prefix[j + 1u] - arr_prefix[j] : arr_n0;
    uint32_t v0 = random_u32(S, tile_slot, seq), v1 = random_u32(S, tile_slot, seq);
    for (uint32_t i = 0; i < na * spp; i++) {
      t[2u * i] = fminf((float)v0 * 2.3283064365386963e-10f, PT_ONE_MINUS_EPS);
      t[2u * i + 1u] = fminf((float)v1 * 2.3283064365386963e-10f, PT_ONE_MINUS_EPS);
      const int z = ffs_(i + 1u) - 1;
      v0 ^= 0x80000000u >> z; v1 ^= c1[z];
    }
    zerotwo_blocks_shuffle(S, tile_slot, seq, t, na, spp, 2u);
    t += 2u * na * spp;
  }
}
// BasePixelSampler::get_1d / get_2d (pixel_sampler.rs:53-77): dim = requests so far (1D low half, 2D high half) + which value
PT_HD float stratified_dim(const DScene& S, int64_t index, uint32_t dim) {
  const uint32_t tile = (uint32_t)((uint64_t)index >> 32), nd = (uint32_t)S.sampler.nd, spp = (uint32_t)S.sampler.spp;
  const float* t = S.tile_arrays + (size_t)tile * S.tile_arr_stride;
  if (dim & 0xc000u) {
    const uint32_t d2 = dim >> 16;
    if (d2 < nd) return t[nd * spp + 2u * (d2 * spp + S.tile_sample[tile]) + ((dim & 0x8000u) ? 1u : 0u)];
  } else {
    const uint32_t d1 = dim & 0x3fffu;
    if (d1 < nd) return t[d1 * spp + S.tile_sample[tile]];
  }
  return random_draw(S, tile, (uint64_t)(uint32_t)index);
}
PT_HDN float sampler_dim(const DScene& S, int64_t index, uint32_t dim, int px, int py) {
  const SamplerDev& c = S.sampler;
  if (c.type == PT_SAMPLER_RANDOM) return random_draw(S, (uint32_t)((uint64_t)index >> 32), (uint64_t)(uint32_t)index);
  if (c.type >= PT_SAMPLER_STRATIFIED) return stratified_dim(S, index, dim);
  if (c.type == PT_SAMPLER_HALTON) {  // halton.rs:149-162
    if (c.at_center && dim < 2) return 0.5f;
    if (dim == 0) return radical_inverse2((uint64_t)(index >> c.base_exponents[0]));
    if (dim == 1) return radical_inverse_base(3u, (uint64_t)(index / (int64_t)c.base_scales[1]));
    if (dim >= 1024u) dim = 1023u;  // the reference panics here; paths that long do not occur (maxdepth <= ~100)
    i4 pr = ldgi4(reinterpret_cast<const i4*>(S.prime_recs + dim));
    uint64_t magic = ((uint64_t)(uint32_t)pr.w << 32) | (uint64_t)(uint32_t)pr.z;
    return scrambled_radical_inverse((uint32_t)pr.x, magic, (uint64_t)index, S.perms + (uint32_t)pr.y);
  }
  float s = sobol_sample_float(S.sobol32, index, dim, 0u);  // sobol.rs:174-185
  if (dim < 2) {
    int cp = dim == 0 ? px : py;
    s = s * (float)c.resolution + (float)c.sb[dim];
    s = clampf(s - (float)cp, 0.0f, PT_ONE_MINUS_EPS);
    s = s - truncf(s);
  }
  return s;
}

struct SampleStream {  // the per-path sampler state
  int64_t index;         // Halton / Sobol sample index; Sampler "random": tile slot << 32 | tile sequence number
  uint32_t dim;
  int px, py;
};

PT_HD float get1d(const DScene& S, SampleStream& st) {
  float v = sampler_dim(S, st.index, st.dim, st.px, st.py);
  st.dim += 1;
  return v;
}
PT_HD f2 get2d(const DScene& S, SampleStream& st) {
  float a = sampler_dim(S, st.index, st.dim + S.sampler.x2d, st.px, st.py);
  float b = sampler_dim(S, st.index, st.dim + S.sampler.y2d, st.px, st.py);
  st.dim += S.sampler.inc2d;
  return mk2(a, b);
}

}  // namespace ptd
