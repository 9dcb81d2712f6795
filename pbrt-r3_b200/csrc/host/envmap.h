// envmap.h — host side of LightSource "infinite": the MIPMap<RGBSpectrum> pyramid of the environment map and the
// Distribution2D the light samples from.  Replaces reference src/lights/infinite.rs:67-93 (make_distribution), 103-108
// (power), src/core/texture/mipmap.rs:225-275,406-441 (box-filter pyramid), 620-637 (lookup), 711-767 (triangle) with
// the wrap modes of create_spectrum_mipmap (:1013-1026: s repeats, t clamps), src/core/sampling/distribution.rs:34-58,
// 110-127.  Maps whose sides are not powers of two are first resampled up to the next powers of two with a four-tap Lanczos
// filter, first along s (repeat) then along t (clamp) (mipmap.rs:291-404, noise.rs:150-162).
// The device reads level 0 (every lookup on the path has width 0, which selects the bilinear "triangle" filter of level 0)
// and the distribution tables; the pyramid only exists here, for make_distribution and power().
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace pth {

struct EnvMapHost {
  struct Level { int w, h; std::vector<float> rgb; };
  std::vector<Level> pyramid;
  // Distribution2D over the 2w x 2h image of luminance * sin(theta)
  int nu = 0, nv = 0;
  std::vector<float> func;      // [nv][nu]
  std::vector<float> cdf;       // [nv][nu + 1]
  std::vector<float> row_int;   // [nv]  conditional_v[v].func_int = marginal.func
  std::vector<float> mcdf;      // [nv + 1]
  float m_int = 0.0f;           // marginal.func_int

  static float lum(const float* c) { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }  // RGBSpectrum::y (rgb.rs)

  // one output sample of the resampling filter: four taps starting at texel `first`
  struct Taps { int first; float w[4]; };
  static float lanczos2(float x) {  // lanczos(x, tau = 2) (noise.rs:150-162)
    const float PI = 3.14159265358979323846f, tau = 2.0f;
    x = fabsf(x);
    if (x < 1e-5f) return 1.0f;
    if (x > 1.0f) return 0.0f;
    x *= PI;
    const float s = sinf(x * tau) / (x * tau);
    return s * (sinf(x) / x);
  }
  static std::vector<Taps> taps_for(int old_res, int new_res) {  // resample_weights (mipmap.rs:291-314)
    std::vector<Taps> out((size_t)new_res);
    for (int i = 0; i < new_res; i++) {
      const float center = ((float)i + 0.5f) * ((float)old_res / (float)new_res);
      const float first = floorf((center - 2.0f) + 0.5f);
      Taps& T = out[(size_t)i];
      T.first = (int)first;
      for (int j = 0; j < 4; j++) T.w[j] = lanczos2(((first + (float)j + 0.5f) - center) / 2.0f);
      const float inv = 1.0f / (T.w[0] + T.w[1] + T.w[2] + T.w[3]);
      for (int j = 0; j < 4; j++) T.w[j] *= inv;
    }
    return out;
  }
  static int round_up_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
  // resample_image (mipmap.rs:316-404): rows are stretched first (s wraps around), then columns (t clamps); negative lobes of
  // the filter are cut off at zero at the end
  static Level resample_to_pow2(int w, int h, const float* texels) {
    Level R{round_up_pow2(w), round_up_pow2(h), {}};
    if (R.w == w && R.h == h) { R.rgb.assign(texels, texels + (size_t)w * h * 3); return R; }
    R.rgb.assign((size_t)R.w * R.h * 3, 0.0f);
    const std::vector<Taps> ts = taps_for(w, R.w), tt = taps_for(h, R.h);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < R.w; x++)
        for (int j = 0; j < 4; j++) {
          const int sx = ((ts[(size_t)x].first + j) % w + w) % w;
          for (int c = 0; c < 3; c++) R.rgb[3 * ((size_t)y * R.w + x) + c] += texels[3 * ((size_t)y * w + sx) + c] * ts[(size_t)x].w[j];
        }
    std::vector<float> column((size_t)h * 3);
    for (int x = 0; x < R.w; x++) {
      for (int y = 0; y < h; y++)
        for (int c = 0; c < 3; c++) column[3 * (size_t)y + c] = R.rgb[3 * ((size_t)y * R.w + x) + c];
      for (int y = 0; y < R.h; y++) {
        float acc[3] = {0.0f, 0.0f, 0.0f};
        for (int j = 0; j < 4; j++) {
          const int sy = std::min(std::max(tt[(size_t)y].first + j, 0), h - 1);
          for (int c = 0; c < 3; c++) acc[c] += column[3 * (size_t)sy + c] * tt[(size_t)y].w[j];
        }
        for (int c = 0; c < 3; c++) R.rgb[3 * ((size_t)y * R.w + x) + c] = acc[c];
      }
    }
    for (float& v : R.rgb) v = std::max(v, 0.0f);
    return R;
  }

  void build_pyramid(int w, int h, const float* texels) {  // make_pyramid: halve x, then y, until 1 x 1
    pyramid.clear();
    Level L = resample_to_pow2(w, h, texels);
    for (;;) {
      pyramid.push_back(L);
      if (L.w * L.h == 1) break;
      Level X = L;
      if (L.w > 1) {
        X.w = L.w / 2;
        X.rgb.assign((size_t)X.w * X.h * 3, 0.0f);
        for (int y = 0; y < L.h; y++)
          for (int x = 0; x < X.w; x++)
            for (int c = 0; c < 3; c++)
              X.rgb[3 * ((size_t)y * X.w + x) + c] = L.rgb[3 * ((size_t)y * L.w + 2 * x) + c] * 0.5f + L.rgb[3 * ((size_t)y * L.w + 2 * x + 1) + c] * 0.5f;
      }
      Level Y = X;
      if (X.h > 1) {
        Y.h = X.h / 2;
        Y.rgb.assign((size_t)Y.w * Y.h * 3, 0.0f);
        for (int y = 0; y < Y.h; y++)
          for (int x = 0; x < Y.w; x++)
            for (int c = 0; c < 3; c++)
              Y.rgb[3 * ((size_t)y * Y.w + x) + c] = X.rgb[3 * ((size_t)(2 * y) * X.w + x) + c] * 0.5f + X.rgb[3 * ((size_t)(2 * y + 1) * X.w + x) + c] * 0.5f;
      }
      L = Y;
    }
  }
  void triangle(int level, float s_, float t_, float out[3]) const {  // mipmap.rs:711-767
    level = std::min(std::max(level, 0), (int)pyramid.size() - 1);
    const Level& I = pyramid[(size_t)level];
    float s = s_ * (float)I.w - 0.5f, t = t_ * (float)I.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    float ds = s - (float)s0, dt = t - (float)t0;
    int s1 = s0 + 1, t1 = t0 + 1;
    s0 &= I.w - 1; s1 &= I.w - 1;
    t0 = std::min(std::max(t0, 0), I.h - 1); t1 = std::min(std::max(t1, 0), I.h - 1);
    const float wa = (1.0f - ds) * (1.0f - dt), wb = (1.0f - ds) * dt, wc = ds * (1.0f - dt), wd = ds * dt;
    for (int c = 0; c < 3; c++) {
      float a = I.rgb[3 * ((size_t)t0 * I.w + s0) + c] * wa, b = I.rgb[3 * ((size_t)t1 * I.w + s0) + c] * wb;
      float cc = I.rgb[3 * ((size_t)t0 * I.w + s1) + c] * wc, d = I.rgb[3 * ((size_t)t1 * I.w + s1) + c] * wd;
      out[c] = a + b + cc + d;
    }
  }
  void lookup(float s, float t, float width, float out[3]) const {  // mipmap.rs:620-637
    const float max_level = (float)((int)pyramid.size() - 1);
    const float level = max_level + log2f(std::max(width, 1e-8f));
    if (level < 0.0f) { triangle(0, s, t, out); return; }
    if (level >= max_level) { const Level& I = pyramid.back(); for (int c = 0; c < 3; c++) out[c] = I.rgb[c]; return; }
    const int il = (int)floorf(level);
    const float delta = std::min(std::max(level - (float)il, 0.0f), 1.0f);
    float a[3], b[3];
    triangle(il, s, t, a);
    triangle(il + 1, s, t, b);
    for (int c = 0; c < 3; c++) out[c] = a[c] * (1.0f - delta) + b[c] * delta;
  }
  // Distribution1D::new over f[0..n) -> cdf[0..n], returns func_int
  static float dist1d(const float* f, int n, float* cdf) {
    cdf[0] = 0.0f;
    for (int i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + f[i - 1] / (float)n;
    const float fi = cdf[n];
    if (fi == 0.0f) { for (int i = 1; i < n + 1; i++) cdf[i] = (float)i / (float)n; }
    else { for (int i = 1; i < n + 1; i++) cdf[i] /= fi; }
    return fi;
  }
  void build_distribution() {  // infinite.rs:67-93
    const float PI = 3.14159265358979323846f;
    nu = pyramid[0].w * 2; nv = pyramid[0].h * 2;
    func.assign((size_t)nu * nv, 0.0f);
    const float fwidth = 0.5f / (float)std::min(nu, nv);
    for (int v = 0; v < nv; v++) {
      const float vp = ((float)v + 0.5f) / (float)nv;
      const float sin_theta = sinf(PI * vp);
      for (int u = 0; u < nu; u++) {
        const float up = ((float)u + 0.5f) / (float)nu;
        float c[3];
        lookup(up, vp, fwidth, c);
        func[(size_t)v * nu + u] = std::max(lum(c), 0.0f) * sin_theta;
      }
    }
    cdf.assign((size_t)nv * (nu + 1), 0.0f);
    row_int.assign((size_t)nv, 0.0f);
    for (int v = 0; v < nv; v++) row_int[(size_t)v] = dist1d(func.data() + (size_t)v * nu, nu, cdf.data() + (size_t)v * (nu + 1));
    mcdf.assign((size_t)nv + 1, 0.0f);
    m_int = dist1d(row_int.data(), nv, mcdf.data());
  }
};

}  // namespace pth
