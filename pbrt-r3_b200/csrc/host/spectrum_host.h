// spectrum_host.h — host-side construction of spectra in the render's colour representation.
// RGB: reference src/core/spectrum/rgb.rs (from_rgb :157, rgb_from_sampled :124-149).
// 60-bin: reference src/core/spectrum/sampled.rs (sampled_from_rgb :150-226, to_xyz :53-67) with the
// tables the reference generates at BUILD time (build/spectrum/{utils,build_cie,build_xyz,build_rgb_refl,
// to_string}.rs: averaged to 60 bins over 400-700 nm, printed with "{:9.8}" and re-parsed as f32 —
// reproduced here by the same snprintf/strtof round trip).
// Raw tables come from pbrt-r3_b200/data/{cie,rgb2spect}.bin (tools/extract_tables.py).
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace pth {

inline std::string data_dir() {
  if (const char* e = getenv("PBRT_R3_B200_DATA")) return e;
  Dl_info info;
  static int anchor;
  if (dladdr((void*)&anchor, &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t s = p.find_last_of('/');
    return (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/data";
  }
  return "data";
}
inline std::vector<float> load_f32(const std::string& name, size_t expect) {
  std::string path = data_dir() + "/" + name;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open table " + path);
  std::vector<float> v(expect);
  size_t n = fread(v.data(), sizeof(float), expect, f);
  fclose(f);
  if (n != expect) throw std::runtime_error("short table " + path);
  return v;
}

class SpectrumFactory {
 public:
  static constexpr int NS = 60;
  static constexpr float CIE_Y_INTEGRAL = 106.856895f;
  explicit SpectrumFactory(int kind) : kind_(kind) {
    if (kind != 0 && kind != 1) throw std::runtime_error("unknown spectrum kind");
  }
  int kind() const { return kind_; }
  size_t n_channels() const { return kind_ == 0 ? 3 : NS; }

  static void xyz_to_rgb(const float* xyz, float* rgb) {  // core/spectrum/convert.rs:3-9
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
  }

  std::vector<float> constant(float v) const { return std::vector<float>(n_channels(), v); }

  std::vector<float> from_rgb(const float* rgb, bool illuminant) {
    if (kind_ == 0) return std::vector<float>{rgb[0], rgb[1], rgb[2]};
    load_sampled_tables();
    const int base = illuminant ? 7 : 0;
    const std::vector<float>*T = &rgb2spect_[(size_t)base];  // white cyan magenta yellow red green blue
    enum { W = 0, C = 1, M = 2, Y = 3, R = 4, G = 5, B = 6 };
    std::vector<float> r((size_t)NS, 0.0f);
    auto add = [&](int which, float s) {
      for (int i = 0; i < NS; i++) r[(size_t)i] += T[which][(size_t)i] * s;
    };
    if (rgb[0] <= rgb[1] && rgb[0] <= rgb[2]) {
      add(W, rgb[0]);
      if (rgb[1] <= rgb[2]) { add(C, rgb[1] - rgb[0]); add(B, rgb[2] - rgb[1]); }
      else { add(C, rgb[2] - rgb[0]); add(G, rgb[1] - rgb[2]); }
    } else if (rgb[1] <= rgb[0] && rgb[1] <= rgb[2]) {
      add(W, rgb[1]);
      if (rgb[0] <= rgb[2]) { add(M, rgb[0] - rgb[1]); add(B, rgb[2] - rgb[0]); }
      else { add(M, rgb[2] - rgb[1]); add(R, rgb[0] - rgb[2]); }
    } else {
      add(W, rgb[2]);
      if (rgb[0] <= rgb[2]) { add(Y, rgb[0] - rgb[2]); add(G, rgb[1] - rgb[0]); }
      else { add(Y, rgb[1] - rgb[2]); add(R, rgb[0] - rgb[1]); }
    }
    const float k = illuminant ? 0.86445f : 0.94f;
    for (auto& v : r) {
      v *= k;
      if (v < 0.0f) v = 0.0f;  // clamp(0, inf)
    }
    return r;
  }

  std::vector<float> from_sampled(std::vector<float> lam, std::vector<float> val) {
    sort_samples(lam, val);
    if (kind_ == 1) return sample_spectrum(lam, val);
    load_cie();
    float xyz[3] = {0, 0, 0};
    for (int i = 0; i < 471; i++) {  // rgb.rs:135-141
      float v = interpolate(lam, val, cie_[3][(size_t)i]);
      xyz[0] += v * cie_[0][(size_t)i];
      xyz[1] += v * cie_[1][(size_t)i];
      xyz[2] += v * cie_[2][(size_t)i];
    }
    float scale = (cie_[3][470] - cie_[3][0]) / (CIE_Y_INTEGRAL * 471.0f);
    for (float& c : xyz) c *= scale;
    float rgb[3];
    xyz_to_rgb(xyz, rgb);
    return std::vector<float>{rgb[0], rgb[1], rgb[2]};
  }

  // "blackbody" parameters (pairs of temperature, scale) go through SampledSpectrum even in an RGB build: SampledSpectrum::
  // from_blackbody (sampled.rs:241-251: blackbody_normalized over CIE_LAMBDA, blackbody.rs:3-33, averaged into the 60 bins)
  // and then Spectrum::from(&SampledSpectrum) = to_rgb for RGBSpectrum (scene_context.rs:763-778, rgb.rs:409-413)
  std::vector<float> from_blackbody(const std::vector<float>& values) {
    load_cie();
    std::vector<float> s((size_t)NS, 0.0f);
    for (size_t i = 0; i + 1 < values.size(); i += 2) {
      const double t = (double)values[i];
      auto planck = [&](double lambda_nm) {
        if (t <= 0.0) return 0.0;
        const double C = 299792458.0, H = 6.62606957e-34, KB = 1.3806488e-23;
        const double l = lambda_nm * 1e-9;
        const double lambda5 = (l * l) * (l * l) * l;
        return (2.0 * H * C * C) / (lambda5 * (std::exp((H * C) / (l * KB * t)) - 1.0));
      };
      const double max_l = planck(2.8977721e-3 / t * 1e9);
      std::vector<float> v(471);
      for (int k = 0; k < 471; k++) v[(size_t)k] = (float)(planck((double)cie_[3][(size_t)k]) / max_l);
      std::vector<float> b = sample_spectrum(cie_[3], v);
      for (int k = 0; k < NS; k++) s[(size_t)k] += b[(size_t)k] * values[i + 1];
    }
    return from_binned(s);
  }
  // Spectrum::from(&SampledSpectrum): the 60 bins themselves, or SampledSpectrum::to_rgb (sampled.rs:53-67, 89-92)
  std::vector<float> from_binned(const std::vector<float>& s60) {
    if (kind_ == 1) return s60;
    load_sampled_tables();
    float xyz[3] = {0, 0, 0};
    for (int i = 0; i < NS; i++) {
      xyz[0] += xyz60_[0][(size_t)i] * s60[(size_t)i];
      xyz[1] += xyz60_[1][(size_t)i] * s60[(size_t)i];
      xyz[2] += xyz60_[2][(size_t)i] * s60[(size_t)i];
    }
    const float scale = (float)(700 - 400) / (CIE_Y_INTEGRAL * (float)NS);
    for (float& c : xyz) c *= scale;
    float rgb[3];
    xyz_to_rgb(xyz, rgb);
    return std::vector<float>{rgb[0], rgb[1], rgb[2]};
  }

  // materials/metal.rs:87-129 default copper eta / k
  std::vector<float> copper_n() { return from_sampled(copper_lambda(), copper(true)); }
  std::vector<float> copper_k() { return from_sampled(copper_lambda(), copper(false)); }

  // Spectrum::to_xyz as weights: xyz[r] = scale * sum_i w[r*60 + i] * c[i]   (rgb.rs:55-63, sampled.rs:53-67)
  void xyz_weights(float* w180, float* scale) {
    for (int i = 0; i < 180; i++) w180[i] = 0.0f;
    if (kind_ == 0) {
      const float m[9] = {0.412453f, 0.357580f, 0.180423f, 0.212671f, 0.715160f, 0.072169f, 0.019334f, 0.119193f, 0.950227f};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) w180[r * 60 + c] = m[r * 3 + c];
      *scale = 1.0f;
    } else {
      load_sampled_tables();
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < NS; c++) w180[r * 60 + c] = xyz60_[(size_t)r][(size_t)c];
      *scale = (float)(700 - 400) / (CIE_Y_INTEGRAL * (float)NS);
    }
  }

 private:
  int kind_;
  std::vector<std::vector<float>> cie_;        // X, Y, Z, LAMBDA (471)
  std::vector<std::vector<float>> rgb2spect_;  // 14 curves resampled to 60 bins
  std::vector<std::vector<float>> xyz60_;      // ARRAY_CIE_X/Y/Z

  static float lerpf(float t, float a, float b) { return (1.0f - t) * a + t * b; }
  static float roundtrip(float v) {  // build/spectrum/to_string.rs "{:9.8}" then Rust literal parse
    char buf[64];
    snprintf(buf, sizeof(buf), "%.8f", (double)v);
    return strtof(buf, nullptr);
  }
  static void sort_samples(std::vector<float>& lam, std::vector<float>& val) {
    bool sorted = true;
    for (size_t i = 0; i + 1 < lam.size(); i++)
      if (lam[i] > lam[i + 1]) sorted = false;
    if (sorted) return;
    std::vector<std::pair<float, float>> v;
    for (size_t i = 0; i < lam.size(); i++) v.emplace_back(lam[i], val[i]);
    std::stable_sort(v.begin(), v.end(), [](const std::pair<float, float>& a, const std::pair<float, float>& b) { return a.first < b.first; });
    for (size_t i = 0; i < lam.size(); i++) { lam[i] = v[i].first; val[i] = v[i].second; }
  }
  static float interpolate(const std::vector<float>& lam, const std::vector<float>& val, float l) {  // utils.rs interpolate_spectrum_samples
    size_t n = lam.size();
    if (l <= lam[0]) return val[0];
    if (l >= lam[n - 1]) return val[n - 1];
    size_t first = 0, len = n;  // find_interval: last index with lam[index] <= l, clamped to [0, n-2]
    while (len > 0) {
      size_t half = len >> 1, mid = first + half;
      if (lam[mid] <= l) { first = mid + 1; len -= half + 1; }
      else len = half;
    }
    size_t off = first == 0 ? 0 : first - 1;
    if (off > n - 2) off = n - 2;
    float t = (l - lam[off]) / (lam[off + 1] - lam[off]);
    return lerpf(t, val[off], val[off + 1]);
  }
  // build/spectrum/utils.rs:8-62
  static float average(const std::vector<float>& lam, const std::vector<float>& val, float ls, float le) {
    size_t n = lam.size();
    if (le <= lam[0]) return val[0];
    if (ls >= lam[n - 1]) return val[n - 1];
    if (n == 1) return val[0];
    float sum = 0.0f;
    if (ls < lam[0]) sum += val[0] * (lam[0] - ls);
    if (le > lam[n - 1]) sum += val[n - 1] * (le - lam[n - 1]);
    size_t i = 0;
    while (ls > lam[i + 1]) i++;
    auto interp = [&](float w, size_t k) { return lerpf((w - lam[k]) / (lam[k + 1] - lam[k]), val[k], val[k + 1]); };
    while (i + 1 < n && le >= lam[i]) {
      float s0 = std::max(ls, lam[i]), s1 = std::min(le, lam[i + 1]);
      sum += 0.5f * (interp(s0, i) + interp(s1, i)) * (s1 - s0);
      i++;
    }
    return sum / (le - ls);
  }
  static std::vector<float> sample_spectrum(const std::vector<float>& lam, const std::vector<float>& val) {  // utils.rs:64-81
    std::vector<float> x((size_t)NS);
    for (int i = 0; i < NS; i++) {
      float wl0 = lerpf((float)i / (float)NS, 400.0f, 700.0f);
      float wl1 = lerpf((float)(i + 1) / (float)NS, 400.0f, 700.0f);
      x[(size_t)i] = average(lam, val, wl0, wl1);
    }
    return x;
  }
  void load_cie() {
    if (!cie_.empty()) return;
    std::vector<float> raw = load_f32("cie.bin", 4 * 471);
    cie_.resize(4);
    for (int k = 0; k < 4; k++) cie_[(size_t)k].assign(raw.begin() + k * 471, raw.begin() + (k + 1) * 471);
    // the runtime CIE tables are the build-printed ones (build_cie.rs -> cie_to_string)
    for (auto& t : cie_) for (auto& v : t) v = roundtrip(v);
  }
  void load_sampled_tables() {
    if (!rgb2spect_.empty()) return;
    std::vector<float> raw = load_f32("rgb2spect.bin", 15 * 32);
    std::vector<float> lam(raw.begin(), raw.begin() + 32);
    rgb2spect_.resize(14);
    for (int k = 0; k < 14; k++) {
      std::vector<float> v(raw.begin() + (k + 1) * 32, raw.begin() + (k + 2) * 32);
      rgb2spect_[(size_t)k] = sample_spectrum(lam, v);
      for (auto& x : rgb2spect_[(size_t)k]) x = roundtrip(x);
    }
    // build_xyz.rs: ARRAY_CIE_{X,Y,Z} = sample_spectrum(CIE_LAMBDA, CIE_{X,Y,Z}) from the RAW cie_data tables
    std::vector<float> craw = load_f32("cie.bin", 4 * 471);
    std::vector<float> clam(craw.begin() + 3 * 471, craw.end());
    xyz60_.resize(3);
    for (int k = 0; k < 3; k++) {
      std::vector<float> v(craw.begin() + k * 471, craw.begin() + (k + 1) * 471);
      xyz60_[(size_t)k] = sample_spectrum(clam, v);
      for (auto& x : xyz60_[(size_t)k]) x = roundtrip(x);
    }
  }
  static std::vector<float> copper_lambda() {
    static const float L[56] = {298.7570554f, 302.4004341f, 306.1337728f, 309.960445f, 313.8839949f, 317.9081487f, 322.036826f, 326.2741526f, 330.6244747f, 335.092373f, 339.6826795f, 344.4004944f, 349.2512056f, 354.2405086f, 359.374429f, 364.6593471f, 370.1020239f, 375.7096303f, 381.4897785f, 387.4505563f, 393.6005651f, 399.9489613f, 406.5055016f, 413.2805933f, 420.2853492f, 427.5316483f, 435.0322035f, 442.8006357f, 450.8515564f, 459.2006593f, 467.8648226f, 476.8622231f, 486.2124627f, 495.936712f, 506.0578694f, 516.6007417f, 527.5922468f, 539.0616435f, 551.0407911f, 563.5644455f, 576.6705953f, 590.4008476f, 604.8008683f, 619.92089f, 635.8162974f, 652.5483053f, 670.1847459f, 688.8009889f, 708.4810171f, 729.3186941f, 751.4192606f, 774.9011125f, 799.8979226f, 826.5611867f, 855.0632966f, 885.6012714f};
    return std::vector<float>(L, L + 56);
  }
  static std::vector<float> copper(bool n) {
    static const float N[56] = {1.400313f, 1.38f, 1.358438f, 1.34f, 1.329063f, 1.325f, 1.3325f, 1.34f, 1.334375f, 1.325f, 1.317812f, 1.31f, 1.300313f, 1.29f, 1.281563f, 1.27f, 1.249062f, 1.225f, 1.2f, 1.18f, 1.174375f, 1.175f, 1.1775f, 1.18f, 1.178125f, 1.175f, 1.172812f, 1.17f, 1.165312f, 1.16f, 1.155312f, 1.15f, 1.142812f, 1.135f, 1.131562f, 1.12f, 1.092437f, 1.04f, 0.950375f, 0.826f, 0.645875f, 0.468f, 0.35125f, 0.272f, 0.230813f, 0.214f, 0.20925f, 0.213f, 0.21625f, 0.223f, 0.2365f, 0.25f, 0.254188f, 0.26f, 0.28f, 0.3f};
    static const float K[56] = {1.662125f, 1.687f, 1.703313f, 1.72f, 1.744563f, 1.77f, 1.791625f, 1.81f, 1.822125f, 1.834f, 1.85175f, 1.872f, 1.89425f, 1.916f, 1.931688f, 1.95f, 1.972438f, 2.015f, 2.121562f, 2.21f, 2.177188f, 2.13f, 2.160063f, 2.21f, 2.249938f, 2.289f, 2.326f, 2.362f, 2.397625f, 2.433f, 2.469187f, 2.504f, 2.535875f, 2.564f, 2.589625f, 2.605f, 2.595562f, 2.583f, 2.5765f, 2.599f, 2.678062f, 2.809f, 3.01075f, 3.24f, 3.458187f, 3.67f, 3.863125f, 4.05f, 4.239563f, 4.43f, 4.619563f, 4.817f, 5.034125f, 5.26f, 5.485625f, 5.717f};
    return n ? std::vector<float>(N, N + 56) : std::vector<float>(K, K + 56);
  }
};

}  // namespace pth
