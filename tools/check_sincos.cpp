// Checks the product's pt_sinf / pt_cosf (pbrt-r3_b200/csrc/cuda/dmath.cuh, compiled for the host) bit-for-bit against
// the host libm (glibc sinf / cosf = what Rust's f32::sin / f32::cos call on Linux).
//   g++ -O2 -ffp-contract=off tools/check_sincos.cpp -o /tmp/check_sincos && /tmp/check_sincos [N]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../pbrt-r3_b200/csrc/cuda/dmath.cuh"

int main(int argc, char** argv) {
  long N = argc > 1 ? atol(argv[1]) : 20000000, bad_s = 0, bad_c = 0;
  std::mt19937 g(1);
  std::uniform_real_distribution<float> d(-7.0f, 7.0f);
  for (long i = 0; i < N; i++) {
    float x = d(g);
    if (i % 4 == 0) x *= 1e-3f;
    if (i % 1000 == 1) x = (float)(i % 13) * 0.78539816339744830962f;  // multiples of pi/4 (branch boundaries)
    if (ptd::fbits(ptd::pt_sinf(x)) != ptd::fbits(sinf(x))) bad_s++;
    if (ptd::fbits(ptd::pt_cosf(x)) != ptd::fbits(cosf(x))) bad_c++;
  }
  printf("%ld %ld %ld\n", bad_s, bad_c, N);
  return (bad_s || bad_c) ? 1 : 0;
}
