// Throughput microbenchmark for the SASS ops the traversal kernel leans on (B200, sm_100a).
// Each test: 8 independent dependency chains per thread, 8 warps per SMSP; prints warp-instructions / clock / SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 8
template <int OP>
__global__ void k(float* out, long long* cyc, float a0, float b0, int i0) {
  float x[CH], y = b0, z = a0 * 0.5f;
  unsigned u[CH], w = (unsigned)i0;
  for (int i = 0; i < CH; i++) { x[i] = a0 + i + threadIdx.x; u[i] = i0 + i * 77 + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == 0) x[i] = x[i] * y;                                  // FMUL
      if (OP == 1) x[i] = x[i] + y;                                  // FADD
      if (OP == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y));            // FMNMX
      if (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y), "f"(z)); // FMNMX3
      if (OP == 4) x[i] = x[i] > y ? x[i] : z;                       // FSETP + FSEL
      if (OP == 5) u[i] = __popc(u[i]) + w;                          // POPC (+IADD)
      if (OP == 6) u[i] = (u[i] & w) ^ (u[i] >> 3);                  // LOP3 + SHF
      if (OP == 7) u[i] = u[i] > w ? u[i] + 1 : w;                   // ISETP + SEL-ish
      if (OP == 8) x[i] = 1.0f / x[i];                               // IEEE division (prec-div)
      if (OP == 9) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));               // MUFU.RCP
      if (OP == 10) u[i] = u[i] * w + 3;                             // IMAD
      if (OP == 11) u[i] = __byte_perm(u[i], w, 0x5410 + (u[i] & 3)); // PRMT
      if (OP == 12) x[i] = sqrtf(x[i]);                              // IEEE sqrt
    }
  }
  long long t1 = clock64();
  float s = 0; unsigned su = 0;
  for (int i = 0; i < CH; i++) { s += x[i]; su += u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + su;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int per) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 1024>>>(out, cyc, 1.0001f, 1.00001f, 12345);
  k<OP><<<148, 1024>>>(out, cyc, 1.0001f, 1.00001f, 12345);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  double ops = (double)ITERS * CH * 8 /*warps per SMSP*/;
  printf("%-22s %8.0f cyc  %6.3f ops/clk/SMSP  (%d SASS instr per op: %6.3f instr/clk/SMSP)\n", name, avg, ops / avg, per, ops * per / avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FMUL", 1); run<1>("FADD", 1); run<2>("FMNMX", 1); run<3>("FMNMX3", 1); run<4>("FSETP+FSEL", 2); run<5>("POPC+IADD", 2);
  run<6>("LOP3+SHF", 2); run<7>("ISETP+SEL", 2); run<8>("IEEE 1/x", 1); run<9>("MUFU.RCP", 1); run<10>("IMAD", 1); run<11>("PRMT(+LOP)", 2); run<12>("IEEE sqrt", 1);
  return 0;
}
