// Pipe-throughput probe for the set-sum epilogue's instruction mix (MUFU.EX2 / int widen / DFMA).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/mufu_probe scripts/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double f2d_pos(float k) { unsigned u = __float_as_uint(k); return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29)); }
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, float seed, double w) {
  float x[16];
  double acc[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + 1e-3f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  float facc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float k = ex2(x[i]);
      if (MODE == 0) { facc += k; }
      if (MODE == 1) { acc[i % 8] = fma((double)0.5 + acc[(i + 1) % 8] * 0.0, w, acc[i % 8]); facc += k; }   // MUFU + independent DFMA
      if (MODE == 2) { acc[i % 8] = fma(f2d_pos(k), w, acc[i % 8]); }                                         // full mix
      if (MODE == 3) { unsigned u = __float_as_uint(k); facc += __uint_as_float((u >> 3) + 0x38000000u) + __uint_as_float(u << 29); }  // MUFU + int ops + FADD
      if (MODE == 4) { acc[i % 8] = fma(__hiloint2double(__float_as_int(x[i]) , it + i), w, acc[i % 8]); if (i == 0) facc += k; }      // DFMA mostly
      x[i] = x[i] * 0.999f - 0.001f;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = facc + (float)s;
}
template <int MODE>
void run(const char* name, float* out) {
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148, 512>>>(out, 64, 0.1f, 1e-7);
  cudaEventRecord(e0);
  probe<MODE><<<148, 512>>>(out, iters, 0.1f, 1e-7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 512 * 16.0 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %7.2f ex2/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  run<0>("mufu + fadd + ffma", out);
  run<1>("mufu + indep dfma", out);
  run<2>("mufu + widen + dfma", out);
  run<3>("mufu + int ops + fadd", out);
  run<4>("dfma mostly", out);
  return 0;
}
