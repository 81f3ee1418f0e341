// FP64 pipe probe: DFMA alone, DFMA + the 2-op fp32->fp64 widening, DFMA with broadcast LDS.128 operands.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double f2d_pos(float k) { unsigned u = __float_as_uint(k); return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29)); }
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(float* out, int iters, double w0, const double* wg) {
  __shared__ double sw[256];
  if (threadIdx.x < 256) sw[threadIdx.x] = wg[threadIdx.x];
  __syncthreads();
  double acc[16];
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = 1e-3 * i; x[i] = 0.5f + 1e-3f * (threadIdx.x + i); }
  double w = w0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) acc[i] = fma(acc[i], w, w0);                         // DFMA only, 16 chains
        if (MODE == 1) acc[i % 8] = fma(acc[i % 8], w, w0);                 // DFMA only, 8 chains
        if (MODE == 2) acc[i % 8] = fma(f2d_pos(x[i]), w, acc[i % 8]);      // widen + DFMA, 8 chains
        if (MODE == 3) acc[i % 8] = fma(f2d_pos(x[i]), sw[(it * 32 + r * 16 + i) & 255], acc[i % 8]);  // + LDS operand
      }
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __int_as_float(__float_as_int(x[i]) + 1);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
template <int MODE, int WARPS>
void run(const char* name, float* out, const double* wg) {
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE, WARPS><<<148, WARPS * 32>>>(out, 64, 0.999, wg);
  cudaEventRecord(e0);
  probe<MODE, WARPS><<<148, WARPS * 32>>>(out, iters, 0.999, wg);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * WARPS * 32 * 32.0 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-30s warps=%2d %8.3f ms  %7.2f dfma/clk/SM\n", name, WARPS, ms, ops / (ms * 1e-3) / 148 / (clk * 1e3));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  double* wg; cudaMalloc(&wg, 256 * 8); cudaMemset(wg, 0, 256 * 8);
  run<0, 8>("dfma 16 chains", out, wg);   run<0, 16>("dfma 16 chains", out, wg);
  run<1, 8>("dfma 8 chains", out, wg);    run<1, 16>("dfma 8 chains", out, wg);
  run<2, 8>("widen + dfma", out, wg);     run<2, 16>("widen + dfma", out, wg);
  run<3, 8>("widen + dfma + lds", out, wg); run<3, 16>("widen + dfma + lds", out, wg);
  return 0;
}
