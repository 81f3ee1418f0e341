// TMEM-read throughput probe: tcgen05.ld alone and mixed with the set-sum epilogue's MUFU/DFMA mix.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ldtm_probe scripts/ldtm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double f2d_pos(float k) { unsigned u = __float_as_uint(k); return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29)); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// MODE 0: LDTM only (xor-reduce the values)   MODE 1: LDTM + ex2 + widen + DFMA per value   MODE 2: same without LDTM
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(float* out, int iters, double w) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32) % 512;
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  uint32_t x = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0x3dcccccd + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE != 2) {
      tmem_ld32(base + (uint32_t)((it * 32) & 255), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x ^= v[i];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float k = ex2(__uint_as_float((v[i] & 0x007fffffu) | 0xbf000000u));   // argument in (-1, -0.5]
        acc[i % 8] = fma(f2d_pos(k), w, acc[i % 8]);
        if (MODE == 2) v[i] += 3;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s + __uint_as_float(x);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
template <int MODE, int WARPS>
void run(const char* name, float* out) {
  const int iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE, WARPS><<<148, WARPS * 32>>>(out, 64, 1e-7);
  cudaEventRecord(e0);
  probe<MODE, WARPS><<<148, WARPS * 32>>>(out, iters, 1e-7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  double vals = 148.0 * WARPS * 32 * 32.0 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s warps=%2d %8.3f ms  %7.2f values/clk/SM (%s)\n", name, WARPS, ms, vals / (ms * 1e-3) / 148 / (clk * 1e3), cudaGetErrorString(e));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<0, 4>("ldtm only", out);
  run<0, 8>("ldtm only", out);
  run<0, 16>("ldtm only", out);
  run<1, 8>("ldtm + ex2 + widen + dfma", out);
  run<1, 16>("ldtm + ex2 + widen + dfma", out);
  run<2, 8>("ex2 + widen + dfma (no ldtm)", out);
  run<2, 16>("ex2 + widen + dfma (no ldtm)", out);
  return 0;
}
