// Grid-wide barrier for persistent cooperative kernels (all CTAs co-resident).
#pragma once
#include "common.cuh"

namespace basq {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Monotone-counter grid barrier (all CTAs are co-resident: cooperative launch).  A watchdog turns
// a would-be hang into an error flag; the return value (uniform over the CTA) says "abort".
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned& target, int* status, int* abort_sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    target += gridDim.x;
    atomicAdd(bar, 1u);
    long long spins = 0;
    while ((int)(ld_acquire_u32(bar) - target) < 0) {
      ++spins;
      if ((spins & 1023) == 0) {
        if (*reinterpret_cast<volatile int*>(status) != 0) break;
        if (spins > (1ll << 26)) { atomicExch(status, 2); break; }
      }
    }
    __threadfence();
    *abort_sh = *reinterpret_cast<volatile int*>(status);
  }
  __syncthreads();
  return *abort_sh != 0;
}

}  // namespace basq
