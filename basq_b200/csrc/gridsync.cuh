// Grid-wide barrier for persistent cooperative kernels (all CTAs co-resident).
#pragma once
#include "common.cuh"

namespace basq {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add_acq_rel_u32(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// Arrival counter and release flag live in DIFFERENT 128-byte lines: bar[0] = monotone arrival
// count, bar[32] = generation flag.  The last arriver of a generation publishes the flag; everyone
// else polls the flag (with a short sleep), so the polls never queue in front of the arrival
// atomics at the L2 slice.  `gen` counts barriers passed (per thread 0).  A watchdog turns a
// would-be hang into an error flag; the return value (uniform over the CTA) says "abort".
constexpr int BASQ_BAR_WORDS = 64;  // unsigned words to reserve (and zero) per barrier

__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned& gen, int* status, int* abort_sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gen += 1;
    const unsigned old = atom_add_acq_rel_u32(bar, 1u);
    if (old + 1u == gen * gridDim.x) {
      st_release_u32(bar + 32, gen);
    } else {
      long long spins = 0;
      while ((int)(ld_acquire_u32(bar + 32) - gen) < 0) {
        __nanosleep(20);
        if ((++spins & 4095) == 0) {
          if (*reinterpret_cast<volatile int*>(status) != 0) break;
          if (spins > (1ll << 26)) { atomicExch(status, 2); break; }
        }
      }
    }
    *abort_sh = *reinterpret_cast<volatile int*>(status);
  }
  __syncthreads();
  return *abort_sh != 0;
}

// ---------------------------------------------------------------------------------------------
// "The data is the flag": a producer streams results into an L2 buffer that the host pre-filled with
// 0xFF bytes (a NaN for doubles, -1 for ints); a consumer polls exactly the words it needs until
// they are no longer the sentinel.  No fences and no barriers: 8-byte stores are atomic and every
// word is validated on its own.  A watchdog turns a dead chain into an error flag.
// ---------------------------------------------------------------------------------------------
constexpr int BASQ_SPIN_LIMIT = 1 << 22;  // polls before a consumer declares the chain dead

__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == -1ll; }

__device__ __forceinline__ double poll_f64(const double* p, int* status) {
  double v = __ldcg(p);
  int spins = 0;
  while (is_sentinel(v)) {
    if ((++spins & 255) == 0 && (*reinterpret_cast<volatile int*>(status) != 0 || spins > BASQ_SPIN_LIMIT)) {
      if (spins > BASQ_SPIN_LIMIT) atomicExch(status, 2);
      return 0.0;
    }
    v = __ldcg(p);
  }
  return v;
}
__device__ __forceinline__ int poll_i32(const int* p, int* status) {
  int v = __ldcg(p);
  int spins = 0;
  while (v == -1) {
    if ((++spins & 255) == 0 && (*reinterpret_cast<volatile int*>(status) != 0 || spins > BASQ_SPIN_LIMIT)) {
      if (spins > BASQ_SPIN_LIMIT) atomicExch(status, 2);
      return 1;
    }
    v = __ldcg(p);
  }
  return v;
}

}  // namespace basq
