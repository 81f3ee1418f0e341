// tcgen05.mma throughput probe: cycles per MMA instruction (M=128, N, K = 32 bytes) for tf32 / bf16
// kinds, shared-memory operand layouts (no swizzle vs 128-byte swizzle) and 1/2 accumulators.
// Descriptors are precomputed and the issue loop is fully unrolled so that the single issuing
// thread is not the bottleneck.  Operand data is garbage (timing only).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int KIND, int N, int LAYOUT, int KSTEPS, int NACC>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 255);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 / bf16
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = s32(smem), b0 = s32(smem) + 64 * 1024;
    uint64_t ad[KSTEPS], bd[KSTEPS];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      if (LAYOUT == 0) {            // [K/16B][rows][16 B]: LBO = rows*16, SBO = 128
        ad[ks] = desc(a0 + ks * 2 * 128 * 16, 128 * 16, 128, 0);
        bd[ks] = desc(b0 + ks * 2 * N * 16, N * 16, 128, 0);
      } else {                      // 128-byte swizzle, rows of 128 B, 8-row atoms of 1024 B
        const int blk = ks / 4, in = ks % 4;
        ad[ks] = desc(a0 + blk * 128 * 128 + in * 32, 16, 1024, 2);
        bd[ks] = desc(b0 + blk * N * 128 + in * 32, 16, 1024, 2);
      }
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) mma<KIND>(tm + (uint32_t)((ks % NACC) * N), ad[ks], bd[ks], idesc);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int KIND, int N, int LAYOUT, int KSTEPS, int NACC>
void run(const char* name) {
  long long* d; cudaMalloc(&d, sizeof(long long) * 148);
  const int reps = 2000;
  auto k = probe<KIND, N, LAYOUT, KSTEPS, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k<<<148, 128, 160 * 1024>>>(d, 10);
  k<<<148, 128, 160 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)reps * KSTEPS);
  const int kel = KIND == 0 ? 8 : 16;
  printf("%-5s N=%3d layout=%s ksteps=%d nacc=%d : %7.1f cycles/MMA  %6.0f MAC/clk/SM  (%s)\n", name, N, LAYOUT ? "sw128" : "none ", KSTEPS, NACC,
         per, 128.0 * N * kel / per, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0, 256, 0, 4, 1>("tf32"); run<0, 256, 2, 4, 1>("tf32"); run<0, 256, 0, 4, 2>("tf32"); run<0, 256, 2, 4, 2>("tf32");
  run<0, 128, 0, 4, 1>("tf32"); run<0, 128, 2, 4, 1>("tf32"); run<0, 128, 0, 4, 2>("tf32"); run<0, 128, 2, 4, 2>("tf32"); run<0, 128, 0, 4, 4>("tf32");
  run<0, 64, 0, 4, 1>("tf32");  run<0, 64, 0, 4, 4>("tf32");
  run<1, 256, 0, 4, 1>("bf16"); run<1, 256, 2, 4, 1>("bf16"); run<1, 128, 0, 4, 2>("bf16"); run<1, 128, 2, 4, 2>("bf16");
  return 0;
}
