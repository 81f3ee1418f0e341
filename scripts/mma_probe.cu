// tcgen05.mma issue-rate probe: cycles per MMA instruction (M=128, N, K = 32 bytes) for tf32 / f16
// kinds and shared-memory operand layouts (no swizzle vs 32/64/128-byte swizzle).  Data is garbage.
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
template <int KIND>  // 0 tf32, 1 f16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
struct Res { long long cycles; };
template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(Res* out, int N, int layout, int ksteps, int reps) {
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
    const uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 / bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = s32(smem), b0 = s32(smem) + 64 * 1024;
    // layout 0: [K/16B][rows][16 B]: LBO = rows*16, SBO = 128, K-step advance = 2 planes
    // layout 6/4/2 (32/64/128 B swizzle): rows of 32/64/128 B, 8-row atoms: SBO = 8 * rowbytes; K-step advance = 32 B inside the row
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < ksteps; ++ks) {
        uint64_t ad, bd;
        if (layout == 0) {
          ad = desc(a0 + ks * 2 * 128 * 16, 128 * 16, 128, 0);
          bd = desc(b0 + ks * 2 * N * 16, N * 16, 128, 0);
        } else {
          const int rowb = layout == 6 ? 32 : (layout == 4 ? 64 : 128);
          const int per_row = rowb / 32;                       // K-steps inside one swizzled row
          const int blk = ks / per_row, in = ks % per_row;     // K blocks of rowb bytes are stored one after another
          ad = desc(a0 + blk * 128 * rowb + in * 32, 16, 8 * rowb, layout);
          bd = desc(b0 + blk * N * rowb + in * 32, 16, 8 * rowb, layout);
        }
        mma<KIND>(tm + (r & 1) * 256, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0) : "memory");
    const long long t1 = clock64();
    out[blockIdx.x].cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int KIND>
void run(const char* name, int N, int layout, int ksteps, int grid) {
  Res* d; cudaMalloc(&d, sizeof(Res) * grid);
  const int reps = 2000;
  cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  probe<KIND><<<grid, 128, 160 * 1024>>>(d, N, layout, ksteps, 10);
  probe<KIND><<<grid, 128, 160 * 1024>>>(d, N, layout, ksteps, reps);
  cudaError_t e = cudaDeviceSynchronize();
  Res h; cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
  const double per = (double)h.cycles / ((double)reps * ksteps);
  const int kel = KIND == 0 ? 8 : 16;
  printf("%-6s N=%3d layout=%d ksteps=%d grid=%3d : %7.1f cycles/MMA  %7.0f MAC/clk/SM  (%s)\n", name, N, layout, ksteps, grid, per,
         128.0 * N * kel / per, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int grid : {1, 148}) {
    for (int layout : {0, 6, 4, 2}) run<0>("tf32", 256, layout, 5, grid);
    for (int layout : {0, 6, 4, 2}) run<1>("bf16", 256, layout, 4, grid);
    run<0>("tf32", 128, 0, 5, grid);
    run<0>("tf32", 128, 2, 4, grid);
    run<0>("tf32", 64, 0, 5, grid);
  }
  return 0;
}
