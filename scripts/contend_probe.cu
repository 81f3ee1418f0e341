// Does a concurrent tcgen05.mma stream slow down the epilogue's CUDA-core work (and vice versa)?
// Warp 0 issues back-to-back tf32 MMAs (M=128,N=256,K=8, no swizzle) while warps 4..11 run one of:
//  0 DFMA  1 MUFU.EX2  2 widen (LEA+IMAD) + FADD  3 MUFU+widen+DFMA  4 broadcast LDS.128  5 LDTM.x32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double f2d_pos(float k) { unsigned u = __float_as_uint(k); return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29)); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
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
struct Res { long long mma_cycles, comp_cycles; };
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(Res* out, float* sink, int mma_reps, int comp_iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) double sw[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 255);
  if (threadIdx.x < 256) sw[threadIdx.x] = 1e-7 * threadIdx.x;
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
  if (warp == 0) {
    if (lane == 0 && mma_reps > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a0 = s32(smem), b0 = s32(smem) + 32 * 1024;
      const long long t0 = clock64();
      for (int r = 0; r < mma_reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          const uint64_t ad = desc(a0 + ks * 2 * 128 * 16, 128 * 16, 128);
          const uint64_t bd = desc(b0 + ks * 2 * 256 * 16, 256 * 16, 128);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tm + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0) : "memory");
      out[blockIdx.x].mma_cycles = clock64() - t0;
    }
  } else if (warp >= 4) {
    const int cw = warp - 4;
    double acc[8];
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 1e-3 * i;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0xbf000000u + threadIdx.x * 64 + i;
    float facc = 0.f;
    const uint32_t taddr = tm + ((uint32_t)((cw & 3) * 32) << 16) + (cw >> 2) * 64;
    const long long t0 = clock64();
    for (int it = 0; it < comp_iters; ++it) {
      if (MODE == 5) { tmem_ld32(taddr + (it & 1) * 32, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); facc += __uint_as_float(v[it & 31]); }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (MODE == 0) acc[i % 8] = fma(acc[i % 8], 0.999, 1e-7);
        if (MODE == 1) { facc += ex2(__uint_as_float(v[i])); }
        if (MODE == 2) { unsigned u = v[i]; facc += __uint_as_float((u >> 3) + 0x38000000u) + __uint_as_float(u << 29); }
        if (MODE == 3) { acc[i % 8] = fma(f2d_pos(ex2(__uint_as_float(v[i]))), 1e-7, acc[i % 8]); }
        if (MODE == 4) { if ((i & 1) == 0) { double wx, wy; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(wx), "=d"(wy) : "r"(s32(&sw[(it * 32 + i) & 254]))); acc[i % 8] += wx; acc[(i + 1) % 8] += wy; } }
      }
      if (MODE == 1 || MODE == 2 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) v[i] += 1;
      }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    sink[blockIdx.x * 384 + threadIdx.x] = facc + (float)s;
    if (warp == 4 && lane == 0) out[blockIdx.x].comp_cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int MODE>
void run(const char* name, int comp_iters, int mma_per_iter_x100) {
  Res* d; cudaMalloc(&d, sizeof(Res) * 148); cudaMemset(d, 0, sizeof(Res) * 148);
  float* sink; cudaMalloc(&sink, 148 * 384 * 4);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  for (int with_mma = 0; with_mma < 2; ++with_mma) {
    // first pass without MMA gives the compute duration; size the MMA stream to cover it
    static long long base_cycles = 0;
    int mma_reps = 0;
    if (with_mma) mma_reps = (int)(base_cycles / 800) + 1;
    probe<MODE><<<148, 384, 96 * 1024>>>(d, sink, mma_reps, comp_iters);
    cudaError_t e = cudaDeviceSynchronize();
    Res h; cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    if (!with_mma) base_cycles = h.comp_cycles;
    const double per32 = (double)h.comp_cycles / comp_iters;             // cycles per 32 elements per warp
    printf("%-22s mma=%d : %8.1f cycles/iter/warp -> %6.2f elem/clk/SM (8 warps) ; mma %7.1f cycles/instr  (%s)\n", name, with_mma, per32,
           8 * 32 * 32.0 / per32, mma_reps ? (double)h.mma_cycles / (mma_reps * 5.0) : 0.0, cudaGetErrorString(e));
  }
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<0>("dfma", 4000, 0);
  run<1>("mufu", 2000, 0);
  run<2>("widen+fadd", 4000, 0);
  run<3>("mufu+widen+dfma", 2000, 0);
  run<4>("lds.128 broadcast", 4000, 0);
  run<5>("ldtm.x32", 4000, 0);
  return 0;
}
