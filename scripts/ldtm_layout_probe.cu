// Which (lane, column) does each register of tcgen05.ld.16x256b hold?  TMEM is filled through
// tcgen05.st.32x32b (lane = row, register i = column i) with value 1000 * lane + column, then read back.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ldtm_layout_probe scripts/ldtm_layout_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(int* out) {
  __shared__ uint32_t slot;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  uint32_t v[16];
  for (int i = 0; i < 16; ++i) v[i] = 1000 * threadIdx.x + i;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t addr = base + ((uint32_t)(half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(half * 32 + threadIdx.x) * 8 + i] = (int)r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(32) : "memory");
}
int main() {
  int* out; cudaMalloc(&out, 64 * 8 * 4);
  probe<<<1, 32>>>(out);
  int h[64 * 8];
  cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status: %s\n", cudaGetErrorString(e));
  for (int half = 0; half < 2; ++half)
    for (int t = 0; t < 32; ++t) {
      printf("half %d thread %2d:", half, t);
      for (int i = 0; i < 8; ++i) printf(" r%d=(%d,%d)", i, h[(half * 32 + t) * 8 + i] / 1000, h[(half * 32 + t) * 8 + i] % 1000);
      printf("\n");
    }
  return 0;
}
