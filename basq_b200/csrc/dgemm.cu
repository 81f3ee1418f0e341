// fp64 GEMM on the CUDA cores (B200 has no fp64 tcgen05 path; DFMA peak is 64/clk/SM).
// Used for the per-round projection U' @ G (reference BASQ/_rchq.py:88), the GP cache products
// (K_ZX W, BASQ/_gp.py:270-273) and the Nystrom subspace iteration.
// 64x64 CTA tile, BK = 16, 256 threads, 4x4 register tile, register-staged double buffering.
#include "common.cuh"

namespace basq {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int64_t lda, const double* __restrict__ B, int64_t ldb, double beta,
                                                    double* __restrict__ C, int64_t ldc) {
  __shared__ __align__(16) double As[BK][BM + PAD];
  __shared__ __align__(16) double Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // global -> register staging maps (coalesced along the contiguous index of each operand)
  int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!TA) { a_k[i] = tid & 15; a_m[i] = (tid >> 4) + 16 * i; } else { a_m[i] = tid & 63; a_k[i] = (tid >> 6) + 4 * i; }
    if (!TB) { b_n[i] = tid & 63; b_k[i] = (tid >> 6) + 4 * i; } else { b_k[i] = tid & 15; b_n[i] = (tid >> 4) + 16 * i; }
  }
  auto fetchA = [&](int k0, double* r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + a_m[i], k = k0 + a_k[i];
      r[i] = (m < M && k < K) ? (TA ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k]) : 0.0;
    }
  };
  auto fetchB = [&](int k0, double* r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + b_n[i], k = k0 + b_k[i];
      r[i] = (n < N && k < K) ? (TB ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) : 0.0;
    }
  };

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  double ra[4], rb[4];
  fetchA(0, ra);
  fetchB(0, rb);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[a_k[i]][a_m[i]] = ra[i];
      Bs[b_k[i]][b_n[i]] = rb[i];
    }
    __syncthreads();
    if (k0 + BK < K) {
      fetchA(k0 + BK, ra);
      fetchB(k0 + BK, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y};
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      double* c = C + (int64_t)m * ldc + n;
      *c = (beta == 0.0) ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * (*c));
    }
  }
}
}  // namespace

int dgemm(basq_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  if (m <= 0 || n <= 0) return BASQ_OK;
  BASQ_CHECK(k >= 0, BASQ_ERR_INVALID, "dgemm: negative k");
  dim3 grid((unsigned)ceil_div(n, BN), (unsigned)ceil_div(m, BM));
  BASQ_CHECK(grid.y <= 65535, BASQ_ERR_UNSUPPORTED, "dgemm: m=%d too large for one launch", m);
  if (!ta && !tb)
    dgemm_kernel<false, false><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (ta && !tb)
    dgemm_kernel<true, false><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!ta && tb)
    dgemm_kernel<false, true><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else
    dgemm_kernel<true, true><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

}  // namespace basq
