// fp64 GEMM for the per-round projection U' @ G (reference BASQ/_rchq.py:88), the GP cache products
// (K_ZX W, BASQ/_gp.py:270-273) and the Nystrom subspace iteration.
//
// There is no fp64 kind in tcgen05; on sm_100a the fp64 tensor path is the warp-level
// mma.sync.m8n8k4.f64 (DMMA), which runs at twice the DFMA-pipe rate measured on this part
// (a register-tiled DFMA version of this kernel topped out at 14 TFLOP/s).
// CTA tile (16*MF) x 128, BK = 8, 8 warps as 2 x 4, warp tile (8*MF) x 32 = MF x 4 fragments of
// 8 x 8, accumulators in registers.  Operands are staged global -> registers -> shared with the next
// tile's loads in flight during the current tile's MMAs; shared layout [k/4][row][k%4] makes every
// fragment load one contiguous 256-byte warp access (conflict-free).  All four transpose
// combinations, arbitrary sizes and leading dimensions.
#include "common.cuh"

namespace basq {

namespace {
constexpr int BN = 128, BK = 8;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool TA, bool TB, int MF>
__global__ void __launch_bounds__(256) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int64_t lda, const double* __restrict__ B, int64_t ldb, double beta,
                                                    double* __restrict__ C, int64_t ldc, int tri_k) {
  constexpr int BM = 16 * MF;
  constexpr int AV = BM * BK / 256;  // A elements staged per thread
  constexpr int BV = BN * BK / 256;  // B elements staged per thread
  __shared__ __align__(16) double As[2][BK / 4][BM][4];
  __shared__ __align__(16) double Bs[2][BK / 4][BN][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // tri_k: op(B)[k][n] is zero for k > n (B lower triangular, used transposed): stop at the tile's last column
  if (tri_k) K = min(K, n0 + BN);

  // staging maps: consecutive threads walk the contiguous index of each operand
  int a_m[AV], a_k[AV], b_n[BV], b_k[BV];
#pragma unroll
  for (int i = 0; i < AV; ++i) {
    const int e = tid + i * 256;
    if (!TA) { a_k[i] = e % BK; a_m[i] = e / BK; } else { a_m[i] = e % BM; a_k[i] = e / BM; }
  }
#pragma unroll
  for (int i = 0; i < BV; ++i) {
    const int e = tid + i * 256;
    if (!TB) { b_n[i] = e % BN; b_k[i] = e / BN; } else { b_k[i] = e % BK; b_n[i] = e / BK; }
  }
  auto fetchA = [&](int k0, double* r) {
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      const int m = m0 + a_m[i], k = k0 + a_k[i];
      r[i] = (m < M && k < K) ? (TA ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k]) : 0.0;
    }
  };
  auto fetchB = [&](int k0, double* r) {
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int n = n0 + b_n[i], k = k0 + b_k[i];
      r[i] = (n < N && k < K) ? (TB ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) : 0.0;
    }
  };
  auto stage = [&](int buf, const double* ra, const double* rb) {
#pragma unroll
    for (int i = 0; i < AV; ++i) As[buf][a_k[i] >> 2][a_m[i]][a_k[i] & 3] = ra[i];
#pragma unroll
    for (int i = 0; i < BV; ++i) Bs[buf][b_k[i] >> 2][b_n[i]][b_k[i] & 3] = rb[i];
  };

  double acc[MF][4][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[AV], rb[BV];
  fetchA(0, ra);
  fetchB(0, rb);
  stage(0, ra, rb);
  __syncthreads();
  int buf = 0;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) {
      fetchA(k0 + BK, ra);
      fetchB(k0 + BK, rb);
    }
#pragma unroll
    for (int kc = 0; kc < BK / 4; ++kc) {
      double af[MF], bf[4];
#pragma unroll
      for (int i = 0; i < MF; ++i) af[i] = As[buf][kc][wm * (8 * MF) + i * 8 + fr][fk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][kc][wn * 32 + j * 8 + fr][fk];
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (more) {
      stage(buf ^ 1, ra, rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  // C fragment: row = lane / 4, columns = (lane % 4) * 2 + {0, 1}
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int m = m0 + wm * (8 * MF) + i * 8 + fr;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + wn * 32 + j * 8 + fk * 2 + h;
        if (n >= N) continue;
        double* c = C + (int64_t)m * ldc + n;
        *c = (beta == 0.0) ? alpha * acc[i][j][h] : fma(alpha, acc[i][j][h], beta * (*c));
      }
    }
  }
}

template <bool TA, bool TB>
void launch(basq_ctx* ctx, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B,
            int64_t ldb, double beta, double* C, int64_t ldc, int tri_k) {
  const int64_t tiles128 = (int64_t)ceil_div(m, 128) * ceil_div(n, BN);
  if (tiles128 >= ctx->num_sms / 2) {
    dim3 grid((unsigned)ceil_div(n, BN), (unsigned)ceil_div(m, 128));
    dgemm_kernel<TA, TB, 8><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  } else {
    dim3 grid((unsigned)ceil_div(n, BN), (unsigned)ceil_div(m, 64));
    dgemm_kernel<TA, TB, 4><<<grid, 256, 0, ctx->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  }
}
}  // namespace

int dgemm(basq_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool b_lower_tri) {
  const int tri_k = (b_lower_tri && tb) ? 1 : 0;
  if (m <= 0 || n <= 0) return BASQ_OK;
  BASQ_CHECK(k >= 0, BASQ_ERR_INVALID, "dgemm: negative k");
  BASQ_CHECK(ceil_div(m, 64) <= 65535, BASQ_ERR_UNSUPPORTED, "dgemm: m=%d too large for one launch", m);
  if (!ta && !tb) launch<false, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  else if (ta && !tb) launch<true, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  else if (!ta && tb) launch<false, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  else launch<true, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri_k);
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

}  // namespace basq
