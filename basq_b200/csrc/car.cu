// Caratheodory step of kernel recombination: reduce S weighted barycentres in R^(n-1) to <= n,
// preserving mass and barycentre (reference Tchernychova_Lyons_CAR, BASQ/_rchq.py:133-175).
//
// The reference builds an orthonormal null-space basis of [1|X]^T by a full SVD and performs
// S - n ratio-test eliminations on it.  Here the same elimination is carried out in tableau form,
// which needs no SVD and half the storage:
//   stage 1  Gauss-Jordan with row pivoting over the set columns turns A [n, S] into [I | T]:
//            n basic columns (sets) and the tableau T [n, m] of the m = S - n non-basic ones.  Each
//            non-basic column j gives the null vector (+1 at j, -T[:, j] on the basic sets) - the
//            same null space the reference gets from the SVD.
//   stage 2  for every non-basic column: ratio test along that null vector (reference :148-152),
//            move the weights until the first one hits zero (:158-159); if a basic set died, pivot
//            it out of the basis (the reference's rank-1 update :165-171 restricted to the tableau).
// Weights are relative (omega, starting at 1 for every set): A omega = A 1 is preserved, so the
// caller rescales the points of set j by omega_j - no division by set masses anywhere.
//
// B200 mapping: ONE persistent cooperative kernel, one CTA per SM.  Stage 1 distributes the
// tableau by rows, stage 2 by columns, so every pivot needs exactly one grid-wide barrier: the
// owner of the next pivot row/column updates it first, runs the pivot search and publishes the
// (scaled) pivot row / column through L2 while the other CTAs are still applying the current
// rank-1 update.  All arithmetic is fp64 (the 1e-8 moment tolerance needs it).
#include <cooperative_groups.h>

#include "common.cuh"
#include "gridsync.cuh"

namespace basq {

namespace {

constexpr int CAR_THREADS = 512;

struct CarDev {
  double* A;
  int n, S;
  int64_t lda;
  double* prow;      // [2][S]
  int* pinfo;        // [2]
  double* rowscale;  // [n]
  int* colbasis;     // [S]
  int* rowpoint;     // [n]
  double* Tc;        // [S][n]
  double* pcol;      // [2][n]
  double* sinfo;     // [2][2]
  double* omega;     // [S]
  unsigned* bar;
  int* status;
  double tol;
};

struct ValIdx {
  double v;
  int i;
};

// block-wide arg-best; better(a, b) is a strict "a beats b"; ties resolved to the smaller index.
template <bool MAX>
__device__ __forceinline__ ValIdx block_arg_best(ValIdx x, ValIdx* scratch) {
  auto beats = [](const ValIdx& a, const ValIdx& b) {
    if (a.i < 0) return false;
    if (b.i < 0) return true;
    if (MAX ? (a.v > b.v) : (a.v < b.v)) return true;
    if (a.v == b.v && a.i < b.i) return true;
    return false;
  };
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_down_sync(0xffffffffu, x.v, o);
    y.i = __shfl_down_sync(0xffffffffu, x.i, o);
    if (beats(y, x)) x = y;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = x;
  __syncthreads();
  if (warp == 0) {
    ValIdx y = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : ValIdx{0.0, -1};
    for (int o = 16; o > 0; o >>= 1) {
      ValIdx z;
      z.v = __shfl_down_sync(0xffffffffu, y.v, o);
      z.i = __shfl_down_sync(0xffffffffu, y.i, o);
      if (beats(z, y)) y = z;
    }
    if (lane == 0) scratch[0] = y;
  }
  __syncthreads();
  const ValIdx r = scratch[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(CAR_THREADS, 1) car_kernel(const CarDev a) {
  extern __shared__ __align__(16) unsigned char car_smem[];
  const int n = a.n, S = a.S;
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  // shared layout (byte offsets kept 16-aligned)
  size_t off = 0;
  auto carve = [&](size_t bytes) { unsigned char* p = car_smem + off; off += (bytes + 15) & ~(size_t)15; return p; };
  double* rowc = reinterpret_cast<double*>(carve(sizeof(double) * S));  // cached pivot row (stage 1)
  double* pc = reinterpret_cast<double*>(carve(sizeof(double) * n));    // cached pivot column (stage 2)
  double* muB = reinterpret_cast<double*>(carve(sizeof(double) * n));
  int* rowpt = reinterpret_cast<int*>(carve(sizeof(int) * n));
  int* nbcol = reinterpret_cast<int*>(carve(sizeof(int) * S));
  int* itmp = reinterpret_cast<int*>(carve(sizeof(int) * S));
  __shared__ ValIdx scratch[32];
  __shared__ int tscan[CAR_THREADS];
  __shared__ double bcast;
  __shared__ int ibcast;
  __shared__ int abort_sh;

  unsigned target = 0;

  // ---------------------------------------------------------------- init
  for (int r = b; r < n; r += G) {
    ValIdx best{0.0, -1};
    for (int c = tid; c < S; c += NT) {
      const double v = fabs(a.A[(int64_t)r * a.lda + c]);
      if (best.i < 0 || v > best.v) best = ValIdx{v, c};
    }
    best = block_arg_best<true>(best, scratch);
    if (tid == 0) {
      a.rowscale[r] = best.v;
      a.rowpoint[r] = -1;
    }
  }
  for (int c = b * NT + tid; c < S; c += G * NT) {
    a.colbasis[c] = -1;
    a.omega[c] = 0.0;
  }
  if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;

  // ---------------------------------------------------------------- stage 1 helpers
  // pivot search on row r (owned), scale it and publish it
  auto publish_row = [&](int r) {
    double* row = a.A + (int64_t)r * a.lda;
    ValIdx best{0.0, -1};
    for (int c = tid; c < S; c += NT) {
      if (__ldcg(&a.colbasis[c]) >= 0) continue;
      const double v = fabs(row[c]);
      if (best.i < 0 || v > best.v) best = ValIdx{v, c};
    }
    best = block_arg_best<true>(best, scratch);
    const double scale = a.rowscale[r];
    const bool skip = (best.i < 0) || !(best.v > a.tol * scale) || !(scale > 0.0);
    if (skip) {
      if (tid == 0) a.pinfo[r & 1] = -1;
      return;
    }
    const int c = best.i;
    const double p = row[c];
    __syncthreads();
    double* pub = a.prow + (int64_t)(r & 1) * S;
    for (int cc = tid; cc < S; cc += NT) {
      const double v = (cc == c) ? 1.0 : row[cc] / p;
      row[cc] = v;
      __stcg(&pub[cc], v);
    }
    if (tid == 0) {
      a.colbasis[c] = r;
      a.rowpoint[r] = c;
      a.pinfo[r & 1] = c;
    }
  };
  // row r2 -= A[r2][c] * pivot row (cached in rowc)
  auto update_row = [&](int r2, int c) {
    double* row = a.A + (int64_t)r2 * a.lda;
    if (tid == 0) bcast = row[c];
    __syncthreads();
    const double f = bcast;
    if (f != 0.0) {
      for (int cc = tid; cc < S; cc += NT) row[cc] = (cc == c) ? 0.0 : fma(-f, rowc[cc], row[cc]);
    }
    __syncthreads();
  };

  // ---------------------------------------------------------------- stage 1: A -> [I | T]
  if (b == 0) publish_row(0);
  for (int r = 0; r < n; ++r) {
    if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;
    const int c = __ldcg(&a.pinfo[r & 1]);
    if (c >= 0) {
      const double* pub = a.prow + (int64_t)(r & 1) * S;
      for (int cc = tid; cc < S; cc += NT) rowc[cc] = __ldcg(&pub[cc]);
    }
    __syncthreads();
    const int rn = r + 1;
    if (rn < n && (rn % G) == b) {
      if (c >= 0) update_row(rn, c);
      publish_row(rn);
    }
    if (c >= 0) {
      for (int r2 = b; r2 < n; r2 += G) {
        if (r2 == r || r2 == rn) continue;
        update_row(r2, c);
      }
    }
  }
  if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;

  // ---------------------------------------------------------------- stage 2 setup
  // non-basic column list (ascending), replicated per CTA
  for (int c = tid; c < S; c += NT) itmp[c] = (__ldcg(&a.colbasis[c]) < 0) ? 1 : 0;
  __syncthreads();
  // exclusive scan of itmp into nbcol-positions: chunked per thread
  {
    const int per = (S + NT - 1) / NT;
    const int lo = tid * per, hi = min(S, lo + per);
    int cnt = 0;
    for (int c = lo; c < hi; ++c) cnt += itmp[c];
    __syncthreads();
    tscan[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < NT; o <<= 1) {
      int t = 0;
      if (tid >= o) t = tscan[tid - o];
      __syncthreads();
      tscan[tid] += t;
      __syncthreads();
    }
    int pos = tscan[tid] - cnt;
    if (tid == NT - 1) ibcast = tscan[tid];
    for (int c = lo; c < hi; ++c)
      if (itmp[c]) nbcol[pos++] = c;
    __syncthreads();
  }
  const int m = ibcast;
  for (int i = tid; i < n; i += NT) {
    const int rp = __ldcg(&a.rowpoint[i]);
    rowpt[i] = rp;
    muB[i] = (rp >= 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  // own columns -> column-major storage Tc[jn][:]
  for (int jn = b; jn < m; jn += G) {
    const int c = nbcol[jn];
    double* col = a.Tc + (int64_t)jn * n;
    for (int i = tid; i < n; i += NT) col[i] = (rowpt[i] >= 0) ? __ldcg(&a.A[(int64_t)i * a.lda + c]) : 0.0;
  }
  __syncthreads();

  // ratio test on own column jn (up to date) and publish it
  auto publish_col = [&](int jn) {
    const double* col = a.Tc + (int64_t)jn * n;
    ValIdx best{0.0, -1};
    for (int i = tid; i < n; i += NT) {
      const double t = col[i];
      if (rowpt[i] >= 0 && t < 0.0) {
        const double ratio = muB[i] / (-t);
        if (best.i < 0 || ratio < best.v) best = ValIdx{ratio, i};
      }
    }
    best = block_arg_best<false>(best, scratch);
    // the non-basic set itself carries weight 1 and has +1 in its null vector: ratio 1
    const bool self = (best.i < 0) || !(best.v < 1.0);
    double* pub = a.pcol + (int64_t)(jn & 1) * n;
    for (int i = tid; i < n; i += NT) __stcg(&pub[i], col[i]);
    if (tid == 0) {
      __stcg(&a.sinfo[(jn & 1) * 2 + 0], self ? 1.0 : best.v);
      __stcg(&a.sinfo[(jn & 1) * 2 + 1], self ? -1.0 : (double)best.i);
    }
  };
  // pivot update of own column jc with the cached pivot column pc, pivot row istar
  auto update_col = [&](int jc, int istar) {
    double* col = a.Tc + (int64_t)jc * n;
    if (tid == 0) bcast = col[istar] / pc[istar];
    __syncthreads();
    const double t = bcast;
    if (t != 0.0) {
      for (int i = tid; i < n; i += NT) col[i] = (i == istar) ? t : fma(-pc[i], t, col[i]);
    } else if (tid == 0) {
      col[istar] = 0.0;
    }
    __syncthreads();
  };

  // ---------------------------------------------------------------- stage 2: eliminate m sets
  if (b == 0 && m > 0) publish_col(0);
  for (int jn = 0; jn < m; ++jn) {
    if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;
    const double alpha = __ldcg(&a.sinfo[(jn & 1) * 2 + 0]);
    const int istar = (int)__ldcg(&a.sinfo[(jn & 1) * 2 + 1]);
    const double* pub = a.pcol + (int64_t)(jn & 1) * n;
    for (int i = tid; i < n; i += NT) pc[i] = __ldcg(&pub[i]);
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      if (rowpt[i] >= 0) {
        const double v = fma(alpha, pc[i], muB[i]);
        muB[i] = v > 0.0 ? v : 0.0;
      }
    }
    __syncthreads();
    if (istar >= 0 && tid == 0) {
      muB[istar] = 1.0 - alpha;          // the entering set keeps what is left of its unit weight
      rowpt[istar] = nbcol[jn];
    }
    __syncthreads();
    const int jnn = jn + 1;
    if (jnn < m && (jnn % G) == b) {
      if (istar >= 0) update_col(jnn, istar);
      publish_col(jnn);
    }
    if (istar >= 0) {
      // first own column index > jn
      int j0 = jn + 1 + ((b - (jn + 1)) % G + G) % G;
      for (int jc = j0; jc < m; jc += G) {
        if (jc == jnn) continue;
        update_col(jc, istar);
      }
    }
  }
  __syncthreads();
  if (b == 0) {
    for (int i = tid; i < n; i += NT)
      if (rowpt[i] >= 0 && muB[i] > 0.0) a.omega[rowpt[i]] = muB[i];
  }
}

}  // namespace

bool caratheodory_fast_supported(const basq_ctx* ctx, int n, int S);
int caratheodory_fast(basq_ctx* ctx, double* A, int n, int S, int lda, double* omega_out, int* status_out);

int caratheodory(basq_ctx* ctx, double* A, int n, int S, int lda, double* omega_out) {
  PhaseTimer timer(ctx, PH_CAR);
  BASQ_CHECK(n >= 1 && S >= 1 && lda >= S, BASQ_ERR_INVALID, "caratheodory: bad shape n=%d S=%d lda=%d", n, S, lda);
  if (S > n && !ctx->force_general_car && caratheodory_fast_supported(ctx, n, S)) {
    // register-resident block-pivot kernel (car2.cu).  It can only run out of room when the
    // system is rank deficient AND large; keep a copy of A for that case.
    DevBuf backup;
    const bool risky = (S + 7) / 8 > ctx->num_sms;
    if (risky) {
      BASQ_TRY(backup.alloc(ctx, sizeof(double) * (size_t)n * lda));
      BASQ_CUDA(cudaMemcpyAsync(backup.p, A, sizeof(double) * (size_t)n * lda, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    int st = 0;
    BASQ_TRY(caratheodory_fast(ctx, A, n, S, lda, omega_out, &st));
    if (st == 0) return BASQ_OK;
    BASQ_CHECK(risky, BASQ_ERR_NUMERIC, "caratheodory: fast kernel overflowed without a backup");
    BASQ_CUDA(cudaMemcpyAsync(A, backup.p, sizeof(double) * (size_t)n * lda, cudaMemcpyDeviceToDevice, ctx->stream));
    BASQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (S <= n) {
    // nothing to eliminate: every set keeps its weight
    std::vector<double> ones((size_t)S, 1.0);
    BASQ_CUDA(cudaMemcpyAsync(omega_out, ones.data(), sizeof(double) * S, cudaMemcpyHostToDevice, ctx->stream));
    BASQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return BASQ_OK;
  }
  auto r16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t smem = r16(sizeof(double) * S) + 2 * r16(sizeof(double) * n) + r16(sizeof(int) * n) +
                      2 * r16(sizeof(int) * S);
  BASQ_CHECK(smem <= ctx->smem_optin, BASQ_ERR_UNSUPPORTED,
             "caratheodory: n=%d S=%d needs %zu B shared memory (limit %zu)", n, S, smem, ctx->smem_optin);
  BASQ_CUDA(cudaFuncSetAttribute(car_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  BASQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, car_kernel, CAR_THREADS, smem));
  BASQ_CHECK(per_sm >= 1, BASQ_ERR_UNSUPPORTED, "caratheodory: kernel does not fit on an SM");
  int G = (n + 1) / 2;
  if (G > ctx->num_sms) G = ctx->num_sms;
  if (G < 1) G = 1;

  DevBuf ws;
  const size_t sz_prow = sizeof(double) * 2 * S, sz_rows = sizeof(double) * n, sz_tc = sizeof(double) * (size_t)S * n,
               sz_pcol = sizeof(double) * 2 * n, sz_sinfo = sizeof(double) * 4;
  const size_t sz_int = sizeof(int) * (2 + (size_t)S + n + 2 + 4);
  BASQ_TRY(ws.alloc(ctx, 512 + sz_prow + sz_rows + sz_tc + sz_pcol + sz_sinfo + sz_int + 256));
  unsigned char* w = ws.as<unsigned char>();
  CarDev d;
  d.A = A; d.n = n; d.S = S; d.lda = lda;
  d.bar = reinterpret_cast<unsigned*>(w);            // 256 B: arrival counter + flag line
  d.status = reinterpret_cast<int*>(w + 256);
  w += 512;
  d.prow = reinterpret_cast<double*>(w); w += sz_prow;
  d.rowscale = reinterpret_cast<double*>(w); w += sz_rows;
  d.Tc = reinterpret_cast<double*>(w); w += sz_tc;
  d.pcol = reinterpret_cast<double*>(w); w += sz_pcol;
  d.sinfo = reinterpret_cast<double*>(w); w += sz_sinfo;
  d.pinfo = reinterpret_cast<int*>(w); w += sizeof(int) * 2;
  d.colbasis = reinterpret_cast<int*>(w); w += sizeof(int) * S;
  d.rowpoint = reinterpret_cast<int*>(w); w += sizeof(int) * n;
  d.omega = omega_out;
  d.tol = 1e-13;
  BASQ_CUDA(cudaMemsetAsync(ws.p, 0, 512, ctx->stream));
  BASQ_CUDA(cudaMemsetAsync(d.pinfo, 0, sz_int, ctx->stream));
  void* args[] = {(void*)&d};
  BASQ_CUDA(cudaLaunchCooperativeKernel((const void*)car_kernel, dim3(G), dim3(CAR_THREADS), args, smem, ctx->stream));
  ctx->launches++;
  int status = 0;
  BASQ_CUDA(cudaMemcpyAsync(&status, d.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BASQ_CUDA(cudaStreamSynchronize(ctx->stream));
  BASQ_CHECK(status == 0, BASQ_ERR_NUMERIC, "caratheodory: grid barrier watchdog fired (status %d)", status);
  return BASQ_OK;
}

}  // namespace basq
