// Caratheodory step, register-resident block-pivot version (the fast path; car.cu keeps the general
// global-memory kernel for shapes beyond its limits).  Same algorithm as car.cu:
//   stage 1  Gauss-Jordan with row pivoting turns A [n, S] into [I | T]      (null-space basis)
//   stage 2  ratio-test eliminations of the m = S - n non-basic sets          (BASQ/_rchq.py:146-171)
//
// B200 mapping.  One persistent cooperative kernel, 512 threads per CTA, one CTA per SM.
//   * The tableau never touches shared or global memory between pivots: CTA k keeps a block of B
//     consecutive rows (stage 1) / non-basic columns (stage 2) in REGISTERS, thread t holding the
//     entries of columns (rows) t, t+512, t+1024, ... of each of them.
//   * A block-step: the owner CTA performs its B pivots back to back on its own registers
//     (pivot search = block arg-reduce; the B-1 sibling updates are register FMAs) and streams the
//     B pivot rows to L2 as they are produced; ONE grid barrier; every other CTA replays the B
//     rank-1 updates from L2 (pivot row entries prefetched one pivot ahead, multipliers broadcast
//     through a double-buffered shared array - one __syncthreads per pivot).
//   So n pivots cost ceil(n/B) grid barriers instead of n, and the per-pivot critical path is a few
//   hundred cycles instead of several L2 round trips.
// Limits: S <= 2048 (4 columns per thread), n <= 1024 (2 rows per thread), n, m <= 8 * #SMs.
#include <algorithm>

#include "common.cuh"
#include "gridsync.cuh"

namespace basq {

namespace {

constexpr int NT = 512;
constexpr int CPT = 4;  // tableau columns per thread in stage 1
constexpr int RPT = 2;  // tableau rows per thread in stage 2

struct Car2Dev {
  double* A;
  int n, S;
  int64_t lda;
  double* prow;   // [n][S]   pivot rows as published
  int* pinfo;     // [n]      pivot column of row r, -1 = row skipped (dependent)
  double* pcol;   // [S][n]   pivot columns as published (stage 2)
  double* sinfo;  // [S][2]   alpha, istar
  double* omega;  // [S]
  unsigned* bar;
  int* status;
  double tol;
};

struct VI {
  double v;
  int i;
};

template <bool MAX>
__device__ __forceinline__ VI block_best(VI x, VI* scratch) {
  auto beats = [](const VI& a, const VI& b) {
    if (a.i < 0) return false;
    if (b.i < 0) return true;
    if (MAX ? (a.v > b.v) : (a.v < b.v)) return true;
    return a.v == b.v && a.i < b.i;
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    VI y;
    y.v = __shfl_down_sync(0xffffffffu, x.v, o);
    y.i = __shfl_down_sync(0xffffffffu, x.i, o);
    if (beats(y, x)) x = y;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scratch[warp] = x;
  __syncthreads();
  if (warp == 0) {
    VI y = (lane < NT / 32) ? scratch[lane] : VI{0.0, -1};
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      VI z;
      z.v = __shfl_down_sync(0xffffffffu, y.v, o);
      z.i = __shfl_down_sync(0xffffffffu, y.i, o);
      if (beats(z, y)) y = z;
    }
    if (lane == 0) scratch[16] = y;
  }
  __syncthreads();
  return scratch[16];
}

template <int B>
__global__ void __launch_bounds__(NT, 1) car2_kernel(const Car2Dev a) {
  __shared__ VI scratch[32];
  __shared__ double fbuf[2][8];
  __shared__ double pbcast;
  __shared__ int abort_sh;
  __shared__ int scan_sh[NT];
  __shared__ int tot_sh[CPT + 1];
  extern __shared__ int nbcol[];  // [S] non-basic column list (stage 2)

  const int n = a.n, S = a.S;
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  unsigned target = 0;
  int par = 0;  // parity of the multiplier buffer (advanced once per applied pivot, CTA-uniform)

  for (int c = b * NT + tid; c < S; c += G * NT) a.omega[c] = 0.0;

  // ------------------------------------------------------------------ stage 1: own rows -> registers
  const int G1 = (n + B - 1) / B;
  const int r0 = b * B;
  const int rows_mine = max(0, min(B, n - r0));
  double reg[B][CPT];
  double rscale[B];
  unsigned elig = 0;
#pragma unroll
  for (int j = 0; j < CPT; ++j)
    if (tid + j * NT < S) elig |= 1u << j;
#pragma unroll
  for (int i = 0; i < B; ++i) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + j * NT;
      reg[i][j] = (i < rows_mine && c < S) ? a.A[(int64_t)(r0 + i) * a.lda + c] : 0.0;
    }
  }
#pragma unroll
  for (int i = 0; i < B; ++i) {
    VI m{0.0, -1};
    if (i < rows_mine) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const double v = fabs(reg[i][j]);
        if ((elig >> j & 1u) && (m.i < 0 || v > m.v)) m = VI{v, tid + j * NT};
      }
    }
    rscale[i] = block_best<true>(m, scratch).v;  // (CTA-uniform loop: every thread calls it B times)
  }

  for (int k = 0; k < G1; ++k) {
    const int kr0 = k * B;
    const int krows = min(B, n - kr0);
    if (b == k) {
      // ---- owner: B local pivots
#pragma unroll
      for (int i = 0; i < B; ++i) {
        if (i < krows) {
          VI m{0.0, -1};
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const double v = fabs(reg[i][j]);
            if ((elig >> j & 1u) && (m.i < 0 || v > m.v)) m = VI{v, tid + j * NT};
          }
          m = block_best<true>(m, scratch);
          const bool skip = (m.i < 0) || !(m.v > a.tol * rscale[i]) || !(rscale[i] > 0.0);
          if (skip) {
            if (tid == 0) a.pinfo[kr0 + i] = -1;
          } else {
            const int cs = m.i;
            const bool mine = (cs % NT) == tid;
            const int js = cs / NT;
            if (mine) {
              double p = 0.0;
#pragma unroll
              for (int j = 0; j < CPT; ++j)
                if (j == js) p = reg[i][j];
              pbcast = p;
            }
            __syncthreads();
            const double inv = 1.0 / pbcast;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              const int c = tid + j * NT;
              reg[i][j] = (c == cs) ? 1.0 : reg[i][j] * inv;
              if (c < S) __stcg(&a.prow[(int64_t)(kr0 + i) * S + c], reg[i][j]);
            }
            if (mine) {
              elig &= ~(1u << js);
#pragma unroll
              for (int i2 = 0; i2 < B; ++i2) {
                double f = 0.0;
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                  if (j == js) f = reg[i2][j];
                fbuf[par][i2] = f;
              }
            }
            if (tid == 0) a.pinfo[kr0 + i] = cs;
            __syncthreads();
#pragma unroll
            for (int i2 = 0; i2 < B; ++i2) {
              if (i2 != i && i2 < krows) {
                const double f = fbuf[par][i2];
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                  reg[i2][j] = (tid + j * NT == cs) ? 0.0 : fma(-f, reg[i][j], reg[i2][j]);
              }
            }
            par ^= 1;
          }
        }
      }
    }
    if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;
    if (b != k) {
      // ---- everyone else: replay the B rank-1 updates
      double cur[CPT], nxt[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * NT;
        cur[j] = (c < S) ? __ldcg(&a.prow[(int64_t)kr0 * S + c]) : 0.0;
      }
      for (int i = 0; i < krows; ++i) {
        const int cs = __ldcg(&a.pinfo[kr0 + i]);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * NT;
          nxt[j] = (i + 1 < krows && c < S) ? __ldcg(&a.prow[(int64_t)(kr0 + i + 1) * S + c]) : 0.0;
        }
        if (cs >= 0) {
          const int js = cs / NT;
          if ((cs % NT) == tid) {
            elig &= ~(1u << js);
#pragma unroll
            for (int i2 = 0; i2 < B; ++i2) {
              double f = 0.0;
#pragma unroll
              for (int j = 0; j < CPT; ++j)
                if (j == js) f = reg[i2][j];
              fbuf[par][i2] = f;
            }
          }
          __syncthreads();
#pragma unroll
          for (int i2 = 0; i2 < B; ++i2) {
            if (i2 < rows_mine) {
              const double f = fbuf[par][i2];
#pragma unroll
              for (int j = 0; j < CPT; ++j)
                reg[i2][j] = (tid + j * NT == cs) ? 0.0 : fma(-f, cur[j], reg[i2][j]);
            }
          }
          par ^= 1;
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) cur[j] = nxt[j];
      }
    }
  }

  // ------------------------------------------------------------------ write [I | T] back, list the non-basic sets
#pragma unroll
  for (int i = 0; i < B; ++i)
    if (i < rows_mine) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * NT;
        if (c < S) __stcg(&a.A[(int64_t)(r0 + i) * a.lda + c], reg[i][j]);
      }
    }
  // ascending list of still-eligible (= non-basic) columns: one block scan per column group
  int base = 0;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int bit = (elig >> j) & 1u;
    __syncthreads();
    scan_sh[tid] = bit;
    __syncthreads();
    for (int o = 1; o < NT; o <<= 1) {
      int t = 0;
      if (tid >= o) t = scan_sh[tid - o];
      __syncthreads();
      scan_sh[tid] += t;
      __syncthreads();
    }
    if (bit) nbcol[base + scan_sh[tid] - 1] = tid + j * NT;
    if (tid == NT - 1) tot_sh[j] = scan_sh[tid];
    __syncthreads();
    base += tot_sh[j];
  }
  const int m = base;
  if ((m + B - 1) / B > G) {
    // more non-basic sets than this launch can hold in registers (rank-deficient system):
    // tell the host to redo the step with the general kernel.  Uniform over the grid.
    if (b == 0 && tid == 0) atomicExch(a.status, 3);
    return;
  }
  if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;

  // ------------------------------------------------------------------ stage 2: own non-basic columns -> registers
  const int G2 = (m + B - 1) / B;
  const int c0 = b * B;
  const int cols_mine = max(0, min(B, m - c0));
  double tc[B][RPT];
  double muB[RPT];
  int rowpt[RPT];
#pragma unroll
  for (int jj = 0; jj < RPT; ++jj) {
    const int i = tid + jj * NT;
    rowpt[jj] = (i < n) ? __ldcg(&a.pinfo[i]) : -1;
    muB[jj] = (rowpt[jj] >= 0) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int l = 0; l < B; ++l) {
#pragma unroll
    for (int jj = 0; jj < RPT; ++jj) {
      const int i = tid + jj * NT;
      tc[l][jj] = (l < cols_mine && i < n && rowpt[jj] >= 0) ? __ldcg(&a.A[(int64_t)i * a.lda + nbcol[c0 + l]]) : 0.0;
    }
  }

  // weights move along the null vector of one non-basic set; pv = its tableau column
  auto apply = [&](const double (&pv)[RPT], double alpha, int istar, int entering, int first) {
#pragma unroll
    for (int jj = 0; jj < RPT; ++jj) {
      if (rowpt[jj] >= 0) {
        const double v = fma(alpha, pv[jj], muB[jj]);
        muB[jj] = v > 0.0 ? v : 0.0;
      }
    }
    if (istar >= 0) {
      const int js = istar / NT;
      if ((istar % NT) == tid) {
        double p = 0.0;
#pragma unroll
        for (int jj = 0; jj < RPT; ++jj)
          if (jj == js) p = pv[jj];
#pragma unroll
        for (int l2 = 0; l2 < B; ++l2) {
          double t = 0.0;
#pragma unroll
          for (int jj = 0; jj < RPT; ++jj)
            if (jj == js) t = tc[l2][jj];
          fbuf[par][l2] = t / p;
        }
#pragma unroll
        for (int jj = 0; jj < RPT; ++jj)
          if (jj == js) {
            muB[jj] = 1.0 - alpha;  // the entering set keeps what is left of its unit weight
            rowpt[jj] = entering;
          }
      }
      __syncthreads();
#pragma unroll
      for (int l2 = 0; l2 < B; ++l2) {
        if (l2 >= first && l2 < cols_mine) {
          const double t = fbuf[par][l2];
#pragma unroll
          for (int jj = 0; jj < RPT; ++jj)
            tc[l2][jj] = (tid + jj * NT == istar) ? t : fma(-pv[jj], t, tc[l2][jj]);
        }
      }
      par ^= 1;
    }
  };

  for (int k = 0; k < G2; ++k) {
    const int kc0 = k * B;
    const int kcols = min(B, m - kc0);
    if (b == k) {
#pragma unroll
      for (int l = 0; l < B; ++l) {
        if (l < kcols) {
          VI best{0.0, -1};
#pragma unroll
          for (int jj = 0; jj < RPT; ++jj) {
            const double t = tc[l][jj];
            if (rowpt[jj] >= 0 && t < 0.0) {
              const double ratio = muB[jj] / (-t);
              if (best.i < 0 || ratio < best.v) best = VI{ratio, tid + jj * NT};
            }
          }
          best = block_best<false>(best, scratch);
          // the non-basic set itself has +1 in its null vector and weight 1: ratio 1
          const bool self = (best.i < 0) || !(best.v < 1.0);
          const double alpha = self ? 1.0 : best.v;
          const int istar = self ? -1 : best.i;
          const int jn = kc0 + l;
#pragma unroll
          for (int jj = 0; jj < RPT; ++jj) {
            const int i = tid + jj * NT;
            if (i < n) __stcg(&a.pcol[(int64_t)jn * n + i], tc[l][jj]);
          }
          if (tid == 0) {
            __stcg(&a.sinfo[2 * jn + 0], alpha);
            __stcg(&a.sinfo[2 * jn + 1], (double)istar);
          }
          double pv[RPT];
#pragma unroll
          for (int jj = 0; jj < RPT; ++jj) pv[jj] = tc[l][jj];
          apply(pv, alpha, istar, nbcol[jn], l + 1);
        }
      }
    }
    if (grid_barrier(a.bar, target, a.status, &abort_sh)) return;
    if (b != k) {
      for (int l = 0; l < kcols; ++l) {
        const int jn = kc0 + l;
        const double alpha = __ldcg(&a.sinfo[2 * jn + 0]);
        const int istar = (int)__ldcg(&a.sinfo[2 * jn + 1]);
        double pv[RPT];
#pragma unroll
        for (int jj = 0; jj < RPT; ++jj) {
          const int i = tid + jj * NT;
          pv[jj] = (i < n) ? __ldcg(&a.pcol[(int64_t)jn * n + i]) : 0.0;
        }
        // CTAs whose columns are already eliminated (b < k) only track the weights
        apply(pv, alpha, istar, nbcol[jn], b > k ? 0 : B);
      }
    }
  }
  if (b == 0) {
#pragma unroll
    for (int jj = 0; jj < RPT; ++jj)
      if (rowpt[jj] >= 0 && muB[jj] > 0.0) a.omega[rowpt[jj]] = muB[jj];
  }
}

template <int B>
int launch_car2(basq_ctx* ctx, const Car2Dev& d, int grid, size_t smem) {
  BASQ_CUDA(cudaFuncSetAttribute(car2_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&d};
  BASQ_CUDA(cudaLaunchCooperativeKernel((const void*)car2_kernel<B>, dim3(grid), dim3(NT), args, smem, ctx->stream));
  return BASQ_OK;
}

}  // namespace

bool caratheodory_fast_supported(const basq_ctx* ctx, int n, int S) {
  const int m = S - n;
  return S <= CPT * NT && n <= RPT * NT && n <= 8 * ctx->num_sms && m <= 8 * ctx->num_sms && m >= 1;
}

// status_out: 0 ok, 3 = more non-basic sets than fit (A holds the stage-1 result; caller must redo)
int caratheodory_fast(basq_ctx* ctx, double* A, int n, int S, int lda, double* omega_out, int* status_out) {
  const int m_max = S;  // non-basic sets: S - rank; S - n for a full-rank system
  int B = 1;
  while (B < 8 && ((n + B - 1) / B > ctx->num_sms || (m_max + B - 1) / B > ctx->num_sms)) B *= 2;
  const int grid = std::min(ctx->num_sms, std::max((n + B - 1) / B, (m_max + B - 1) / B));
  DevBuf ws;
  const size_t sz_prow = sizeof(double) * (size_t)n * S, sz_pcol = sizeof(double) * (size_t)S * n,
               sz_sinfo = sizeof(double) * 2 * S, sz_int = sizeof(int) * ((size_t)n + 8);
  BASQ_TRY(ws.alloc(ctx, 512 + sz_prow + sz_pcol + sz_sinfo + sz_int + 64));
  unsigned char* w = ws.as<unsigned char>();
  Car2Dev d;
  d.A = A; d.n = n; d.S = S; d.lda = lda;
  d.bar = reinterpret_cast<unsigned*>(w);            // 256 B: arrival counter + flag line
  d.status = reinterpret_cast<int*>(w + 256);
  w += 512;
  d.prow = reinterpret_cast<double*>(w); w += sz_prow;
  d.pcol = reinterpret_cast<double*>(w); w += sz_pcol;
  d.sinfo = reinterpret_cast<double*>(w); w += sz_sinfo;
  d.pinfo = reinterpret_cast<int*>(w);
  d.omega = omega_out;
  d.tol = 1e-13;
  BASQ_CUDA(cudaMemsetAsync(ws.p, 0, 512, ctx->stream));
  const size_t smem = sizeof(int) * (size_t)S;
  switch (B) {
    case 1: BASQ_TRY(launch_car2<1>(ctx, d, grid, smem)); break;
    case 2: BASQ_TRY(launch_car2<2>(ctx, d, grid, smem)); break;
    case 4: BASQ_TRY(launch_car2<4>(ctx, d, grid, smem)); break;
    default: BASQ_TRY(launch_car2<8>(ctx, d, grid, smem)); break;
  }
  ctx->launches++;
  int status = 0;
  BASQ_CUDA(cudaMemcpyAsync(&status, d.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BASQ_CUDA(cudaStreamSynchronize(ctx->stream));
  *status_out = status;
  BASQ_CHECK(status == 0 || status == 3, BASQ_ERR_NUMERIC, "caratheodory: grid barrier watchdog fired (status %d)",
             status);
  return BASQ_OK;
}

}  // namespace basq
