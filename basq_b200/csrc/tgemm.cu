// Dense fp32-accuracy GEMM on the 5th-generation tensor cores (3xTF32 split, fp32 accumulate in TMEM):
//
//   D[I, J] = A[I, Kd] * B[J, Kd]^T
//
// used for the Nystrom subspace iteration (K(Z,Z) Y, reference torch.svd_lowrank inside
// ker_svd_sparsify, BASQ/_rchq.py:28-31) and the GP posterior-variance contraction
// (k^T W k per candidate, BASQ/_gp.py:213-230) when the kernel is evaluated in fp32.
//
// Operands are stored "blocked + split" (BlkOperand): two fp32 arrays hi / lo with
// x = hi + lo, hi = tf32(x), lo = tf32(x - hi), laid out [row tile of 128][K chunk of 4][128 rows][4]
// so that one (row tile, 32-wide K block) is a contiguous 16 KB piece that a single bulk copy
// (cp.async.bulk, UBLKCP) drops into shared memory in exactly the K-major no-swizzle form the
// tcgen05 shared-memory descriptor expects (8-row core matrices contiguous, SBO = 128 B,
// LBO = 128 rows * 16 B).  Three MMAs per K step (lo*hi, hi*lo, hi*hi; kind::tf32, M = 128, N = 128)
// reproduce the fp32 product to ~2^-21.
//
// CTA = 6 warps, persistent, one CTA per SM; tile 128 x 256 (two 128-column accumulators):
//   warp 0    producer: six 16 KB bulk copies per K block (A hi/lo, B0 hi/lo, B1 hi/lo), 2-stage ring
//   warp 1    TMEM allocation + tcgen05.mma issue (24 MMAs per K block, one lane)
//   warps 2-5 epilogue: tcgen05.ld -> fp64 store (plain or transposed); double-buffered accumulators
//             (2 x 256 TMEM columns) so the epilogue of one tile overlaps the main loop of the next
#include "common.cuh"
#include "setsum_mma.cuh"  // mma:: helpers (mbarrier, bulk copy, tcgen05 wrappers, smem descriptor)
#include "tgemm.cuh"

namespace basq {

namespace {

constexpr int TG_THREADS = 192;
constexpr int TG_NSTAGE = 2;
constexpr int TG_PIECE = 16384;                   // bytes of one (row tile, K block) piece
constexpr int TG_STAGE_BYTES = 6 * TG_PIECE;      // A hi, A lo, B0 hi, B0 lo, B1 hi, B1 lo
constexpr int TG_SMEM = TG_NSTAGE * TG_STAGE_BYTES + 256;

struct TgDev {
  const float *Ahi, *Alo, *Bhi, *Blo;
  int I, J;        // logical output size
  int RT_A, RT_B;  // row tiles
  int KC;          // K chunks (of 4) = 8 * K blocks
  double* out;
  int64_t ldo;
  int transposed;  // 0: out[i * ldo + j] ; 1: out[j * ldo + i]
  double alpha;
  int accumulate;  // out += alpha * A B^T instead of out = ...
};

__global__ void __launch_bounds__(TG_THREADS, 1) tgemm_kernel(const TgDev a) {
  extern __shared__ __align__(1024) unsigned char smem_tg[];
  unsigned char* const smem = smem_tg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TG_NSTAGE * TG_STAGE_BYTES);
  uint64_t* s_full = bars;                    // [NSTAGE]
  uint64_t* s_empty = bars + TG_NSTAGE;       // [NSTAGE]
  uint64_t* t_full = bars + 2 * TG_NSTAGE;    // [2]
  uint64_t* t_empty = t_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < TG_NSTAGE; ++s) {
      mma::mbar_init(&s_full[s], 1);
      mma::mbar_init(&s_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mma::mbar_init(&t_full[b], 1);
      mma::mbar_init(&t_empty[b], 4);
    }
    mma::fence_barrier_init();
  }
  if (warp == 1) mma::tmem_alloc(tmem_slot, 512);
  mma::tc_fence_before();
  __syncthreads();
  mma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tb = (a.RT_B + 1) / 2;
  const int n_items = a.RT_A * n_tb;
  const int nkb = a.KC / 8;
  const size_t tile_floats = (size_t)a.KC * 128 * 4;  // floats per row tile

  if (warp == 0) {
    // ======================================================================== producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ta = item % a.RT_A, tb = item / a.RT_A;
        const int tb0 = 2 * tb, tb1 = min(2 * tb + 1, a.RT_B - 1);  // an odd tail re-reads the last tile (discarded)
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % TG_NSTAGE;
          mma::mbar_wait(&s_empty[s], ((it / TG_NSTAGE) & 1u) ^ 1u);
          unsigned char* st = smem + (size_t)s * TG_STAGE_BYTES;
          mma::mbar_expect_tx(&s_full[s], TG_STAGE_BYTES);
          const size_t koff = (size_t)kb * 8 * 128 * 4;
          mma::bulk_g2s(st + 0 * TG_PIECE, a.Ahi + ta * tile_floats + koff, TG_PIECE, &s_full[s]);
          mma::bulk_g2s(st + 1 * TG_PIECE, a.Alo + ta * tile_floats + koff, TG_PIECE, &s_full[s]);
          mma::bulk_g2s(st + 2 * TG_PIECE, a.Bhi + tb0 * tile_floats + koff, TG_PIECE, &s_full[s]);
          mma::bulk_g2s(st + 3 * TG_PIECE, a.Blo + tb0 * tile_floats + koff, TG_PIECE, &s_full[s]);
          mma::bulk_g2s(st + 4 * TG_PIECE, a.Bhi + tb1 * tile_floats + koff, TG_PIECE, &s_full[s]);
          mma::bulk_g2s(st + 5 * TG_PIECE, a.Blo + tb1 * tile_floats + koff, TG_PIECE, &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================== MMA issuer
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t it = 0, tc = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tc) {
      const uint32_t buf = tc & 1u;
      mma::mbar_wait(&t_empty[buf], ((tc >> 1) & 1u) ^ 1u);
      mma::tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % TG_NSTAGE;
        mma::mbar_wait(&s_full[s], (it / TG_NSTAGE) & 1u);
        mma::tc_fence_after();
        if (lane == 0) {
          const uint32_t st = mma::smem_u32(smem + (size_t)s * TG_STAGE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t d = tmem_base + buf * 256 + h * 128;
            const uint32_t ahi = st, alo = st + TG_PIECE;
            const uint32_t bhi = st + (2 + 2 * h) * TG_PIECE, blo = bhi + TG_PIECE;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              const uint32_t aa = (p == 0) ? alo : ahi;   // lo*hi, hi*lo, hi*hi
              const uint32_t bb = (p == 1) ? blo : bhi;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = mma::smem_desc(aa + ks * 2 * (128 * 16), 128 * 16, 128);
                const uint64_t bd = mma::smem_desc(bb + ks * 2 * (128 * 16), 128 * 16, 128);
                mma::umma_tf32(d, ad, bd, IDESC, (kb > 0 || p > 0 || ks > 0) ? 1u : 0u);
              }
            }
          }
          mma::umma_commit(&s_empty[s]);
          if (kb == nkb - 1) mma::umma_commit(&t_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================================================== epilogue
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t tc = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tc) {
      const int ta = item % a.RT_A, tb = item / a.RT_A;
      const uint32_t buf = tc & 1u;
      mma::mbar_wait(&t_full[buf], (tc >> 1) & 1u);
      mma::tc_fence_after();
      const int i = ta * 128 + row;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256;
      const int ncols = min(256, a.J - tb * 256);  // an odd tail's second half lies beyond J
#pragma unroll 1
      for (int cb = 0; cb < ncols; cb += 32) {
        uint32_t v[32];
        mma::tmem_ld32(taddr + cb, v);
        mma::tmem_ld_wait();
        if (i < a.I) {
          const int j0 = tb * 256 + cb;
          if (a.transposed) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j0 + c < a.J) {
                double* dst = a.out + (int64_t)(j0 + c) * a.ldo + i;
                const double val = a.alpha * (double)__uint_as_float(v[c]);
                *dst = a.accumulate ? *dst + val : val;
              }
          } else {
            double* dst = a.out + (int64_t)i * a.ldo + j0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j0 + c < a.J) {
                const double val = a.alpha * (double)__uint_as_float(v[c]);
                dst[c] = a.accumulate ? dst[c] + val : val;
              }
          }
        }
      }
      mma::tc_fence_before();
      __syncwarp();
      if (lane == 0) mma::mbar_arrive(&t_empty[buf]);
    }
  }

  mma::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    mma::tc_fence_after();
    mma::tmem_dealloc(tmem_base, 512);
  }
}

// one thread per (row tile, chunk, row): 4 consecutive K elements of one operand row
template <bool TRANSPOSED>
__global__ void blk_from_f64_kernel(const double* __restrict__ src, int64_t ld, int rows, int kdim, int RT, int KC,
                                    float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)RT * KC * 128;
  if (t >= total) return;
  const int r = (int)(t % 128);
  const int kc = (int)((t / 128) % KC);
  const int rt = (int)(t / (128 * (int64_t)KC));
  const int row = rt * 128 + r;
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kc * 4 + e;
    double x = 0.0;
    if (row < rows && k < kdim) x = TRANSPOSED ? src[(int64_t)k * ld + row] : src[(int64_t)row * ld + k];
    const float xf = (float)x;
    h[e] = mma::tf32_rna(xf);
    // the residual is taken against the fp64 value: hi + lo carries ~22 bits of x
    l[e] = mma::tf32_rna((float)(x - (double)h[e]));
  }
  reinterpret_cast<float4*>(hi)[t] = make_float4(h[0], h[1], h[2], h[3]);
  reinterpret_cast<float4*>(lo)[t] = make_float4(l[0], l[1], l[2], l[3]);
}

}  // namespace

int BlkOperand::alloc(basq_ctx* ctx, int rows_, int kdim_) {
  rows = rows_;
  kdim = kdim_;
  RT = ceil_div(rows_, 128);
  KC = ceil_div(kdim_, 32) * 8;
  const size_t bytes = sizeof(float) * (size_t)RT * KC * 128 * 4;
  BASQ_TRY(hi.alloc(ctx, bytes));
  BASQ_TRY(lo.alloc(ctx, bytes));
  return BASQ_OK;
}

int blk_from_f64(basq_ctx* ctx, const double* src, int64_t ld, bool transposed, BlkOperand* op) {
  const int64_t total = (int64_t)op->RT * op->KC * 128;
  const unsigned grid = (unsigned)ceil_div64(total, 256);
  if (transposed)
    blk_from_f64_kernel<true><<<grid, 256, 0, ctx->stream>>>(src, ld, op->rows, op->kdim, op->RT, op->KC,
                                                             op->hi.as<float>(), op->lo.as<float>());
  else
    blk_from_f64_kernel<false><<<grid, 256, 0, ctx->stream>>>(src, ld, op->rows, op->kdim, op->RT, op->KC,
                                                              op->hi.as<float>(), op->lo.as<float>());
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

int tgemm(basq_ctx* ctx, const BlkOperand& A, const BlkOperand& B, double alpha, double* out, int64_t ldo,
          bool transposed, bool accumulate) {
  BASQ_CHECK(A.kdim == B.kdim && A.KC == B.KC, BASQ_ERR_INVALID, "tgemm: inner dimensions differ (%d vs %d)", A.kdim,
             B.kdim);
  BASQ_CHECK((size_t)TG_SMEM <= ctx->smem_optin, BASQ_ERR_UNSUPPORTED, "tgemm: needs %d B shared memory", TG_SMEM);
  if (A.rows <= 0 || B.rows <= 0) return BASQ_OK;
  TgDev d;
  d.Ahi = A.hi.as<float>(); d.Alo = A.lo.as<float>();
  d.Bhi = B.hi.as<float>(); d.Blo = B.lo.as<float>();
  d.I = A.rows; d.J = B.rows;
  d.RT_A = A.RT; d.RT_B = B.RT;
  d.KC = A.KC;
  d.out = out; d.ldo = ldo;
  d.transposed = transposed ? 1 : 0;
  d.alpha = alpha;
  d.accumulate = accumulate ? 1 : 0;
  const int n_items = A.RT * ((B.RT + 1) / 2);
  BASQ_CUDA(cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
  const int grid = std::min(ctx->num_sms, n_items);
  tgemm_kernel<<<grid, TG_THREADS, TG_SMEM, ctx->stream>>>(d);
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

}  // namespace basq
