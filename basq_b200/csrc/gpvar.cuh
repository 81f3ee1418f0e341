// Fused GP posterior variance over candidates on the tensor cores (fp32 inputs).
//
//   var(x) = sigma_f^2 + sigma_n^2 - k_x^T W k_x,   k_x = k(Xobs, x),  W = (K_XX + sigma_n^2 I)^-1
//   (predict() of BASQ/_gp.py:213-230 with the exact variance; consumers: calc_weights,
//    BASQ/_sampler.py:194-217; wsabim_predict, BASQ/_wsabi.py:265-277; gspace_predict,
//    SOBER/BASQ/_scale_mmlt.py:211-223; lfi, SOBER/_pi.py:121-139)
//
// k^T W k = sum_o k_o (T k)_o with T = tril(W + W^T) (diagonal taken once): the contraction T k over a
// tile of 128 candidates is a GEMM whose result never leaves the SM:
//   D[p, o] = sum_{o' <= o} kx[p, o'] T[o, o']      tcgen05.mma kind::f16, M = 128 candidates (TMEM
//                                                   lanes), N = 256 observations, fp16 hi / lo split
//                                                   operands, three products (fp32 accuracy), the K
//                                                   loop stops at the diagonal block (half the flop);
//   epilogue: thread = candidate; for every observation column k(xobs_o, x_p) is recomputed on the
//             FMA / MUFU pipes (the candidate's coordinates live in registers, the observation's arrive
//             as broadcast shared-memory loads), multiplied with D[p, o] and accumulated in fp64;
//             var = sigma_f^2 + sigma_n^2 - sum.
// Nothing of size n_obs x P is written to HBM in fp64 (round 1: V, Y = 4 x 8 B x n_obs per candidate);
// the only staged operand is kx itself as fp16 hi / lo (4 B x n_obs per candidate, written once by
// kxgen_points_kernel, read back through L2 by cp.async.bulk).
//
// CTA = 11 warps as nlsum_kernel: warp 8 streams operand stages, warp 9 issues the MMAs, warp 10 streams
// the observation pack of a column tile, warps 0-7 are the epilogue (two column halves per lane quarter).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "nlsum.cuh"
#include "prep.cuh"

namespace basq {

constexpr int GPV_MT = 128;   // candidates per tile (MMA M)
constexpr int GPV_NT = 256;   // observations per column tile (MMA N)
constexpr int GPV_A_PIECE = (NLS_KB / 8) * GPV_MT * 16;   // 8 KB
constexpr int GPV_B_PIECE = (NLS_KB / 8) * GPV_NT * 16;   // 16 KB
constexpr int GPV_STAGE_BYTES = 2 * GPV_A_PIECE + 2 * GPV_B_PIECE;

struct GpvDev {
  const __half* kxh; const __half* kxl;   // [n_ptiles][KP / 8][128][8]
  const __half* th; const __half* tl;     // [KP / 256][KP / 8][256][8]  T rows scaled per row
  const unsigned char* obspack;           // [KP / 256][256][OBR]: zz[DP], b, 1 / (row scale * kx scale)
  const float* ppack;                     // [n_ptiles * 128][PPF]: a, x'[DP] of every candidate of the chunk
  int KP;                                 // n_obs padded to a multiple of 256
  int n_ptiles;
  int64_t n_points;                       // candidates in this chunk
  float os_f;
  double base;                            // sigma_f^2 + sigma_n^2
  double* var_out;                        // [n_points]
};

template <int DP>
struct GpvCfg {
  static constexpr int OBR = ((DP + 2) * 4 + 15) / 16 * 16;   // bytes per observation in the pack
  static constexpr int PPF = (DP + 1 + 3) / 4 * 4;            // floats per candidate in ppack
  static constexpr int OBS_BYTES = GPV_NT * OBR;
  static constexpr int NSTAGE = 3;
  static constexpr int OFF_STAGE = 0;
  static constexpr int OFF_OBS = NSTAGE * GPV_STAGE_BYTES;
  static constexpr int OFF_COMB = OFF_OBS + 2 * OBS_BYTES;
  static constexpr int OFF_BAR = OFF_COMB + 128 * 8;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "gpvar: shared memory budget");
};

template <int FAM, int DP>
__global__ void __launch_bounds__(NLS_THREADS, 1) gpvar_kernel(const GpvDev a) {
  using Cfg = GpvCfg<DP>;
  constexpr int NSTAGE = Cfg::NSTAGE, OBR = Cfg::OBR, NT = GPV_NT;
  extern __shared__ __align__(1024) unsigned char smem_gpv[];
  unsigned char* const smem = smem_gpv;
  unsigned char* sObs = smem + Cfg::OFF_OBS;
  double* sComb = reinterpret_cast<double*>(smem + Cfg::OFF_COMB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* s_full = bars;
  uint64_t* s_empty = bars + NSTAGE;
  uint64_t* t_full = bars + 2 * NSTAGE;
  uint64_t* t_empty = t_full + 2;
  uint64_t* o_full = t_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mma::mbar_init(&s_full[s], 1);
      mma::mbar_init(&s_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mma::mbar_init(&t_full[b], 1);
      mma::mbar_init(&t_empty[b], NLS_EPI_WARPS);
      mma::mbar_init(&o_full[b], 1);
    }
    mma::fence_barrier_init();
  }
  if (warp == NLS_EPI_WARPS + 1) mma::tmem_alloc(tmem_slot, 512);
  mma::tc_fence_before();
  __syncthreads();
  mma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_ctiles = a.KP / NT;
  auto nkb_of = [&](int c) { return min(a.KP, NT * (c + 1)) / NLS_KB; };   // K blocks up to the diagonal block

  if (warp < NLS_EPI_WARPS) {
    // ======================================================================== epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    uint32_t tc = 0, items_done = 0;
    for (int item = blockIdx.x; item < a.n_ptiles; item += gridDim.x, ++items_done) {
      const int64_t p = (int64_t)item * GPV_MT + row;
      float xr[DP];
      float pa = 0.f;
      {
        const float* pp = a.ppack + (size_t)p * Cfg::PPF;   // padded tiles are fully written by kxgen_points
        pa = pp[0];
#pragma unroll
        for (int i = 0; i < DP; ++i) xr[i] = pp[1 + i];
      }
      double acc = 0.0;
      for (int c = 0; c < n_ctiles; ++c, ++tc) {
        const uint32_t buf = tc & 1u, ph = (tc >> 1) & 1u;
        mma::mbar_wait(&o_full[buf], ph);
        mma::mbar_wait(&t_full[buf], ph);
        mma::tc_fence_after();
        const unsigned char* obs = sObs + buf * Cfg::OBS_BYTES;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NT + half * (NT / 2);
#pragma unroll 1
        for (int cb = 0; cb < NT / 2; cb += 32) {
          uint32_t v[32];
          mma::tmem_ld32(taddr + cb, v);
          mma::tmem_ld_wait();
          float part = 0.f;   // 32 terms in fp32 (each ~1e-7 accurate), widened once per block
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const unsigned char* ob = obs + (half * (NT / 2) + cb + cc) * OBR;
            constexpr int NF4 = OBR / 16;
            float f[NF4 * 4];   // zz[DP], b, tinv, pad
#pragma unroll
            for (int q4 = 0; q4 < NF4; ++q4) {
              const float4 w4 = *reinterpret_cast<const float4*>(ob + q4 * 16);
              f[q4 * 4 + 0] = w4.x; f[q4 * 4 + 1] = w4.y; f[q4 * 4 + 2] = w4.z; f[q4 * 4 + 3] = w4.w;
            }
            const float k = pair_eval_f32<FAM, DP>(xr, pa, f, f[DP], a.os_f);
            const float corr = __fmul_rn(__uint_as_float(v[cc]), f[DP + 1]);
            part = __fmaf_rn(k, corr, part);
          }
          acc += (double)part;
        }
        mma::tc_fence_before();
        __syncwarp();
        if (lane == 0) mma::mbar_arrive(&t_empty[buf]);
      }
      // combine the two column halves
      if (half == 1) sComb[row] = acc;
      mma::named_bar_sync(1, NLS_EPI_WARPS * 32);
      if (half == 0 && p < a.n_points) a.var_out[p] = a.base - (acc + sComb[row]);
      mma::named_bar_sync(2, NLS_EPI_WARPS * 32);   // sComb is reused by the next item
    }
  } else if (warp == NLS_EPI_WARPS) {
    // ======================================================================== operand producer
    if (lane == 0) {
      uint32_t it = 0;
      const size_t a_tile = (size_t)(a.KP / 8) * GPV_MT * 8;
      const size_t b_tile = (size_t)(a.KP / 8) * NT * 8;
      for (int item = blockIdx.x; item < a.n_ptiles; item += gridDim.x) {
        for (int c = 0; c < n_ctiles; ++c) {
          const int nkb = nkb_of(c);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % NSTAGE;
            mma::mbar_wait(&s_empty[s], ((it / NSTAGE) & 1u) ^ 1u);
            unsigned char* st = smem + Cfg::OFF_STAGE + (size_t)s * GPV_STAGE_BYTES;
            mma::mbar_expect_tx(&s_full[s], GPV_STAGE_BYTES);
            const size_t ka = (size_t)kb * (GPV_A_PIECE / 2), kbo = (size_t)kb * (GPV_B_PIECE / 2);
            mma::bulk_g2s(st, a.kxh + item * a_tile + ka, GPV_A_PIECE, &s_full[s]);
            mma::bulk_g2s(st + GPV_A_PIECE, a.kxl + item * a_tile + ka, GPV_A_PIECE, &s_full[s]);
            mma::bulk_g2s(st + 2 * GPV_A_PIECE, a.th + c * b_tile + kbo, GPV_B_PIECE, &s_full[s]);
            mma::bulk_g2s(st + 2 * GPV_A_PIECE + GPV_B_PIECE, a.tl + c * b_tile + kbo, GPV_B_PIECE, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == NLS_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t it = 0, tc = 0;
    for (int item = blockIdx.x; item < a.n_ptiles; item += gridDim.x) {
      for (int c = 0; c < n_ctiles; ++c, ++tc) {
        const uint32_t buf = tc & 1u;
        const int nkb = nkb_of(c);
        mma::mbar_wait(&t_empty[buf], ((tc >> 1) & 1u) ^ 1u);
        mma::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NSTAGE;
          mma::mbar_wait(&s_full[s], (it / NSTAGE) & 1u);
          mma::tc_fence_after();
          if (lane == 0) {
            const uint32_t st = mma::smem_u32(smem + Cfg::OFF_STAGE + (size_t)s * GPV_STAGE_BYTES);
            const uint32_t ahi = st, alo = st + GPV_A_PIECE;
            const uint32_t bhi = st + 2 * GPV_A_PIECE, blo = bhi + GPV_B_PIECE;
            const uint32_t d = tmem_base + buf * NT;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              const uint32_t aa = (p == 0) ? alo : ahi;
              const uint32_t bb = (p == 1) ? blo : bhi;
#pragma unroll
              for (int ks = 0; ks < NLS_KB / 16; ++ks) {
                const uint64_t ad = mma::smem_desc(aa + ks * 2 * (GPV_MT * 16), GPV_MT * 16, 128);
                const uint64_t bd = mma::smem_desc(bb + ks * 2 * (NT * 16), NT * 16, 128);
                umma_f16(d, ad, bd, IDESC, (kb > 0 || p > 0 || ks > 0) ? 1u : 0u);
              }
            }
            mma::umma_commit(&s_empty[s]);
            if (kb == nkb - 1) mma::umma_commit(&t_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================================================================== observation-pack producer
    if (lane == 0) {
      uint32_t tc = 0;
      for (int item = blockIdx.x; item < a.n_ptiles; item += gridDim.x) {
        for (int c = 0; c < n_ctiles; ++c, ++tc) {
          const uint32_t buf = tc & 1u;
          mma::mbar_wait(&t_empty[buf], ((tc >> 1) & 1u) ^ 1u);
          mma::mbar_expect_tx(&o_full[buf], Cfg::OBS_BYTES);
          mma::bulk_g2s(sObs + buf * Cfg::OBS_BYTES, a.obspack + (size_t)c * Cfg::OBS_BYTES, Cfg::OBS_BYTES, &o_full[buf]);
        }
      }
    }
  }

  mma::tc_fence_before();
  __syncthreads();
  if (warp == NLS_EPI_WARPS + 1) {
    mma::tc_fence_after();
    mma::tmem_dealloc(tmem_base, 512);
  }
}

// k(Xobs, x) of a chunk of raw candidates as the A operand ([tile of 128][KP / 8][128][8] fp16 hi / lo)
// plus the prepared coordinates of every candidate; one CTA per tile, one thread per candidate.
struct KxpDev {
  const float* X;          // [n_points, d] raw candidates of the chunk
  int64_t n_points;
  const float* ozz; const float* obz;
  int n_obs, KP;
  float kx_scale;
  __half* kxh; __half* kxl;
  float* ppack;
};

template <int FAM, int DP>
__global__ void __launch_bounds__(GPV_MT) kxgen_points_kernel(KParams kp, const KxpDev a) {
  constexpr int OB = 64, PPF = GpvCfg<DP>::PPF;
  __shared__ __align__(16) float s_oz[OB * DP];
  __shared__ float s_ob[OB];
  const int tile = blockIdx.x, r = threadIdx.x;
  const int64_t p = (int64_t)tile * GPV_MT + r;
  const bool ok = p < a.n_points;
  float x[DP];
  float pa = 0.f;
#pragma unroll
  for (int i = 0; i < DP; ++i) x[i] = 0.f;
  if (ok) {
    float xl[BASQ_MAX_DIM];
    float nrm;
    prep_point_f32(kp, a.X + p * kp.d, xl, &nrm);
#pragma unroll
    for (int i = 0; i < DP; ++i) x[i] = xl[i];
    pa = point_a_term(kp, nrm);
  }
  {
    float* pp = a.ppack + (size_t)p * PPF;
    pp[0] = pa;
#pragma unroll
    for (int i = 0; i < DP; ++i) pp[1 + i] = x[i];
  }
  const size_t a_tile = (size_t)(a.KP / 8) * GPV_MT * 8;
  uint4* oh = reinterpret_cast<uint4*>(a.kxh + (size_t)tile * a_tile) + r;
  uint4* ol = reinterpret_cast<uint4*>(a.kxl + (size_t)tile * a_tile) + r;
  for (int o0 = 0; o0 < a.KP; o0 += OB) {
    __syncthreads();
    for (int i = threadIdx.x; i < OB * DP; i += GPV_MT) {
      const int o = o0 + i / DP;
      s_oz[i] = o < a.n_obs ? a.ozz[(size_t)o * DP + i % DP] : 0.f;
    }
    if (threadIdx.x < OB) s_ob[threadIdx.x] = (o0 + (int)threadIdx.x < a.n_obs) ? a.obz[o0 + threadIdx.x] : 0.f;
    __syncthreads();
    for (int kc = 0; kc < OB / 8; ++kc) {
      __half2 h2[4], l2[4];
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        float vv[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int ol_ = kc * 8 + u + w;
          float kv = 0.f;
          if (ok && o0 + ol_ < a.n_obs) kv = pair_eval_f32<FAM, DP>(x, pa, &s_oz[ol_ * DP], s_ob[ol_], kp.os_f);
          vv[w] = __fmul_rn(kv, a.kx_scale);
        }
        const __half h0 = __float2half_rn(vv[0]), h1 = __float2half_rn(vv[1]);
        h2[u / 2] = __halves2half2(h0, h1);
        l2[u / 2] = __halves2half2(__float2half_rn(__fsub_rn(vv[0], __half2float(h0))),
                                   __float2half_rn(__fsub_rn(vv[1], __half2float(h1))));
      }
      const size_t chunk = (size_t)(o0 / 8 + kc) * GPV_MT;
      oh[chunk] = *reinterpret_cast<const uint4*>(h2);
      ol[chunk] = *reinterpret_cast<const uint4*>(l2);
    }
  }
}

template <int FAM, int DP>
int launch_gpvar_dp(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, const GpvDev& dev) {
  using Cfg = GpvCfg<DP>;
  BASQ_CHECK((size_t)Cfg::SMEM_BYTES <= ctx->smem_optin, BASQ_ERR_UNSUPPORTED,
             "gpvar: kernel needs %d B shared memory (limit %zu)", Cfg::SMEM_BYTES, ctx->smem_optin);
  if (dev.n_ptiles <= 0) return BASQ_OK;
  kxgen_points_kernel<FAM, DP><<<dev.n_ptiles, GPV_MT, 0, ctx->stream>>>(kp, kx);
  BASQ_CUDA(cudaFuncSetAttribute(gpvar_kernel<FAM, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = std::min(ctx->num_sms, dev.n_ptiles);
  gpvar_kernel<FAM, DP><<<grid, NLS_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(dev);
  ctx->launches += 2;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

template <int FAM>
int launch_gpvar_family(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, const GpvDev& dev) {
  switch (kp.dp) {
    case 2: return launch_gpvar_dp<FAM, 2>(ctx, kp, kx, dev);
    case 4: return launch_gpvar_dp<FAM, 4>(ctx, kp, kx, dev);
    case 6: return launch_gpvar_dp<FAM, 6>(ctx, kp, kx, dev);
    case 8: return launch_gpvar_dp<FAM, 8>(ctx, kp, kx, dev);
    case 10: return launch_gpvar_dp<FAM, 10>(ctx, kp, kx, dev);
    case 12: return launch_gpvar_dp<FAM, 12>(ctx, kp, kx, dev);
    case 16: return launch_gpvar_dp<FAM, 16>(ctx, kp, kx, dev);
    case 20: return launch_gpvar_dp<FAM, 20>(ctx, kp, kx, dev);
    case 24: return launch_gpvar_dp<FAM, 24>(ctx, kp, kx, dev);
    case 32: return launch_gpvar_dp<FAM, 32>(ctx, kp, kx, dev);
  }
  set_error("gpvar: no kernel compiled for padded dimension %d", kp.dp);
  return BASQ_ERR_UNSUPPORTED;
}

template <int DP>
__global__ void obspack_kernel(const float* __restrict__ ozz, const float* __restrict__ obz,
                               const float* __restrict__ tinv, int n_obs, int KP, unsigned char* __restrict__ pack) {
  constexpr int OBR = GpvCfg<DP>::OBR;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= KP) return;
  float* f = reinterpret_cast<float*>(pack + (size_t)o * OBR);
  for (int i = 0; i < OBR / 4; ++i) f[i] = 0.f;
  if (o < n_obs) {
    for (int i = 0; i < DP; ++i) f[i] = ozz[(size_t)o * DP + i];
    f[DP] = obz[o];
    f[DP + 1] = tinv[o];
  }
}

int launch_gpvar_rbf(basq_ctx*, const KParams&, const KxpDev&, const GpvDev&);
int launch_gpvar_m15(basq_ctx*, const KParams&, const KxpDev&, const GpvDev&);
int launch_gpvar_m25(basq_ctx*, const KParams&, const KxpDev&, const GpvDev&);
int launch_obspack(basq_ctx* ctx, int dp, const float* ozz, const float* obz, const float* tinv, int n_obs, int KP,
                   unsigned char* pack, int* obr_out);

}  // namespace basq
