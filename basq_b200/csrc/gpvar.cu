// Host side of the fused posterior-variance kernel (gpvar.cuh): T = tril(W + W^T) as row-scaled fp16
// hi / lo tiles, the observation pack, chunking over the candidates.
#include <math.h>

#include <algorithm>

#include "gpvar.cuh"

namespace basq {

namespace {

// T[i, j] = 2 W[i, j] (j < i), W[i, i] (j == i), 0 (j > i):  v^T W v = sum_i v_i (T v)_i for symmetric W
__global__ void tri_fold16_kernel(const double* __restrict__ W, int n, double* __restrict__ T) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * n) return;
  const int i = (int)(t / n), j = (int)(t % n);
  T[t] = j < i ? W[t] + W[(int64_t)j * n + i] : (j == i ? W[t] : 0.0);
}

template <int DP>
int obspack_dp(basq_ctx* ctx, const float* ozz, const float* obz, const float* tinv, int n_obs, int KP,
               unsigned char* pack) {
  obspack_kernel<DP><<<ceil_div(KP, 128), 128, 0, ctx->stream>>>(ozz, obz, tinv, n_obs, KP, pack);
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

int obr_of(int dp) { return ((dp + 2) * 4 + 15) / 16 * 16; }
int ppf_of(int dp) { return (dp + 1 + 3) / 4 * 4; }

}  // namespace

int launch_obspack(basq_ctx* ctx, int dp, const float* ozz, const float* obz, const float* tinv, int n_obs, int KP,
                   unsigned char* pack, int* obr_out) {
  if (obr_out) *obr_out = obr_of(dp);
  switch (dp) {
    case 2: return obspack_dp<2>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 4: return obspack_dp<4>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 6: return obspack_dp<6>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 8: return obspack_dp<8>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 10: return obspack_dp<10>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 12: return obspack_dp<12>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 16: return obspack_dp<16>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 20: return obspack_dp<20>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 24: return obspack_dp<24>(ctx, ozz, obz, tinv, n_obs, KP, pack);
    case 32: return obspack_dp<32>(ctx, ozz, obz, tinv, n_obs, KP, pack);
  }
  set_error("gpvar: no kernel compiled for padded dimension %d", dp);
  return BASQ_ERR_UNSUPPORTED;
}

// var_out[N] = sigma_f^2 + sigma_n^2 - k_x^T W k_x for fp32 candidates X [N, d]
int gp_variance_tc(basq_ctx* ctx, const basq_kernel_desc* desc, const KParams& kp, const LmView& lmobs, const void* X,
                   int64_t N, double* var_out) {
  BASQ_CHECK(desc->dtype == BASQ_F32 && lmobs.dtype == BASQ_F32, BASQ_ERR_INVALID, "gpvar: fp32 inputs only");
  const int n_obs = desc->n_obs;
  const int KP = ceil_div(n_obs, GPV_NT) * GPV_NT;
  const int n_ct = KP / GPV_NT;
  const float kx_scale = ldexpf(1.f, NLS_KX_SHIFT - (int)ceil(log2(kp.outputscale)));
  DevBuf T, rscale, tinv, th, tl, pack;
  BASQ_TRY(T.alloc(ctx, sizeof(double) * (size_t)n_obs * n_obs));
  BASQ_TRY(rscale.alloc(ctx, sizeof(float) * KP));
  BASQ_TRY(tinv.alloc(ctx, sizeof(float) * KP));
  const size_t t_halves = (size_t)n_ct * (KP / 8) * GPV_NT * 8;
  BASQ_TRY(th.alloc(ctx, t_halves * 2));
  BASQ_TRY(tl.alloc(ctx, t_halves * 2));
  BASQ_TRY(pack.alloc(ctx, (size_t)KP * obr_of(kp.dp)));
  tri_fold16_kernel<<<(unsigned)ceil_div64((int64_t)n_obs * n_obs, 256), 256, 0, ctx->stream>>>(desc->W, n_obs,
                                                                                              T.as<double>());
  rowscale16_kernel<256><<<KP, 256, 0, ctx->stream>>>(T.as<double>(), n_obs, n_obs, n_obs, kx_scale, rscale.as<float>(),
                                                      tinv.as<float>());
  const int64_t tot = (int64_t)n_ct * (KP / 8) * GPV_NT;
  split16_kernel<GPV_NT><<<(unsigned)ceil_div64(tot, 256), 256, 0, ctx->stream>>>(
      T.as<double>(), n_obs, n_obs, n_obs, KP, n_ct, rscale.as<float>(), th.as<__half>(), tl.as<__half>());
  ctx->launches += 3;
  BASQ_CUDA(cudaGetLastError());
  BASQ_TRY(launch_obspack(ctx, kp.dp, reinterpret_cast<const float*>(lmobs.zz), lmobs.b, tinv.as<float>(), n_obs, KP,
                          pack.as<unsigned char>(), nullptr));
  // chunk of candidates: the kx operand (4 B x KP per candidate) takes at most ~2 GB
  static const int64_t budget = [] {
    const char* e = getenv("BASQ_GPV_CHUNK_MB");
    return (int64_t)(e ? atoll(e) : 2048) << 20;
  }();
  int64_t P = std::max<int64_t>(GPV_MT, budget / (4ll * KP) / GPV_MT * GPV_MT);
  P = std::min<int64_t>(P, ceil_div64(N, GPV_MT) * GPV_MT);
  DevBuf kxh, kxl, ppack;
  BASQ_TRY(kxh.alloc(ctx, (size_t)P * KP * 2));
  BASQ_TRY(kxl.alloc(ctx, (size_t)P * KP * 2));
  BASQ_TRY(ppack.alloc(ctx, sizeof(float) * (size_t)P * ppf_of(kp.dp)));
  for (int64_t p0 = 0; p0 < N; p0 += P) {
    const int64_t cnt = std::min<int64_t>(P, N - p0);
    KxpDev kx;
    kx.X = reinterpret_cast<const float*>(X) + p0 * desc->d;
    kx.n_points = cnt;
    kx.ozz = reinterpret_cast<const float*>(lmobs.zz);
    kx.obz = lmobs.b;
    kx.n_obs = n_obs;
    kx.KP = KP;
    kx.kx_scale = kx_scale;
    kx.kxh = kxh.as<__half>();
    kx.kxl = kxl.as<__half>();
    kx.ppack = ppack.as<float>();
    GpvDev d;
    d.kxh = kx.kxh; d.kxl = kx.kxl;
    d.th = th.as<__half>(); d.tl = tl.as<__half>();
    d.obspack = pack.as<unsigned char>();
    d.ppack = kx.ppack;
    d.KP = KP;
    d.n_ptiles = (int)ceil_div64(cnt, GPV_MT);
    d.n_points = cnt;
    d.os_f = kp.os_f;
    d.base = desc->outputscale + desc->noise;
    d.var_out = var_out + p0;
    switch (kp.family) {
      case BASQ_RBF: BASQ_TRY(launch_gpvar_rbf(ctx, kp, kx, d)); break;
      case BASQ_MATERN15: BASQ_TRY(launch_gpvar_m15(ctx, kp, kx, d)); break;
      default: BASQ_TRY(launch_gpvar_m25(ctx, kp, kx, d)); break;
    }
  }
  BASQ_CUDA(cudaStreamSynchronize(ctx->stream));  // scratch goes out of scope
  return BASQ_OK;
}

}  // namespace basq
