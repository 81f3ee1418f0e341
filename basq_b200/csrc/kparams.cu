// Kernel-parameter preparation: scaling constants, centring, landmark arrays.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <time.h>

#include "common.cuh"
#include "prep.cuh"
#include "setsum_mma.cuh"

namespace basq {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

void trace_point(basq_ctx* ctx, const char* label) {
  if (!ctx->trace) return;
  cudaStreamSynchronize(ctx->stream);
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  fprintf(stderr, "[basq trace] %-28s +%9.3f ms\n", label, ctx->trace_t0 > 0 ? now - ctx->trace_t0 : 0.0);
  ctx->trace_t0 = now;
}

int make_kparams(const basq_kernel_desc* desc, KParams* kp) {
  BASQ_CHECK(desc != nullptr, BASQ_ERR_INVALID, "kernel descriptor is NULL");
  BASQ_CHECK(desc->d >= 1 && desc->d <= BASQ_MAX_DIM, BASQ_ERR_UNSUPPORTED,
             "input dimension d=%d outside 1..%d", desc->d, BASQ_MAX_DIM);
  BASQ_CHECK(desc->family >= BASQ_RBF && desc->family <= BASQ_MATERN25, BASQ_ERR_INVALID,
             "unknown kernel family %d", desc->family);
  BASQ_CHECK(desc->mode >= BASQ_PLAIN && desc->mode <= BASQ_MMLT_G, BASQ_ERR_INVALID, "unknown kernel mode %d",
             desc->mode);
  BASQ_CHECK(desc->dtype == BASQ_F32 || desc->dtype == BASQ_F64, BASQ_ERR_INVALID, "unknown dtype %d",
             desc->dtype);
  BASQ_CHECK(desc->outputscale > 0.0 && isfinite(desc->outputscale), BASQ_ERR_INVALID,
             "outputscale must be positive and finite");
  if (desc->mode != BASQ_PLAIN) {
    BASQ_CHECK(desc->n_obs > 0 && desc->Xobs && desc->W && desc->alpha, BASQ_ERR_INVALID,
               "mode %d needs n_obs > 0 and the Xobs / W / alpha caches", desc->mode);
  }
  memset(kp, 0, sizeof(*kp));
  kp->family = desc->family;
  kp->d = desc->d;
  kp->dp = padded_dim(desc->d);
  kp->outputscale = desc->outputscale;
  kp->log2_os = (float)log2(desc->outputscale);
  kp->os_f = (float)desc->outputscale;
  const double rbf_scale = sqrt(0.5 * 1.4426950408889634);  // sqrt(log2(e) / 2)
  for (int i = 0; i < desc->d; ++i) {
    const double l = desc->lengthscale[i];
    BASQ_CHECK(l > 0.0 && isfinite(l), BASQ_ERR_INVALID, "lengthscale[%d] must be positive and finite", i);
    kp->scale_d[i] = 1.0 / l;
    kp->scale_f[i] = (float)((desc->family == BASQ_RBF ? rbf_scale : 1.0) / l);
  }
  return BASQ_OK;
}

// ---------------------------------------------------------------------------------------------
// centre = mean of the landmark rows (removes the common offset before the fp32 expansion)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void column_mean_kernel(const T* __restrict__ Z, int64_t M, int d, double* __restrict__ out) {
  // one block per column; fixed-order tree reduction => deterministic
  __shared__ double sh[256];
  const int c = blockIdx.x;
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < M; r += blockDim.x) s += (double)Z[r * d + c];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = sh[0] / (double)M;
}

int compute_center(basq_ctx* ctx, const basq_kernel_desc* desc, const void* Z, int64_t M, KParams* kp) {
  DevBuf tmp;
  BASQ_TRY(tmp.alloc(ctx, sizeof(double) * BASQ_MAX_DIM));
  if (desc->dtype == BASQ_F32)
    column_mean_kernel<float><<<desc->d, 256, 0, ctx->stream>>>((const float*)Z, M, desc->d, tmp.as<double>());
  else
    column_mean_kernel<double><<<desc->d, 256, 0, ctx->stream>>>((const double*)Z, M, desc->d, tmp.as<double>());
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  double host[BASQ_MAX_DIM];
  BASQ_CUDA(cudaMemcpyAsync(host, tmp.p, sizeof(double) * desc->d, cudaMemcpyDeviceToHost, ctx->stream));
  BASQ_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < desc->d; ++i) {
    BASQ_CHECK(isfinite(host[i]), BASQ_ERR_NUMERIC, "landmark coordinates are not finite (column %d)", i);
    // the fp32 path subtracts the centre in fp32: keep it exactly representable there
    kp->center[i] = (desc->dtype == BASQ_F32) ? (double)(float)host[i] : host[i];
  }
  return BASQ_OK;
}

// ---------------------------------------------------------------------------------------------
// landmark arrays
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void prep_landmarks_f32_kernel(KParams kp, const T* __restrict__ Z, int64_t M, int64_t row0,
                                          float* __restrict__ zz, float* __restrict__ b) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float xs[BASQ_MAX_DIM];
  float nrm;
  prep_point_f32(kp, Z + r * kp.d, xs, &nrm);
  const float sgn = (kp.family == BASQ_RBF) ? 2.f : -2.f;
  for (int i = 0; i < kp.dp; ++i) zz[(row0 + r) * kp.dp + i] = (i < kp.d) ? __fmul_rn(sgn, xs[i]) : 0.f;
  b[row0 + r] = (kp.family == BASQ_RBF) ? __fadd_rn(-nrm, kp.log2_os) : nrm;
}

template <typename T>
__global__ void prep_landmarks_f64_kernel(KParams kp, const T* __restrict__ Z, int64_t M, int64_t row0,
                                          double* __restrict__ zs) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  for (int i = 0; i < kp.dp; ++i)
    zs[(row0 + r) * kp.dp + i] = (i < kp.d) ? ((double)Z[r * kp.d + i] - kp.center[i]) * kp.scale_d[i] : 0.0;
}

int prep_landmarks(basq_ctx* ctx, const KParams& kp, int dtype, const void* Z0, int64_t M0, const void* Z1,
                   int64_t M1, Landmarks* out) {
  const int64_t M = M0 + M1;
  BASQ_CHECK(M > 0 && M < (1ll << 30), BASQ_ERR_INVALID, "landmark count %lld out of range", (long long)M);
  out->count = (int)M;
  out->dp = kp.dp;
  out->dtype = dtype;
  const void* src[2] = {Z0, Z1};
  const int64_t cnt[2] = {M0, M1};
  if (dtype == BASQ_F32) {
    BASQ_TRY(out->zz.alloc(ctx, sizeof(float) * M * kp.dp));
    BASQ_TRY(out->b.alloc(ctx, sizeof(float) * M));
    int64_t row0 = 0;
    for (int s = 0; s < 2; ++s) {
      if (cnt[s] == 0) continue;
      prep_landmarks_f32_kernel<float><<<ceil_div(cnt[s], 256), 256, 0, ctx->stream>>>(
          kp, (const float*)src[s], cnt[s], row0, out->zz.as<float>(), out->b.as<float>());
      ctx->launches++;
      row0 += cnt[s];
    }
    BASQ_CUDA(cudaGetLastError());
    BASQ_TRY(out->lmA.alloc(ctx, sizeof(float) * lmA_floats(kp.dp, (int)M)));
    BASQ_TRY(build_lmA(ctx, kp.dp, out->zz.as<float>(), out->b.as<float>(), (int)M, out->lmA.as<float>()));
  } else {
    BASQ_TRY(out->zz.alloc(ctx, sizeof(double) * M * kp.dp));
    int64_t row0 = 0;
    for (int s = 0; s < 2; ++s) {
      if (cnt[s] == 0) continue;
      prep_landmarks_f64_kernel<double><<<ceil_div(cnt[s], 256), 256, 0, ctx->stream>>>(
          kp, (const double*)src[s], cnt[s], row0, out->zz.as<double>());
      ctx->launches++;
      row0 += cnt[s];
    }
  }
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

}  // namespace basq
