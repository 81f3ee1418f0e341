// Point preparation shared by every kernel that consumes raw coordinates, so that the scaled
// coordinates (and hence every fp32 kernel value) are bit-identical wherever they are rebuilt.
#pragma once
#include "common.cuh"

namespace basq {

// xs[0..dp) = ((float)x - (float)centre) * scale ; *nrm = sum xs^2 (fixed order)
template <typename T>
__device__ __forceinline__ void prep_point_f32(const KParams& kp, const T* __restrict__ x, float* xs, float* nrm) {
  float n = 0.f;
  for (int i = 0; i < kp.dp; ++i) {
    float v = 0.f;
    if (i < kp.d) v = __fmul_rn(__fsub_rn((float)x[i], (float)kp.center[i]), kp.scale_f[i]);
    xs[i] = v;
    n = __fmaf_rn(v, v, n);
  }
  *nrm = n;
}

// the "a" term of the fp32 expansion for a point
__device__ __forceinline__ float point_a_term(const KParams& kp, float nrm) {
  return (kp.family == BASQ_RBF) ? -nrm : nrm;
}

template <typename T>
__device__ __forceinline__ void prep_point_f64(const KParams& kp, const T* __restrict__ x, double* xs) {
  for (int i = 0; i < kp.dp; ++i) xs[i] = (i < kp.d) ? ((double)x[i] - kp.center[i]) * kp.scale_d[i] : 0.0;
}

}  // namespace basq
