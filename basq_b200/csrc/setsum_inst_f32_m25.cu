// Instantiations of the set-sum kernel: float evaluation, BASQ_MATERN25.
#include "setsum_impl.cuh"
namespace basq {
int launch_setsum_f32_m25(basq_ctx* ctx, int dp, const SetSumDev& dev) {
  return launch_setsum_family<float, BASQ_MATERN25>(ctx, dp, dev);
}
}  // namespace basq
