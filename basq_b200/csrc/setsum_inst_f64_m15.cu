// Instantiations of the set-sum kernel: double evaluation, BASQ_MATERN15.
#include "setsum_impl.cuh"
namespace basq {
int launch_setsum_f64_m15(basq_ctx* ctx, int dp, const SetSumDev& dev) {
  return launch_setsum_family<double, BASQ_MATERN15>(ctx, dp, dev);
}
}  // namespace basq
