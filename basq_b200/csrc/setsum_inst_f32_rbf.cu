// Instantiations of the set-sum kernel: float evaluation, BASQ_RBF.
#include "setsum_impl.cuh"
namespace basq {
int launch_setsum_f32_rbf(basq_ctx* ctx, int dp, const SetSumDev& dev) {
  return launch_setsum_family<float, BASQ_RBF>(ctx, dp, dev);
}
}  // namespace basq
