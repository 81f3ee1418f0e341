// Instantiations of the set-sum kernel: double evaluation, BASQ_RBF.
#include "setsum_impl.cuh"
namespace basq {
int launch_setsum_f64_rbf(basq_ctx* ctx, int dp, const SetSumDev& dev) {
  return launch_setsum_family<double, BASQ_RBF>(ctx, dp, dev);
}
}  // namespace basq
