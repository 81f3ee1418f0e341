// Instantiation of the tensor-core set-sum kernel for one kernel family (see setsum_mma.cuh).
#include <algorithm>

#include "setsum_mma.cuh"
namespace basq {
int launch_setsum_mma_rbf(basq_ctx* ctx, int dp, const SetSumMmaDev& dev) {
  return launch_setsum_mma_family<0>(ctx, dp, dev);
}
}  // namespace basq
