// Instantiation of the tensor-core set-sum kernel for one kernel family (see setsum_mma.cuh).
#include <algorithm>

#include "setsum_mma.cuh"
namespace basq {
int launch_setsum_mma_m25(basq_ctx* ctx, int dp, const SetSumMmaDev& dev) {
  return launch_setsum_mma_family<2>(ctx, dp, dev);
}
}  // namespace basq
