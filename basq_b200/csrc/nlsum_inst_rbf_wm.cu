// nlsum_kernel instantiations: rbf, WSABI-M (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_rbf_wm(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_RBF, NL_WSABIM>(ctx, dp, dev, mode);
}
}  // namespace basq
