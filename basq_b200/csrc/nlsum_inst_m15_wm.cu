// nlsum_kernel instantiations: m15, WSABI-M (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_m15_wm(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_MATERN15, NL_WSABIM>(ctx, dp, dev, mode);
}
}  // namespace basq
