// nlsum_kernel instantiations: m25, WSABI-M (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_m25_wm(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_MATERN25, NL_WSABIM>(ctx, dp, dev, mode);
}
}  // namespace basq
