// nlsum2_kernel (CTA pairs) instantiations: m15, WSABI-M (see nlsum2.cuh)
#include "nlsum2.cuh"
namespace basq {
int launch_nlsum2_m15_wm(basq_ctx* ctx, int dp, const NlsDev& dev) {
  return launch_nlsum2_family<BASQ_MATERN15, NL_WSABIM>(ctx, dp, dev);
}
}  // namespace basq
