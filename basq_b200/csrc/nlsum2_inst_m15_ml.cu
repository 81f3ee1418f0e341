// nlsum2_kernel (CTA pairs) instantiations: m15, MMLT (see nlsum2.cuh)
#include "nlsum2.cuh"
namespace basq {
int launch_nlsum2_m15_ml(basq_ctx* ctx, int dp, const NlsDev& dev) {
  return launch_nlsum2_family<BASQ_MATERN15, NL_MMLT>(ctx, dp, dev);
}
}  // namespace basq
