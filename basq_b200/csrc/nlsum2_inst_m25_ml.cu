// nlsum2_kernel (CTA pairs) instantiations: m25, MMLT (see nlsum2.cuh)
#include "nlsum2.cuh"
namespace basq {
int launch_nlsum2_m25_ml(basq_ctx* ctx, int dp, const NlsDev& dev) {
  return launch_nlsum2_family<BASQ_MATERN25, NL_MMLT>(ctx, dp, dev);
}
}  // namespace basq
