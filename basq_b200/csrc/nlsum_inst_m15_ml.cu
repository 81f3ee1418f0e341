// nlsum_kernel / kxgen_kernel instantiations: m15, MMLT (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_m15_ml(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_MATERN15, NL_MMLT>(ctx, dp, dev, mode);
}
int launch_kxgen_m15(basq_ctx* ctx, int dp, const KxDev& dev) { return launch_kxgen_family<BASQ_MATERN15>(ctx, dp, dev); }
}  // namespace basq
