// nlsum_kernel / kxgen_kernel instantiations: m25, MMLT (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_m25_ml(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_MATERN25, NL_MMLT>(ctx, dp, dev, mode);
}
int launch_kxgen_m25(basq_ctx* ctx, int dp, const KxDev& dev) { return launch_kxgen_family<BASQ_MATERN25>(ctx, dp, dev); }
}  // namespace basq
