// nlsum_kernel / kxgen_kernel instantiations: rbf, MMLT (see nlsum.cuh)
#include "nlsum.cuh"
namespace basq {
int launch_nlsum_rbf_ml(basq_ctx* ctx, int dp, const NlsDev& dev, int mode) {
  return launch_nlsum_family<BASQ_RBF, NL_MMLT>(ctx, dp, dev, mode);
}
int launch_kxgen_rbf(basq_ctx* ctx, int dp, const KxDev& dev) { return launch_kxgen_family<BASQ_RBF>(ctx, dp, dev); }
}  // namespace basq
