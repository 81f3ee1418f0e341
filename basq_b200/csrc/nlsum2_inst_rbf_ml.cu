// nlsum2_kernel (CTA pairs) instantiations: rbf, MMLT (see nlsum2.cuh)
#include "nlsum2.cuh"
namespace basq {
int launch_nlsum2_rbf_ml(basq_ctx* ctx, int dp, const NlsDev& dev) {
  return launch_nlsum2_family<BASQ_RBF, NL_MMLT>(ctx, dp, dev);
}
}  // namespace basq
