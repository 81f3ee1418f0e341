// gpvar_kernel / kxgen_points_kernel instantiations: rbf (see gpvar.cuh)
#include "gpvar.cuh"
namespace basq {
int launch_gpvar_rbf(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, const GpvDev& dev) {
  return launch_gpvar_family<BASQ_RBF>(ctx, kp, kx, dev);
}
}  // namespace basq
