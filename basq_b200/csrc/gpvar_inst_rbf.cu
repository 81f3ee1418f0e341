// kxgen_points_kernel instantiations: rbf (see gpvar.cuh)
#include "gpvar.cuh"
namespace basq {
int launch_kxgen_points_rbf(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, int n_ptiles) {
  return launch_kxgen_points_family<BASQ_RBF>(ctx, kp, kx, n_ptiles);
}
}  // namespace basq
