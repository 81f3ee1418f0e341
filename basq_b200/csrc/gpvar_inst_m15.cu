// gpvar_fused_kernel / kxgen_points_kernel instantiations: m15 (see gpvar.cuh)
#include "gpvar.cuh"
namespace basq {
int launch_kxgen_points_m15(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, int n_ptiles) {
  return launch_kxgen_points_family<BASQ_MATERN15>(ctx, kp, kx, n_ptiles);
}
int launch_gpvar_fused_m15(basq_ctx* ctx, const KParams& kp, const GpfDev& dev) {
  return launch_gpvar_fused_family<BASQ_MATERN15>(ctx, kp, dev);
}
}  // namespace basq
