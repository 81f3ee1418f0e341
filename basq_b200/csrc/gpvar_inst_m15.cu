// gpvar_kernel / kxgen_points_kernel instantiations: m15 (see gpvar.cuh)
#include "gpvar.cuh"
namespace basq {
int launch_gpvar_m15(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, const GpvDev& dev) {
  return launch_gpvar_family<BASQ_MATERN15>(ctx, kp, kx, dev);
}
}  // namespace basq
