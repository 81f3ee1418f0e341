// gpvar_kernel / kxgen_points_kernel instantiations: m25 (see gpvar.cuh)
#include "gpvar.cuh"
namespace basq {
int launch_gpvar_m25(basq_ctx* ctx, const KParams& kp, const KxpDev& kx, const GpvDev& dev) {
  return launch_gpvar_family<BASQ_MATERN25>(ctx, kp, kx, dev);
}
}  // namespace basq
