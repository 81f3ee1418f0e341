// Landmark operand of the tensor-core set-sum kernel (see setsum_mma.cuh).
#include <algorithm>

#include "setsum_mma.cuh"

namespace basq {

namespace {
template <int DP>
void cfg_of(int* mt, int* ka) {
  *mt = MmaCfg<DP>::MT;
  *ka = MmaCfg<DP>::KA;
}
bool cfg(int dp, int* mt, int* ka) {
  switch (dp) {
    case 2: cfg_of<2>(mt, ka); return true;
    case 4: cfg_of<4>(mt, ka); return true;
    case 6: cfg_of<6>(mt, ka); return true;
    case 8: cfg_of<8>(mt, ka); return true;
    case 10: cfg_of<10>(mt, ka); return true;
    case 12: cfg_of<12>(mt, ka); return true;
    case 16: cfg_of<16>(mt, ka); return true;
    case 20: cfg_of<20>(mt, ka); return true;
    case 24: cfg_of<24>(mt, ka); return true;
    case 32: cfg_of<32>(mt, ka); return true;
  }
  return false;
}
}  // namespace

int lmA_tiles(int dp, int count) {
  int mt = 1, ka = 8;
  if (!cfg(dp, &mt, &ka)) return 0;
  return ceil_div(count, mt * 128) * mt;
}

size_t lmA_floats(int dp, int count) {
  int mt = 1, ka = 8;
  if (!cfg(dp, &mt, &ka)) return 0;
  return (size_t)lmA_tiles(dp, count) * ka * 128;
}

int build_lmA(basq_ctx* ctx, int dp, const float* zz, const float* bz, int Mtot, float* lmA) {
  const int nt = lmA_tiles(dp, Mtot);
  switch (dp) {
    case 2: return launch_build_lmA_dp<2>(ctx, zz, bz, Mtot, nt, lmA);
    case 4: return launch_build_lmA_dp<4>(ctx, zz, bz, Mtot, nt, lmA);
    case 6: return launch_build_lmA_dp<6>(ctx, zz, bz, Mtot, nt, lmA);
    case 8: return launch_build_lmA_dp<8>(ctx, zz, bz, Mtot, nt, lmA);
    case 10: return launch_build_lmA_dp<10>(ctx, zz, bz, Mtot, nt, lmA);
    case 12: return launch_build_lmA_dp<12>(ctx, zz, bz, Mtot, nt, lmA);
    case 16: return launch_build_lmA_dp<16>(ctx, zz, bz, Mtot, nt, lmA);
    case 20: return launch_build_lmA_dp<20>(ctx, zz, bz, Mtot, nt, lmA);
    case 24: return launch_build_lmA_dp<24>(ctx, zz, bz, Mtot, nt, lmA);
    case 32: return launch_build_lmA_dp<32>(ctx, zz, bz, Mtot, nt, lmA);
  }
  set_error("no tensor-core set-sum kernel for padded dimension %d", dp);
  return BASQ_ERR_UNSUPPORTED;
}

}  // namespace basq
