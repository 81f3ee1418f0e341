// nlsum_kernel on CTA pairs (tcgen05 cta_group::2): SETSUM mode of nlsum.cuh with M = 256.
//
// Why: in the one-CTA kernel every tcgen05.mma (M = 128, N = 256, K = 16) reads 4 KB of A and 8 KB of B
// from shared memory in 128 cycles, and the operand ring writes 64 B/clk on top: 160 B/clk against the
// 128 B/clk the shared-memory pipe delivers, so the tensor pipe cannot exceed ~80 % (measured 66-71 %,
// profiles/r02).  A CTA pair on the two SMs of a TPC runs one 256 x 256 x 16 MMA: each SM keeps the A
// rows of ITS landmark tile and HALF of the candidates' B rows, the pair's tensor cores read each other's
// half, and both the shared-memory reads (64 B/clk) and the L2 -> SM stream (43 B/clk, stages of 32 KB:
// six in flight) drop by a third.
//
// Roles per CTA (11 warps as nlsum_kernel):
//   warp 8   operand producer: A hi / lo of the CTA's own landmark tile + hi / lo of its half of the
//            candidate tile, local full barrier (expect_tx)
//   warp 9   rank 0: waits for BOTH CTAs' stages, issues tcgen05.mma.cta_group::2, commits the stage's
//            empty barrier and the accumulator's full barrier to both CTAs (multicast commit);
//            rank 1: relays "my stage is full" to rank 0 (remote mbarrier arrive)
//   warp 10  record producer (each CTA stages the tile's candidate records for its own epilogue)
//   warps 0-7 epilogue on the CTA's own TMEM half (its landmark tile); "accumulator drained" is counted on
//            rank 0's barrier by all 16 epilogue warps of the pair
// Candidate tiles arrive from kxgen_kernel in the split-half layout [tile][half][K / 8][128][8 halves].
#pragma once
#include "nlsum.cuh"

namespace basq {

constexpr int NLS2_PIECE = (NLS_KB / 8) * 128 * 16;      // 8 KB: 128 rows x 32 halves
constexpr int NLS2_STAGE_BYTES = 4 * NLS2_PIECE;         // A hi, A lo, B-half hi, B-half lo

template <int DP>
struct Nls2Cfg {
  static constexpr int RB = NlsCfg<DP>::RB;
  static constexpr int REC_BYTES = NLS_NT * RB;
  static constexpr int COMB_BYTES = 2 * 128 * NLS_JT * 8;
  static constexpr int FIXED = REC_BYTES + COMB_BYTES + 512;
  static constexpr int NSTAGE_RAW = (227 * 1024 - FIXED) / NLS2_STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 6 ? 6 : NSTAGE_RAW;
  static_assert(NSTAGE >= 2, "nlsum2: shared memory budget");
  static constexpr int OFF_STAGE = 0;
  static constexpr int OFF_REC = NSTAGE * NLS2_STAGE_BYTES;
  static constexpr int OFF_COMB = OFF_REC + REC_BYTES;
  static constexpr int OFF_BAR = OFF_COMB + COMB_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
};

namespace mma {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// commit all MMAs issued so far by this thread; arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
}  // namespace mma

// a.n_mtiles counts landmark tiles (padded to an even number by the host); kx tiles in the split-half layout
template <int FAM, int DP, int NL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NLS_THREADS, 1) nlsum2_kernel(const NlsDev a) {
  using Cfg = Nls2Cfg<DP>;
  constexpr int NSTAGE = Cfg::NSTAGE, NT = NLS_NT;
  extern __shared__ __align__(1024) unsigned char smem_nls2[];
  unsigned char* const smem = smem_nls2;
  unsigned char* sRec = smem + Cfg::OFF_REC;
  double* sComb = reinterpret_cast<double*>(smem + Cfg::OFF_COMB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* s_full = bars;                  // [NSTAGE] this CTA's stage landed
  uint64_t* p_full = bars + NSTAGE;         // [NSTAGE] rank 0 only: the peer's stage landed
  uint64_t* s_empty = bars + 2 * NSTAGE;    // [NSTAGE] MMAs that read the stage are done (multicast commit)
  uint64_t* t_full = bars + 3 * NSTAGE;     // [2] accumulator ready (multicast commit)
  uint64_t* t_empty = t_full + 2;           // [2] rank 0 only: drained by the 16 epilogue warps of the pair
  uint64_t* r_full = t_empty + 2;
  uint64_t* r_empty = r_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = mma::cluster_ctarank();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mma::mbar_init(&s_full[s], 1);
      mma::mbar_init(&p_full[s], 1);
      mma::mbar_init(&s_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mma::mbar_init(&t_full[b], 1);
      mma::mbar_init(&t_empty[b], 2 * NLS_EPI_WARPS);
    }
    mma::mbar_init(r_full, 1);
    mma::mbar_init(r_empty, NLS_EPI_WARPS);
    mma::fence_barrier_init();
  }
  if (warp == NLS_EPI_WARPS + 1) mma::tmem_alloc2(tmem_slot, 512);
  mma::tc_fence_before();
  mma::cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  mma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_mpairs = a.n_mtiles / 2;
  const int n_items = n_mpairs * a.n_jg;
  const int nkb = a.KP / NLS_KB;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp < NLS_EPI_WARPS) {
    // ======================================================================== epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_empty_leader[2] = {mma::map_to_cta(&t_empty[0], 0), mma::map_to_cta(&t_empty[1], 0)};
    uint32_t tc = 0, items_done = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int mt = 2 * (item % n_mpairs) + (int)rank, jgl = item / n_mpairs;
      const int j0 = (a.jg0 + jgl) * a.JT;
      int64_t e_lo;
      int n_tiles;
      nls_item_range(a, j0, e_lo, n_tiles);
      if (n_tiles == 0) continue;
      const int m = mt * 128 + row;
      const bool mok = m < a.M;
      float zr[DP];
#pragma unroll
      for (int i = 0; i < DP; ++i) zr[i] = mok ? a.zz[(int64_t)m * DP + i] : 0.f;
      const float bm = mok ? a.bz[m] : 0.f;
      const float szm = mok ? a.szf[m] : 0.f;
      const float ainv = a.ainv[mt * 128 + row];
      double acc[NLS_JT];
#pragma unroll
      for (int jj = 0; jj < NLS_JT; ++jj) acc[jj] = 0.0;
      for (int t = 0; t < n_tiles; ++t, ++tc) {
        const uint32_t buf = tc & 1u, ph = (tc >> 1) & 1u;
        mma::mbar_wait(r_full, tc & 1u);
        mma::mbar_wait(&t_full[buf], ph);
        mma::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NT + half * (NT / 2);
        nls_epilogue_tile<FAM, DP, NL, 0>(a, sRec, taddr, half, zr, bm, szm, ainv, acc, m, mok, j0);
        mma::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mma::mbar_arrive_cluster(t_empty_leader[buf]);
          mma::mbar_arrive(r_empty);
        }
      }
      double* comb = sComb + (items_done & 1u) * (128 * NLS_JT);
      if (half == 1) {
#pragma unroll
        for (int jj = 0; jj < NLS_JT; ++jj) comb[jj * 128 + row] = acc[jj];
      }
      mma::named_bar_sync(1, NLS_EPI_WARPS * 32);
      if (half == 0 && mok) {
        double* dst = a.G + (int64_t)m * a.ldg + j0;
#pragma unroll
        for (int jj = 0; jj < NLS_JT; ++jj)
          if (j0 + jj < a.S) dst[jj] = acc[jj] + comb[jj * 128 + row];
      }
      ++items_done;
    }
  } else if (warp == NLS_EPI_WARPS) {
    // ======================================================================== operand producer
    if (lane == 0) {
      uint32_t it = 0;
      const size_t a_tile = (size_t)(a.KP / 8) * 128 * 8;   // halves per landmark tile
      const size_t b_half = (size_t)(a.KP / 8) * 128 * 8;   // halves per half candidate tile
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int mt = 2 * (item % n_mpairs) + (int)rank, jgl = item / n_mpairs;
        int64_t e_lo;
        int n_tiles;
        nls_item_range(a, (a.jg0 + jgl) * a.JT, e_lo, n_tiles);
        for (int t = 0; t < n_tiles; ++t) {
          const size_t tile = (size_t)jgl * a.tiles_per_jg + t;
          const __half* bh = a.kxh + (2 * tile + rank) * b_half;
          const __half* bl = a.kxl + (2 * tile + rank) * b_half;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % NSTAGE;
            mma::mbar_wait(&s_empty[s], ((it / NSTAGE) & 1u) ^ 1u);
            unsigned char* st = smem + Cfg::OFF_STAGE + (size_t)s * NLS2_STAGE_BYTES;
            mma::mbar_expect_tx(&s_full[s], NLS2_STAGE_BYTES);
            const size_t ko = (size_t)kb * (NLS2_PIECE / 2);
            mma::bulk_g2s(st, a.azh + mt * a_tile + ko, NLS2_PIECE, &s_full[s]);
            mma::bulk_g2s(st + NLS2_PIECE, a.azl + mt * a_tile + ko, NLS2_PIECE, &s_full[s]);
            mma::bulk_g2s(st + 2 * NLS2_PIECE, bh + ko, NLS2_PIECE, &s_full[s]);
            mma::bulk_g2s(st + 3 * NLS2_PIECE, bl + ko, NLS2_PIECE, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == NLS_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer (rank 0) / relay (rank 1)
    // instruction descriptor: D fp32, A / B fp16 K-major, N = 256, M = 256 (128 rows per CTA)
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((256u >> 4) << 24);
    uint32_t it = 0, tc = 0;
    uint32_t p_full_leader[NSTAGE];
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) p_full_leader[s] = mma::map_to_cta(&p_full[s], 0);
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int jgl = item / n_mpairs;
      int64_t e_lo;
      int n_tiles;
      nls_item_range(a, (a.jg0 + jgl) * a.JT, e_lo, n_tiles);
      for (int t = 0; t < n_tiles; ++t, ++tc) {
        const uint32_t buf = tc & 1u;
        if (rank == 0) {
          mma::mbar_wait(&t_empty[buf], ((tc >> 1) & 1u) ^ 1u);
          mma::tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1u;
          mma::mbar_wait(&s_full[s], ph);
          if (rank != 0) {
            if (lane == 0) mma::mbar_arrive_cluster(p_full_leader[s]);
            __syncwarp();
            continue;
          }
          mma::mbar_wait(&p_full[s], ph);
          mma::tc_fence_after();
          if (lane == 0) {
            const uint32_t st = mma::smem_u32(smem + Cfg::OFF_STAGE + (size_t)s * NLS2_STAGE_BYTES);
            const uint32_t ahi = st, alo = st + NLS2_PIECE, bhi = st + 2 * NLS2_PIECE, blo = st + 3 * NLS2_PIECE;
            const uint32_t d = tmem_base + buf * NT;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              const uint32_t aa = (p == 0) ? alo : ahi;   // lo*hi, hi*lo, hi*hi (small terms first)
              const uint32_t bb = (p == 1) ? blo : bhi;
#pragma unroll
              for (int ks = 0; ks < NLS_KB / 16; ++ks) {
                const uint64_t ad = mma::smem_desc(aa + ks * 2 * (128 * 16), 128 * 16, 128);
                const uint64_t bd = mma::smem_desc(bb + ks * 2 * (128 * 16), 128 * 16, 128);
                mma::umma2_f16(d, ad, bd, IDESC, (kb > 0 || p > 0 || ks > 0) ? 1u : 0u);
              }
            }
            mma::umma_commit2(&s_empty[s], 3);
            if (kb == nkb - 1) mma::umma_commit2(&t_full[buf], 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================================================================== record producer
    if (lane == 0) {
      uint32_t tc = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int jgl = item / n_mpairs;
        int64_t e_lo;
        int n_tiles;
        nls_item_range(a, (a.jg0 + jgl) * a.JT, e_lo, n_tiles);
        for (int t = 0; t < n_tiles; ++t, ++tc) {
          mma::mbar_wait(r_empty, (tc & 1u) ^ 1u);
          const size_t tile = (size_t)jgl * a.tiles_per_jg + t;
          mma::mbar_expect_tx(r_full, Cfg::REC_BYTES);
          mma::bulk_g2s(sRec, a.trec + tile * Cfg::REC_BYTES, Cfg::REC_BYTES, r_full);
        }
      }
    }
  }

  mma::tc_fence_before();
  mma::cluster_sync_all();   // the peer's shared memory and TMEM stay valid until both CTAs are done
  if (warp == NLS_EPI_WARPS + 1) {
    mma::tc_fence_after();
    mma::tmem_dealloc2(tmem_base, 512);
  }
}

template <int FAM, int DP, int NL>
int launch_nlsum2_dp(basq_ctx* ctx, const NlsDev& dev) {
  using Cfg = Nls2Cfg<DP>;
  BASQ_CHECK((size_t)Cfg::SMEM_BYTES <= ctx->smem_optin, BASQ_ERR_UNSUPPORTED,
             "nlsum2: kernel needs %d B shared memory (limit %zu)", Cfg::SMEM_BYTES, ctx->smem_optin);
  BASQ_CHECK(dev.n_mtiles % 2 == 0, BASQ_ERR_INVALID, "nlsum2: the landmark tiles must be padded to an even count");
  const int64_t n_items = (int64_t)(dev.n_mtiles / 2) * dev.n_jg;
  int grid = (int)std::min<int64_t>(ctx->num_sms / 2, n_items) * 2;
  if (grid <= 0) return BASQ_OK;
  BASQ_CUDA(cudaFuncSetAttribute(nlsum2_kernel<FAM, DP, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  nlsum2_kernel<FAM, DP, NL><<<grid, NLS_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(dev);   // __cluster_dims__(2)
  ctx->launches++;
  BASQ_CUDA(cudaGetLastError());
  return BASQ_OK;
}

template <int FAM, int NL>
int launch_nlsum2_family(basq_ctx* ctx, int dp, const NlsDev& dev) {
  switch (dp) {
    case 2: return launch_nlsum2_dp<FAM, 2, NL>(ctx, dev);
    case 4: return launch_nlsum2_dp<FAM, 4, NL>(ctx, dev);
    case 6: return launch_nlsum2_dp<FAM, 6, NL>(ctx, dev);
    case 8: return launch_nlsum2_dp<FAM, 8, NL>(ctx, dev);
    case 10: return launch_nlsum2_dp<FAM, 10, NL>(ctx, dev);
    case 12: return launch_nlsum2_dp<FAM, 12, NL>(ctx, dev);
    case 16: return launch_nlsum2_dp<FAM, 16, NL>(ctx, dev);
    case 20: return launch_nlsum2_dp<FAM, 20, NL>(ctx, dev);
    case 24: return launch_nlsum2_dp<FAM, 24, NL>(ctx, dev);
    case 32: return launch_nlsum2_dp<FAM, 32, NL>(ctx, dev);
  }
  set_error("nlsum2: no kernel compiled for padded dimension %d", dp);
  return BASQ_ERR_UNSUPPORTED;
}

int launch_nlsum2(basq_ctx* ctx, int fam, int nl, int dp, const NlsDev& dev);
int launch_nlsum2_rbf_wm(basq_ctx*, int, const NlsDev&);
int launch_nlsum2_rbf_ml(basq_ctx*, int, const NlsDev&);
int launch_nlsum2_m15_wm(basq_ctx*, int, const NlsDev&);
int launch_nlsum2_m15_ml(basq_ctx*, int, const NlsDev&);
int launch_nlsum2_m25_wm(basq_ctx*, int, const NlsDev&);
int launch_nlsum2_m25_ml(basq_ctx*, int, const NlsDev&);

}  // namespace basq
