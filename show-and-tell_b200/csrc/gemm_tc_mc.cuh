// EXPERIMENTAL (written at the end of round 1 without a GPU run; opt-in through SNT_GEMM_MC=1).
//
// Cluster-multicast variant of the contraction core in gemm_tc.cuh: the CL (2 or 4) CTAs of a (CL,1,1) cluster work on CL
// row tiles of the SAME column tile n, so they need the same B tile.  Each CTA fetches its own A tile and 1/CL of the
// B tile; the B parts are TMA-multicast into every CTA's shared memory.  Per k-block a CTA then pulls
// 16 KB (A) + BN/CL x 128 B from L2 instead of 16 KB + BN x 128 B: 32 KB (CL = 2) or 24 KB (CL = 4) instead of 48 KB at
// BN = 256.
//
// Why (a hypothesis to measure, tools/gemm_step_shapes.py): at 128 x 256 x 64 tiles a CTA requests 48 KB per 4.2 MFLOP; 148
// SMs at ~1000 TFLOP/s ask L2 for ~11.4 TB/s, about what its slices deliver chip-wide, and cuBLAS uses 2-SM tiles for
// that reason.  Against it: the same core reaches 1337 TFLOP/s on 8192^3, where L2 de-duplicates the tiles many CTAs of a
// wave request together - the short-K vocabulary contractions may be epilogue-bound instead.
//
// Everything else - tile shape, swizzle, descriptors, the epilogue functors, the double-buffered accumulator - is the
// validated code of gemm_tc.cuh, used unchanged.  The pipeline differences, all mirrored from the cluster mode of the
// persistent LSTM kernel (lstm_tc.cu, validated there with clusters of 4):
//   * empty[stage] counts 2 arrivals: the MMA issuers of BOTH CTAs release a stage (tcgen05.commit ... multicast), because
//     the peer's multicast writes into this CTA's stage too;
//   * full[stage] still expects STAGE_BYTES: own A + own B half + the peer's B half land on this CTA's barrier;
//   * cluster barrier after the mbarrier initialisation (no multicast / remote arrive may hit an uninitialised barrier) and
//     before exit (no CTA may leave while its peer can still write into it);
//   * tiles are handed out to clusters in pairs of row tiles; an odd last row tile is paired with an out-of-range one,
//     whose A box is zero-filled by TMA and whose epilogue is skipped.
//
// Desk checks done without a GPU: tools/mc_pipeline_sim.py replays this barrier protocol (producer / MMA issuer / epilogue
// of CL CTAs, random latencies) and asserts deadlock freedom, that no stage is overwritten while a consumer may still
// read it, and that every consumer sees exactly its round's bytes; the tile enumeration was checked to cover every
// (m, n, split) exactly once; tools/sass_diff.py shows the validated kernels' SASS untouched.
#pragma once
#include "gemm_tc.cuh"

namespace snt {
namespace tc {

template <int BN, bool A_MN, bool B_MN, class Epi, int CL = 2>
__global__ void __launch_bounds__(128 + 32 * Epi::kWarps, 1)
gemm_tc_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const TileSched ts, const Epi epi) {
  static_assert(CL == 2 || CL == 4, "clusters of 2 or 4 row tiles");
  static_assert(BN / CL >= 64, "the B tile is split into CL parts of whole 64-row boxes");
  using C = Cfg<BN, Epi::kStages>;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* epi_smem = smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CL);  // released by the MMA issuer of every CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], Epi::kWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();  // the peer's barriers are initialised before any multicast traffic or remote arrive
  const uint32_t tmem_base = *tmem_slot;
  const int crank = (int)cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x / CL), n_clusters = (int)(gridDim.x / CL);

  TileSched tp = ts;  // the same enumeration over GROUPS of CL row tiles ("pairs" below, for CL = 2)
  tp.num_m = (ts.num_m + CL - 1) / CL;
  const int total_pairs = tp.num_m * tp.num_n * tp.splits;

  if (warp == 0) {
    if (elect_one()) {
      // ================= TMA producer =================
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int pair = cluster_id; pair < total_pairs; pair += n_clusters) {
        int mp, n_blk, split;
        decode_tile(tp, pair, mp, n_blk, split);
        const int m_blk = CL * mp + crank;  // may be >= ts.num_m (ragged last group): TMA zero-fills that A box
        const int kb0 = split * ts.kblocks_per_split;
        const int kb1 = min(kb0 + ts.kblocks_per_split, ts.kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);  // both CTAs have finished reading this stage
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);  // own A + own B part + the peers' B parts
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full[stage], kb * BK, ts.a_row0 + m_blk * BM);
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h)
              tma_load_2d(sA + h * (BK * 128), &tmA, &full[stage], ts.a_row0 + m_blk * BM + h * 64, kb * BK);
          }
          if (!B_MN) {
            // tmB's box is {BK, BN/CL}: rows [crank*BN/CL, +BN/CL) of the B tile, 128 swizzled bytes per row
            tma_load_2d_mc(sB + crank * (BN / CL) * 128, &tmB, &full[stage], kb * BK,
                           ts.b_row0 + n_blk * BN + crank * (BN / CL), CMASK);
          } else {
#pragma unroll
            for (int q = 0; q < BN / (64 * CL); ++q) {
              const int h = crank * (BN / (64 * CL)) + q;
              tma_load_2d_mc(sB + h * (BK * 128), &tmB, &full[stage], ts.b_row0 + n_blk * BN + h * 64, kb * BK, CMASK);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pair = cluster_id; pair < total_pairs; pair += n_clusters) {
        const int split = pair / (tp.num_m * tp.num_n);
        const int kb0 = split * ts.kblocks_per_split;
        const int kb1 = min(kb0 + ts.kblocks_per_split, ts.kblocks);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(a_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(b_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_mc(&empty[stage], CMASK);  // one arrival on this stage's empty barrier in EVERY CTA of the cluster
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    pdl_wait();
    for (int pair = cluster_id; pair < total_pairs; pair += n_clusters) {
      int mp, n_blk, split;
      decode_tile(tp, pair, mp, n_blk, split);
      const int m_blk = CL * mp + crank;
      const bool valid = m_blk < ts.num_m;  // warp-uniform
      typename Epi::Pre pre;
      if (valid) epi.prefetch(pre, m_blk, n_blk, ew, lane);
      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t tmem_rows = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)((ew & 3) * 32) << 16);
      if (valid) epi.tile(tmem_rows, m_blk, n_blk, split, ew, lane, pre, epi_smem + ew * Epi::kSmemPerWarp);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// tmB must have been made with box rows BN/CL when B is K-major (make_operand_tmap(..., BN / CL)); MN-major B uses the
// ordinary map (its boxes are 64 rows already).
template <int BN, bool A_MN, bool B_MN, class Epi, int CL = 2>
int launch_gemm_tc_mc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TileSched& ts, const Epi& epi,
                      cudaStream_t st) {
  using C = Cfg<BN, Epi::kStages>;
  auto kern = gemm_tc_mc_kernel<BN, A_MN, B_MN, Epi, CL>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    SNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES + Epi::kWarps * Epi::kSmemPerWarp));
    configured = true;
  }
  const int pairs = ((ts.num_m + CL - 1) / CL) * ts.num_n * ts.splits;
  if (pairs <= 0) return SNT_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(128 + 32 * Epi::kWarps);
  cfg.dynamicSmemBytes = C::SMEM_BYTES + Epi::kWarps * Epi::kSmemPerWarp;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // pairs of SMs of one GPC that can hold a cluster at the same time (a GPC with an odd SM count leaves one SM out):
  // the persistent grid is sized to what is co-resident, like grid_sms() does for single CTAs
  static int resident = 0;  // per instantiation
  if (resident == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = sm_count() / CL;
    }
    resident = n;
  }
  int clusters = min(pairs, min(resident, grid_sms() / CL));
  if (clusters <= 0) { set_error("launch_gemm_tc_mc: no SM group available"); return SNT_EINVAL; }
  cfg.gridDim = dim3((unsigned)(CL * clusters));
  count_launch();
  SNT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ts, epi));
  return SNT_OK;
}

// SNT_GEMM_MC: unset / 0 = off, 1 or 2 = clusters of 2, 4 = clusters of 4 (only where a 4-way instantiation exists)
inline int mc_cluster() {
  const char* e = getenv("SNT_GEMM_MC");
  if (!e) return 0;
  if (e[0] == '1' || e[0] == '2') return 2;
  if (e[0] == '4') return 4;
  return 0;
}
inline bool mc_enabled() { return mc_cluster() != 0; }

}  // namespace tc
}  // namespace snt
