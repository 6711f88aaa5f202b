// tcgen05 / TMEM / TMA contraction core for sm_100a.
//
//   D[M,N] (fp32, TMEM) = A . B,  bf16 operands, 128 x BN x 64 tiles, persistent CTAs, warp-specialised:
//     warp 0  : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled shared memory, mbarrier complete_tx)
//     warp 1  : MMA issuer     (one elected thread issues tcgen05.mma, tcgen05.commit frees the stage)
//     warp 2  : TMEM allocator
//     warps 4+: epilogue       (tcgen05.ld -> registers -> functor), double-buffered accumulator so the
//                               epilogue of tile i overlaps the MMAs of tile i+1
//   Operand layouts: "K-major" (the contraction index is contiguous in global memory) or "MN-major" (the M or N
//   index is contiguous); both are fetched with 128B-swizzle TMA boxes and described to the tensor core with
//   the matching shared-memory descriptors, so x.W^T, dY.W and dY^T.X all run without a transpose pass.
//   The epilogue is a functor (plain store, LSTM gates, cross-entropy statistics, ...) given the TMEM address
//   of its warp's 32 accumulator rows.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace snt {
namespace tc {

constexpr int BM = 128;  // tile rows = TMEM lanes
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle row

struct TileSched {
  int num_m, num_n, splits;  // tile grid; tile id = (split * num_n + n) * num_m + m, or with n_fastest
                             // (split * num_m + m) * num_n + n: the CTAs of one wave then share few A row panels (each is
                             // fetched from DRAM once and served to its num_n column tiles out of L2)
  int n_fastest;
  int kblocks;               // total K blocks of BK
  int kblocks_per_split;
  int a_row0, b_row0;        // coordinate offsets into the M / N extents of the tensor maps
};

__device__ __forceinline__ void decode_tile(const TileSched& ts, int tile, int& m_blk, int& n_blk, int& split) {
  if (ts.n_fastest) {
    n_blk = tile % ts.num_n;
    const int rest = tile / ts.num_n;
    m_blk = rest % ts.num_m;
    split = rest / ts.num_m;
  } else {
    m_blk = tile % ts.num_m;
    const int rest = tile / ts.num_m;
    n_blk = rest % ts.num_n;
    split = rest / ts.num_n;
  }
}

template <int BN, int STAGES_OVERRIDE = 0>
struct Cfg {
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN must be 64, 128 or 256");
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_OVERRIDE ? STAGES_OVERRIDE : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (128, 256 or 512 columns)
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES;
};

// Per-warp 32 x 64-byte transpose stage of the epilogues, XOR-swizzled in 16-byte pieces: conflict-free both for the
// row-per-lane writes (lane r writes piece q of row r: a quarter warp covers 8 rows) and for the reads of the store
// phase (a quarter warp reads the 4 pieces of 2 consecutive rows).  ncu on the padded 80-byte pitch it replaces showed
// 2.5 shared-memory wavefronts per ideal one on the reads.
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int row, int piece) {
  return base + (uint32_t)(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
}

using ::snt::pdl_wait;
using ::snt::pdl_launch_dependents;
using ::snt::pdl_all;
using ::snt::launch_chained;

// Epi must provide:
//   static constexpr int kWarps            (4, 8 or 16 epilogue warps)
//   static constexpr int kStages           (TMA ring depth; 0 = the default for BN)
//   struct Pre; __device__ void prefetch(Pre&, int m_blk, int n_blk, int epi_warp, int lane) const
//       global loads the epilogue will need, issued BEFORE waiting for the accumulator so that their latency
//       overlaps the TMA/MMA phase (NoPre = nothing to prefetch)
//   static constexpr int kSmemPerWarp   bytes of shared scratch each epilogue warp gets (16-byte aligned), may be 0
//   __device__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int split, int epi_warp, int lane, const Pre&,
//                        uint8_t* warp_smem) const
// where tmem_rows addresses lane 32*(epi_warp%4), first column of this tile's accumulator.
struct NoPre {};
template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(128 + 32 * Epi::kWarps, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TileSched ts, const Epi epi) {
  using C = Cfg<BN, Epi::kStages>;
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
      mbar_init(&empty[i], 1);
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
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // the next kernel may begin its own prologue; its griddepcontrol.wait still
                            // blocks until this grid has completed and flushed

  const int total_tiles = ts.num_m * ts.num_n * ts.splits;

  if (warp == 0) {
    if (elect_one()) {
      // ================= TMA producer =================
      pdl_wait();  // operands may be produced by the preceding kernel in the stream
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, split;
        decode_tile(ts, tile, m_blk, n_blk, split);
        const int kb0 = split * ts.kblocks_per_split;
        const int kb1 = min(kb0 + ts.kblocks_per_split, ts.kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full[stage], kb * BK, ts.a_row0 + m_blk * BM);
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h)
              tma_load_2d(sA + h * (BK * 128), &tmA, &full[stage], ts.a_row0 + m_blk * BM + h * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full[stage], kb * BK, ts.b_row0 + n_blk * BN);
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sB + h * (BK * 128), &tmB, &full[stage], ts.b_row0 + n_blk * BN + h * 64, kb * BK);
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (ts.num_m * ts.num_n);
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
            // K-major: 16 bf16 = 32 B along the swizzled row.  MN-major: 16 k-rows of 128 B = 2048 B.
            const uint64_t da = A_MN ? make_smem_desc(a_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(b_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    pdl_wait();  // the epilogue may read tensors written by the preceding kernel (C for beta, Gx, lse, ...)
    // The global loads an epilogue needs (Epi::prefetch) are issued one tile AHEAD: when the kernel is epilogue-bound the
    // accumulator is already complete at the wait below, and loads issued right before it would be fully exposed.
    typename Epi::Pre pre, pre_next;
    if ((int)blockIdx.x < total_tiles) {
      int m0, n0, s0;
      decode_tile(ts, blockIdx.x, m0, n0, s0);
      epi.prefetch(pre, m0, n0, ew, lane);
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk, split;
      decode_tile(ts, tile, m_blk, n_blk, split);
      const int next = tile + (int)gridDim.x;
      if (next < total_tiles) {
        int m1, n1, s1;
        decode_tile(ts, next, m1, n1, s1);
        epi.prefetch(pre_next, m1, n1, ew, lane);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t tmem_rows = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)((ew & 3) * 32) << 16);
      epi.tile(tmem_rows, m_blk, n_blk, split, ew, lane, pre, epi_smem + ew * Epi::kSmemPerWarp);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      pre = pre_next;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
// 2-D bf16 tensor map with 128B swizzle: `inner` contiguous elements per row, `outer` rows, row pitch ld elements.
int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                   int box_outer);
// operand map for the kernel above: mn_major=false -> [rows, K] K contiguous, box {BK, box_rows};
// mn_major=true -> [K, rows] rows contiguous, box {64, BK}.
int make_operand_tmap(CUtensorMap* out, const void* base, bool mn_major, int64_t rows, int64_t K, int64_t ld,
                      int box_rows);
int sm_count();
int grid_sms();  // SMs a persistent grid may occupy: sm_count() minus the reserve set through snt_set_sm_reserve()
// Caps the persistent grids launched by THIS thread until reset with 0 (returns the previous cap): a contraction that runs
// beside a cooperative kernel is sized for the SMs that kernel leaves free.
int set_grid_cap(int n);

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TileSched& ts, const Epi& epi,
                   cudaStream_t st, bool pdl = false) {
  using C = Cfg<BN, Epi::kStages>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, Epi>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    SNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES + Epi::kWarps * Epi::kSmemPerWarp));
    configured = true;
  }
  const int total = ts.num_m * ts.num_n * ts.splits;
  if (total <= 0) return SNT_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)min(total, grid_sms()));
  cfg.blockDim = dim3(128 + 32 * Epi::kWarps);
  cfg.dynamicSmemBytes = C::SMEM_BYTES + Epi::kWarps * Epi::kSmemPerWarp;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl || pdl_all()) ? 1 : 0;
  count_launch();
  SNT_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ts, epi));
  return SNT_OK;
}

// ---- the plain epilogue: C = alpha*acc + bias[col] + beta*C, fp32 and/or bf16 out, optional split-K slices ----
// WIDE (BN = 256 only): 16 epilogue warps and a 3-stage ring, for short contractions whose tiles finish their MMAs faster
// than 8 warps can drain 128 x 256 outputs (e.g. the input projection Gx' with K = 256).
template <int BN, bool WIDE = false>
struct PlainEpi {
  static_assert(!WIDE || BN == 256, "the wide epilogue is a BN = 256 variant");
  static constexpr int kWarps = WIDE ? 16 : (BN >= 128 ? 8 : 4);
  static constexpr int kStages = WIDE ? 3 : 0;
  static constexpr int kSmemPerWarp = 32 * 64;  // swizzled 32 x 64-byte transpose stage (stage_addr)
  int M, N;                // valid extent
  float alpha, beta;
  const float* alpha_dev;  // optional device scalar multiplied into alpha (e.g. the incoming dloss)
  float* C;                // may be NULL
  __nv_bfloat16* Cb;       // may be NULL
  int64_t ldc;
  const float* bias;       // may be NULL, length N
  int64_t split_stride;    // elements between split-K partial slices of C
  int row_perm_h;          // != 0: accumulator row 4*j+g is stored to row g*row_perm_h + j (LSTM gate un-interleave)
  const float* row_scale;  // optional [M]: row m of alpha*acc is multiplied by row_scale[m] (before bias / beta)

  using Pre = NoPre;
  __device__ __forceinline__ void prefetch(Pre&, int, int, int, int) const {}
  // Accumulator rows live one per lane, so a direct store makes every instruction touch 32 different rows (16 bytes
  // each).  Full 32-column chunks are instead transposed through a per-warp shared-memory stage (row pitch padded
  // against bank conflicts) and written as whole 64-byte row segments (two rounds per chunk for fp32); the beta read of
  // C happens in the same coalesced pattern.
  struct Ctx {
    int row0, n_blk, split, lane;
    float a;
    bool vec_ok, bvec_ok;
    uint32_t wsa;
  };
  // one 32-column chunk of this warp's 32 accumulator rows
  __device__ __forceinline__ void chunk(const Ctx& x, const uint32_t (&r)[32], int c) const {
    const int col0 = x.n_blk * BN + c * 32;
    if (col0 >= N) return;  // warp-uniform
    const int lane = x.lane;
    const float a = x.a;
    const bool full = col0 + 32 <= N;
    float4 bv[8];
    if (bias && full && (x.vec_ok || x.bvec_ok)) {
#pragma unroll
      for (int q = 0; q < 8; ++q) bv[q] = __ldg(reinterpret_cast<const float4*>(bias + col0) + q);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) bv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (x.vec_ok && full) {
      float* cbase = C + (int64_t)x.split * split_stride + col0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // 16 columns (64 bytes per row) per round
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = hf * 16 + q * 4;
          const float4 b4 = bv[hf * 4 + q];
          sts128(stage_addr(x.wsa, lane, q), __float_as_uint(fmaf(a, __uint_as_float(r[j]), b4.x)),
                 __float_as_uint(fmaf(a, __uint_as_float(r[j + 1]), b4.y)),
                 __float_as_uint(fmaf(a, __uint_as_float(r[j + 2]), b4.z)),
                 __float_as_uint(fmaf(a, __uint_as_float(r[j + 3]), b4.w)));
        }
        __syncwarp();
        float4* dst[4];
        float4 o[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int grow = x.row0 + it * 8 + (lane >> 2);
          const int dr = row_perm_h ? (grow & 3) * row_perm_h + (grow >> 2) : grow;
          dst[it] = grow < M ? reinterpret_cast<float4*>(cbase + (int64_t)dr * ldc + hf * 16) + (lane & 3) : nullptr;
        }
        if (beta != 0.f) {  // all loads first: a load placed after a (possibly aliasing) store cannot be hoisted
#pragma unroll
          for (int it = 0; it < 4; ++it) o[it] = dst[it] ? __ldcg(dst[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const uint4 u = lds128(stage_addr(x.wsa, it * 8 + (lane >> 2), lane & 3));
          float4 v = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
          if (beta != 0.f) {
            v.x += beta * o[it].x; v.y += beta * o[it].y; v.z += beta * o[it].z; v.w += beta * o[it].w;
          }
          if (dst[it]) *dst[it] = v;
        }
        __syncwarp();
      }
    } else if (x.bvec_ok && full) {  // bf16-only output
      const float bf_[32] = {bv[0].x, bv[0].y, bv[0].z, bv[0].w, bv[1].x, bv[1].y, bv[1].z, bv[1].w,
                             bv[2].x, bv[2].y, bv[2].z, bv[2].w, bv[3].x, bv[3].y, bv[3].z, bv[3].w,
                             bv[4].x, bv[4].y, bv[4].z, bv[4].w, bv[5].x, bv[5].y, bv[5].z, bv[5].w,
                             bv[6].x, bv[6].y, bv[6].z, bv[6].w, bv[7].x, bv[7].y, bv[7].z, bv[7].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int j = q * 8 + h * 2;
          __nv_bfloat162 t = __floats2bfloat162_rn(fmaf(a, __uint_as_float(r[j]), bf_[j]),
                                                   fmaf(a, __uint_as_float(r[j + 1]), bf_[j + 1]));
          pk[h] = *reinterpret_cast<uint32_t*>(&t);
        }
        sts128(stage_addr(x.wsa, lane, q), pk[0], pk[1], pk[2], pk[3]);
      }
      __syncwarp();
      uint4 v[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) v[it] = lds128(stage_addr(x.wsa, it * 8 + (lane >> 2), lane & 3));
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int grow = x.row0 + it * 8 + (lane >> 2), cq = lane & 3;
        if (grow < M) {
          const int dr = row_perm_h ? (grow & 3) * row_perm_h + (grow >> 2) : grow;
          *reinterpret_cast<uint4*>(Cb + (int64_t)dr * ldc + col0 + cq * 8) = v[it];
        }
      }
      __syncwarp();
    } else {
      const int row = x.row0 + lane;
      if (row < M) {
        const int drow = row_perm_h ? (row & 3) * row_perm_h + (row >> 2) : row;
        float* crow = C ? C + (int64_t)x.split * split_stride + (int64_t)drow * ldc : nullptr;
        __nv_bfloat16* brow = Cb ? Cb + (int64_t)drow * ldc : nullptr;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < N) {
            float v = a * __uint_as_float(r[j]);
            if (bias) v += bias[col];
            if (crow) {
              if (beta != 0.f) v += beta * crow[col];
              crow[col] = v;
            }
            if (brow) brow[col] = __float2bfloat16_rn(v);
          }
        }
      }
    }
  }
  // Each epilogue warp owns kChunks consecutive 32-column chunks of its lane quadrant; the TMEM load of chunk c+1 is in
  // flight while chunk c is converted, staged and stored (two register buffers, loop unrolled by two).
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int split, int ew, int lane,
                                       const Pre&, uint8_t* wsm) const {
    Ctx x;
    x.row0 = m_blk * BM + (ew & 3) * 32;
    x.n_blk = n_blk; x.split = split; x.lane = lane;
    x.a = alpha_dev ? alpha * alpha_dev[0] : alpha;
    if (row_scale) x.a *= (x.row0 + lane < M) ? __ldg(row_scale + x.row0 + lane) : 0.f;  // one accumulator row per lane
    const bool bias_al = !bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
    x.vec_ok = C && !Cb && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && bias_al &&
               ((split_stride & 3) == 0);
    x.bvec_ok = !C && Cb && ((ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0) && bias_al;
    x.wsa = smem_u32(wsm);
    constexpr int kChunks = BN / 32 / (kWarps / 4);
    static_assert(kChunks % 2 == 0, "chunk loop is unrolled by two");
    const int c0 = (ew >> 2) * kChunks;
    if (WIDE) {  // 16 warps hide the latencies by themselves: one register buffer, no software pipelining
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_rows + (uint32_t)((c0 + c) * 32), r);
        tmem_ld_wait();
        chunk(x, r, c0 + c);
      }
      return;
    }
    uint32_t ra[32], rb[32];
    tmem_ld32(tmem_rows + (uint32_t)(c0 * 32), ra);
#pragma unroll 1
    for (int c = 0; c < kChunks; c += 2) {
      tmem_ld_wait();
      tmem_ld32(tmem_rows + (uint32_t)((c0 + c + 1) * 32), rb);
      chunk(x, ra, c0 + c);
      tmem_ld_wait();
      if (c + 2 < kChunks) tmem_ld32(tmem_rows + (uint32_t)((c0 + c + 2) * 32), ra);
      chunk(x, rb, c0 + c + 1);
    }
  }
};

// Generic bf16 contraction, C[M,N] = alpha * op(A).op(B) + bias + beta*C.
//   a_mn = false: A is [M,K] row-major (lda);  true: A is [K,M] row-major.
//   b_mn = false: B is [N,K] row-major (ldb);  true: B is [K,N] row-major.
//   C fp32 and/or Cb bf16 (same ldc).  splits > 1: partial sums go to split_ws[splits][M][ldc] and are reduced
//   (deterministically) into C by a second kernel; bias/beta are applied there.
int64_t gemm_tc_split_ws_elems(int64_t M, int64_t ldc, int splits);
int choose_splits(int64_t M, int64_t N, int64_t K, int bn);
int gemm_tc(bool a_mn, bool b_mn, int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A,
            int64_t lda, const __nv_bfloat16* B, int64_t ldb, float beta, float* C, __nv_bfloat16* Cb,
            int64_t ldc, const float* bias, int splits, float* split_ws, cudaStream_t st, int row_perm_h = 0,
            const float* alpha_dev = nullptr, bool keep_partials = false, int* splits_used = nullptr,
            int force_bn = 0, int n_fastest = 0, const float* row_scale = nullptr);

// C[M,N] (fp32) = alpha * diag(row_scale) . op(A).op(B), wave-balanced: whole waves of unsplit tiles plus a K-split tail
// (see gemm_tc.cu).  split_ws_elems floats of scratch for the tail's partial sums.
int gemm_tc_balanced(bool a_mn, bool b_mn, int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A,
                     int64_t lda, const __nv_bfloat16* B, int64_t ldb, float* C, int64_t ldc, float* split_ws,
                     int64_t split_ws_elems, cudaStream_t st, const float* alpha_dev, int bn,
                     const float* row_scale = nullptr);

}  // namespace tc
}  // namespace snt
