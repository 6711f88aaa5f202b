// a7 / a10 on tensor cores: the LSTM recurrence with the gate math fused into the contraction's epilogue.
//
// Layout trick: gate rows of W_ih / W_hh / bias are interleaved per hidden unit, r' = 4*j + g (g = i,f,g,o), so any
// accumulator tile holds all four gates of its hidden units and one epilogue thread (= one batch row) can finish
// c_t and h_t straight out of TMEM.  Pre-activations never reach HBM:
//   forward step t : TMEM = h_{t-1}[bs_t,H] . W_hh'^T ; epilogue adds the batched input projection Gx'[t]
//                    (one tcgen05 GEMM over all timesteps), applies sigma/tanh, writes c_t, h_t (bf16, also as next
//                    step's A operand) and the bf16 activations kept for BPTT.
//   backward step t: dG'_{t+1}[bs_{t+1},4H] . W_hh' as a split-K contraction, then a coalesced pointwise kernel that sums
//                    the partials, adds dL/dh_t, runs the cell backward and writes dG'_t as bf16 — the A operand of
//                    step t-1 and of the weight-gradient GEMMs.  (A variant with the cell backward fused into the GEMM
//                    epilogue was measured slower at both B=1024/H=512 and B=2048/H=1024 and removed.)
// Forward steps are chained with programmatic dependent launch, so each step's prologue overlaps its predecessor's tail.
// At H <= 512 both directions run as ONE persistent cooperative kernel each (further down).
#include "bf16.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

#include <stdlib.h>
#include <string.h>

namespace snt {
namespace bf16 {

typedef __nv_bfloat16 bf;
constexpr int MAX_SPLITS = 16;

// Gate non-linearities on the SFU: ONE MUFU.TANH per value (tanh.approx.f32, max relative error 2^-11 — below the
// bf16 rounding (2^-9) applied to every activation that leaves the epilogue).  The exp2 + reciprocal formulation costs
// two MUFU ops per value and made the 10-transcendental-per-unit gate epilogue MUFU-bound (16 lanes/clk/SM).
__device__ __forceinline__ float tanh_(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// ---- weight preparation lives in lstm_prep_kernel (below) ----------------------------------------------------------
__global__ void unperm_vec_kernel(const float* __restrict__ in, int H, float* __restrict__ out) {
  const int rp = blockIdx.x * blockDim.x + threadIdx.x;
  if (rp >= 4 * H) return;
  out[(rp & 3) * H + (rp >> 2)] = in[rp];
}

// ---- forward step epilogue ---------------------------------------------------------------------------------------------
struct LstmFwdEpi {
  // 16 epilogue warps: warp -> (TMEM lane quadrant ew & 3, 32-column chunk ew >> 2), i.e. 8 hidden units of 32 rows each.
  // (With 4 warps - one per quadrant, four chunks each, 243 registers - the gate math ran one warp per scheduler with
  // nothing to hide its MUFU / load latencies behind: ncu showed the tensor pipe 28 % active in the greedy step.)
  static constexpr int kWarps = 16;
  static constexpr int kStages = 0;
  static constexpr int kSmemPerWarp = 0;
  int bs, bs_next, H;
  const bf* gx;         // [bs, 4H]   input projection + biases of this step's rows (bf16), interleaved columns;
                        //            NULL: the contraction already covers [x | h] (greedy decode) and only `bias` is added
  const float* bias;    // [4H]       interleaved b_ih + b_hh, used when gx == NULL
  const float* c_prev;  // [>=bs, H]  c_{t-1} (NULL at t = 0)
  float* cs;            // [bs, H]    c_t
  bf* hs;               // [bs, ldh]  h_t (layer output rows of this step)
  bf* hprev_next;       // [bs_next, ldn] h_t again: the packed rows of step t+1 / the next layer's input (NULL: skip)
  bf* act;              // [bs, 4H]   sigma(i), sigma(f), tanh(g), sigma(o), interleaved, kept for BPTT (NULL: skip)
  int ldh, ldn;         // row pitches of hs / hprev_next in elements

  // everything the epilogue reads from global memory for its 32 columns (8 hidden units), fetched one tile ahead
  struct Pre {
    uint4 gx[4];   // bf16 input projection (or the bias packed the same way), 32 values
    float4 cp[2];  // c_{t-1}, 8 values
  };
  __device__ __forceinline__ void prefetch(Pre& p, int m_blk, int n_blk, int ew, int lane) const {
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    const int H4 = 4 * H;
    const int col0 = n_blk * 128 + (ew >> 2) * 32;
    if (row >= bs || col0 >= H4) return;
    if (gx) {
      const uint4* g4 = reinterpret_cast<const uint4*>(gx + (int64_t)row * H4 + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) p.gx[q] = __ldg(g4 + q);
    } else {  // bias only, packed to the same bf16 layout as a Gx' row would have
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * q));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * q + 4));
        p.gx[q] = make_uint4(pack_bf2(b0.x, b0.y), pack_bf2(b0.z, b0.w), pack_bf2(b1.x, b1.y), pack_bf2(b1.z, b1.w));
      }
    }
    if (c_prev) {
      const float4* c4 = reinterpret_cast<const float4*>(c_prev + (int64_t)row * H + (col0 >> 2));
      p.cp[0] = c4[0];
      p.cp[1] = c4[1];
    } else {
      p.cp[0] = p.cp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int, int ew, int lane,
                                       const Pre& p, uint8_t*) const {
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    const bool ok = row < bs;
    const int H4 = 4 * H;
    const int c = ew >> 2;
    const int col0 = n_blk * 128 + c * 32;
    if (col0 >= H4) return;  // warp-uniform
    uint32_t r[32];
    tc::tmem_ld32(tmem_rows + (uint32_t)(c * 32), r);
    tc::tmem_ld_wait();
    if (!ok) return;
    const int j0 = col0 >> 2;
    float4 g[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 8 bf16 = 2 hidden units per 16-byte load
      const uint4 v = p.gx[q];
      const float2 a = unpack_bf2(v.x), b = unpack_bf2(v.y), c2 = unpack_bf2(v.z), d = unpack_bf2(v.w);
      g[2 * q] = make_float4(a.x, a.y, b.x, b.y);
      g[2 * q + 1] = make_float4(c2.x, c2.y, d.x, d.y);
    }
    const float cp[8] = {p.cp[0].x, p.cp[0].y, p.cp[0].z, p.cp[0].w, p.cp[1].x, p.cp[1].y, p.cp[1].z, p.cp[1].w};
    float cn[8], hn[8];
    uint32_t ap[16];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float i_ = sigm(__uint_as_float(r[4 * u]) + g[u].x);
      const float f_ = sigm(__uint_as_float(r[4 * u + 1]) + g[u].y);
      const float g_ = tanh_(__uint_as_float(r[4 * u + 2]) + g[u].z);
      const float o_ = sigm(__uint_as_float(r[4 * u + 3]) + g[u].w);
      cn[u] = f_ * cp[u] + i_ * g_;
      hn[u] = o_ * tanh_(cn[u]);
      ap[2 * u] = pack_bf2(i_, f_);
      ap[2 * u + 1] = pack_bf2(g_, o_);
    }
    float4* cd = reinterpret_cast<float4*>(cs + (int64_t)row * H + j0);
    cd[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
    cd[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
    const uint4 hv = make_uint4(pack_bf2(hn[0], hn[1]), pack_bf2(hn[2], hn[3]), pack_bf2(hn[4], hn[5]),
                                pack_bf2(hn[6], hn[7]));
    *reinterpret_cast<uint4*>(hs + (int64_t)row * ldh + j0) = hv;
    if (row < bs_next) *reinterpret_cast<uint4*>(hprev_next + (int64_t)row * ldn + j0) = hv;
    if (act) {  // training only
      uint4* ad = reinterpret_cast<uint4*>(act + (int64_t)row * H4 + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) ad[q] = make_uint4(ap[4 * q], ap[4 * q + 1], ap[4 * q + 2], ap[4 * q + 3]);
    }
  }
};

// =====================================================================================================================
// Persistent forward recurrence: ONE cooperative launch runs all T steps.  CTA (m_blk, n_blk) owns 128 batch rows x
// 32 hidden units (128 interleaved gate columns) for the whole sequence:
//   * its W_hh' slice [128, H] is TMA-loaded into shared memory once and stays resident; c_t stays in registers,
//   * per step only h_{t-1}[128, H] streams in (TMA, 4-stage ring), tcgen05.mma accumulates into one of two TMEM
//     accumulators, 16 epilogue warps (one 32-row x 32-column chunk each) run the gate math,
//   * sequences are independent, so step t+1 of a row block depends only on the num_n CTAs of the SAME row block.
//     They synchronise through global arrival counters, one per (row block, step, 64-unit k-block): a consumer
//     starts fetching k-block kb of h_t as soon as the two CTAs that produce it have published, so the exchange
//     overlaps the stragglers' epilogues.  No grid-wide barrier, no relaunch.
//   * the epilogue publishes h_t (the only thing other CTAs wait for) first; c_t and the saved activations are
//     stored after the release, off the critical path.
// h_t is written with generic stores and read back by other CTAs through TMA (async proxy): see publish_fence().
// With clusters (CL = 4) the CL CTAs that share a row block take turns issuing each k-block as ONE multicast TMA.
// =====================================================================================================================
constexpr int PF_STAGES = 6;
constexpr int PF_EPI_WARPS = 16;
constexpr int PF_THREADS = 128 + 32 * PF_EPI_WARPS;
constexpr int PF_FLAGS_PER_STEP = 8;  // k-blocks of 64 hidden units (H <= 512)

struct PersistFwdParams {
  int H, T, num_n;
  int m_base;  // first 128-row block of this launch (batches wider than one co-resident group take several launches)
  const bf* gx; float* cs; bf* hs; bf* hprev; bf* act;
  int* flags;  // [num_m0][T][PF_FLAGS_PER_STEP] arrival counters, zeroed before the launch
#ifdef SNT_LSTM_DBG
  long long* dbg;  // [T][8] clock64 stamps of one block
#endif
};
#ifdef SNT_LSTM_DBG
#define DBG_STAMP(t, k) do { if (blockIdx.x == SNT_DBG_BLOCK) p.dbg[(t) * 12 + (k)] = clock64(); } while (0)
#else
#define DBG_STAMP(t, k) do {} while (0)
#endif

// Publishing global data that other CTAs will fetch with TMA (async proxy): the writer orders its generic-proxy stores
// before later async-proxy accesses with ONE proxy fence restricted to the global space, then releases at gpu scope.
// The consumer's TMA is issued behind the acquire load of the counter (control dependency), i.e. after the fence in
// causality order.  (The unrestricted fence.proxy.async measured ~1.8k cycles per step on the writer and ~3k on the
// reader side of this kernel; the .global form ~150.)
__device__ __forceinline__ void publish_fence() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int CL>
__global__ void __launch_bounds__(PF_THREADS, 1)
lstm_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ PackInfo pk, const PersistFwdParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KB = (p.H + BK - 1) / BK;                 // k-blocks of the contraction (<= 8)
  uint8_t* sW = smem;                                 // KB x 16 KB, resident
  uint8_t* sA = smem + KB * 16384;                    // PF_STAGES x 16 KB ring
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + PF_STAGES * 16384);
  uint64_t* empty = full + PF_STAGES;
  uint64_t* wbar = empty + PF_STAGES;
  uint64_t* tfull = wbar + 1;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x % p.num_n, m_blk = p.m_base + blockIdx.x / p.num_n;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < PF_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    mbar_init(wbar, 1);
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast traffic
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);

  // number of steps this row block is alive (batch_sizes is non-increasing)
  int t_end = 0;
  while (t_end < p.T && m_blk * BM < pk.off[t_end + 1] - pk.off[t_end]) ++t_end;

  // step 0 needs no contraction: h_{-1} = 0, so the recurrent term vanishes and the epilogue starts from Gx' alone
  if (warp == 0) {
    // ================= TMA producer (lane 0 issues, lanes 0..KB-1 poll one k-block counter each) =================
    int stage = 0;
    uint32_t phase = 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, (uint32_t)(KB * 16384));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * 16384, &tmB, wbar, kb * BK, n_blk * 128);
    }
    const int need = min(2, p.num_n - 2 * lane);  // CTAs that produce the 64 hidden units of k-block `lane`
    for (int t = 1; t < t_end; ++t) {
      const int* f = p.flags + ((int64_t)m_blk * p.T + (t - 1)) * PF_FLAGS_PER_STEP;
      int kb = 0;
      const long long t0 = clock64();
      while (kb < KB) {
        const int v = lane < KB ? ld_acquire_gpu(f + lane) : 0;
        const unsigned waiting = __ballot_sync(0xffffffffu, lane < KB && v < need);
        const int ready = waiting ? __ffs((int)waiting) - 1 : KB;  // k-blocks [0, ready) are published
        if (ready <= kb) {
          if (clock64() - t0 > SNT_MBAR_TIMEOUT_CYCLES) {
            if (lane == 0) printf("snt: lstm persistent flag timeout block %d step %d\n", (int)blockIdx.x, t);
            __trap();
          }
          continue;
        }
        if (lane == 0) {
          if (kb == 0) DBG_STAMP(t, 0);
          for (int k = kb; k < ready; ++k) {
            mbar_wait(&empty[stage], phase ^ 1);  // released by every CTA of the cluster
            mbar_arrive_expect_tx(&full[stage], 16384);
            if (CL == 1)
              tma_load_2d(sA + stage * 16384, &tmA, &full[stage], k * BK, pk.off[t] + m_blk * BM);
            else if ((uint32_t)(k % CL) == crank)  // one L2 read feeds the CL CTAs that share this row block
              tma_load_2d_mc(sA + stage * 16384, &tmA, &full[stage], k * BK, pk.off[t] + m_blk * BM, CMASK);
            if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
          }
          if (ready == KB) DBG_STAMP(t, 1);
        }
        kb = ready;
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(BM, 128, false, false);
      mbar_wait(wbar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 1; t < t_end; ++t) {
        // accumulator (t & 1) is free: its previous reader (step t-2's epilogue) finished before h_{t-1} could exist
        const uint32_t tmem_d = tmem_base + (uint32_t)((t & 1) * 128);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          if (kb == 0) DBG_STAMP(t, 2);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * 16384);
          const uint32_t b_addr = smem_u32(sW + kb * 16384);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (CL == 1) umma_commit(&empty[stage]); else umma_commit_mc(&empty[stage], CMASK);
          if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
        }
        DBG_STAMP(t, 3);
        umma_commit(&tfull[t & 1]);
      }
    }
  } else if (warp >= 4) {
    // ================= gate epilogue: warp -> (TMEM lane quadrant q, 32-column chunk c) =================
    const int q = warp & 3, c = (warp - 4) >> 2;
    const int row = m_blk * BM + q * 32 + lane;
    const int H = p.H, H4 = 4 * p.H;
    const int col0 = n_blk * 128 + c * 32;  // first interleaved gate column of this thread's 8 hidden units
    const int j0 = col0 >> 2;
    const bool col_ok = col0 < H4;
    const bool stamp = threadIdx.x == 128;
    float creg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) creg[u] = 0.f;
    uint4 gxv[4];
    if (col_ok && t_end > 0 && row < pk.off[1]) {
      const uint4* g4 = reinterpret_cast<const uint4*>(p.gx + (int64_t)row * H4 + col0);
#pragma unroll
      for (int k = 0; k < 4; ++k) gxv[k] = __ldg(g4 + k);
    }
    for (int t = 0; t < t_end; ++t) {
      const int bs = pk.off[t + 1] - pk.off[t];
      const int bs_next = t + 1 < p.T ? pk.off[t + 2] - pk.off[t + 1] : 0;
      const bool ok = col_ok && row < bs;
      uint32_t r[32];
      if (t > 0) {
        // tfull[a] completes once per use of accumulator a: steps 1,3,5,.. for a = 1 and 2,4,6,.. for a = 0
        const int a = t & 1;
        mbar_wait(&tfull[a], (uint32_t)(((t >> 1) & 1) ^ (a ^ 1)));
        if (stamp) DBG_STAMP(t, 4);
        tcgen05_fence_after();
        tmem_ld32(tmem_base + (uint32_t)(a * 128 + c * 32) + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
        tcgen05_fence_before();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = 0u;
      }
      float cn[8];
      uint32_t ap[16];
      if (ok) {
        float hn[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 8 bf16 = 2 hidden units per 16-byte load
          const uint4 v = gxv[k];
          const float2 a0 = unpack_bf2(v.x), a1 = unpack_bf2(v.y), b0 = unpack_bf2(v.z), b1 = unpack_bf2(v.w);
          const float gi[2][4] = {{a0.x, a0.y, a1.x, a1.y}, {b0.x, b0.y, b1.x, b1.y}};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int u = 2 * k + e;
            const float i_ = sigm(__uint_as_float(r[4 * u]) + gi[e][0]);
            const float f_ = sigm(__uint_as_float(r[4 * u + 1]) + gi[e][1]);
            const float g_ = tanh_(__uint_as_float(r[4 * u + 2]) + gi[e][2]);
            const float o_ = sigm(__uint_as_float(r[4 * u + 3]) + gi[e][3]);
            cn[u] = f_ * creg[u] + i_ * g_;
            hn[u] = o_ * tanh_(cn[u]);
            creg[u] = cn[u];
            ap[2 * u] = pack_bf2(i_, f_);
            ap[2 * u + 1] = pack_bf2(g_, o_);
          }
        }
        // publish h_t first: it is all the other CTAs of this row block wait for
        const uint4 hv = make_uint4(pack_bf2(hn[0], hn[1]), pack_bf2(hn[2], hn[3]), pack_bf2(hn[4], hn[5]),
                                    pack_bf2(hn[6], hn[7]));
        if (row < bs_next) *reinterpret_cast<uint4*>(p.hprev + ((int64_t)pk.off[t + 1] + row) * H + j0) = hv;
        *reinterpret_cast<uint4*>(p.hs + ((int64_t)pk.off[t] + row) * H + j0) = hv;
      }
      if (stamp) DBG_STAMP(t, 5);
      if (t + 1 < t_end) {
        publish_fence();
        if (stamp) DBG_STAMP(t, 6);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * PF_EPI_WARPS) : "memory");
        if (threadIdx.x == 128)
          atomicAdd(p.flags + ((int64_t)m_blk * p.T + t) * PF_FLAGS_PER_STEP + (n_blk >> 1), 1);
        if (stamp) DBG_STAMP(t, 7);
      }
      // off the critical path: c_t and the activations kept for BPTT, then next step's input projection
      if (ok) {
        float4* cd = reinterpret_cast<float4*>(p.cs + ((int64_t)pk.off[t] + row) * H + j0);
        cd[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
        cd[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
        if (p.act) {  // training only
          uint4* ad = reinterpret_cast<uint4*>(p.act + ((int64_t)pk.off[t] + row) * H4 + col0);
#pragma unroll
          for (int k = 0; k < 4; ++k) ad[k] = make_uint4(ap[4 * k], ap[4 * k + 1], ap[4 * k + 2], ap[4 * k + 3]);
        }
      }
      if (col_ok && t + 1 < t_end && row < bs_next) {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.gx + ((int64_t)pk.off[t + 1] + row) * H4 + col0);
#pragma unroll
        for (int k = 0; k < 4; ++k) gxv[k] = __ldg(g4 + k);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================================
// Persistent BPTT recurrence.  Per step and 128-row block the contraction is dL/dh_t[128, H] = dG'_{t+1}[128, 4H] . W_hh'
// (K = 4H).  A CTA that owned whole output columns would have to stream the full 4H-wide dG'_{t+1} every step (512 KB at
// H = 512: measured 15-20k cycles per step, bound by TMA latency x ring depth), so the work is split over K as well:
//   CTA (m_blk, nq, ks), ns = H/128, ns*ns CTAs per row block:
//     * resident in shared memory: W_hh'^T[units 128nq..+128, gate columns 512ks..+512]  (128 KB, K-major),
//     * per step it streams only dG'_{t+1}[128 rows, 512ks..+512] (128 KB) into tcgen05.mma (M = N = 128) and writes the
//       fp32 partial tile P[128, 128] to an L2-resident scratch (phase 1),
//     * phase 2: it owns the final 128/ns units [128nq + ks*128/ns, ..) of that tile: sums the ns partial slices in a
//       fixed order (deterministic), adds dL/dh_t from above, runs the cell backward (dL/dc carried in registers) and
//       publishes dG'_t (bf16) - the A operand of step t-1 and later of the weight-gradient GEMMs.
//   Two release/acquire counters per (row block, step, CTA): "partial written" and "dG' published" (publish_fence()).
// =====================================================================================================================
constexpr int PB_STAGES = 6;
constexpr int PB_FLAGS_PER_STEP = 16;  // CTAs per row block (ns * ns, ns <= 4)
constexpr int PB_PART_FLOATS = 128 * 128;
constexpr int PS_THREADS = 128 + 32 * (4 + 16);  // phase-split BPTT kernel: control warps, P1 group, P2 group

struct PersistBwdParams {
  int H, T, ns;
  int m_base;  // first 128-row block of this launch
  const float* d_hs; const bf* act; const float* cs; bf* dg;
  float* part;   // [num_m0][ns*ns][2][128*128] partial dL/dh tiles: [finalizer slice][8-unit group][row][8]
  int* dgflag;   // [num_m0][T][PB_FLAGS_PER_STEP]
  int* pflag;    // [num_m0][T][PB_FLAGS_PER_STEP]
#ifdef SNT_LSTM_DBG
  long long* dbg;
#endif
};

// RB = row blocks per CTA.  A step of this recurrence is a chain of L2 round trips (two counter exchanges, two fenced
// publications) during which the CTA's tensor core, TMA ring and epilogue warps mostly wait.  With RB = 2 a CTA owns the
// same weight slice for TWO consecutive 128-row blocks and every role walks the items (t, row block 0), (t, row block 1),
// (t-1, row block 0), ... in that one order: while row block 0's step is on the wire, row block 1's is computed.  The
// recurrence then needs half the CTAs (64 instead of 128 at B = 1024) for about the same wall time, and the step executor
// runs the dW_out contraction on the SMs that are left (csrc/step.cu).  Item i depends only on items < i of the CTAs of
// the same group, and every role of every CTA walks the items in the same order: no circular wait.
template <int NS, int RB>
__global__ void __launch_bounds__(PF_THREADS, 1)
lstm_bwd_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ PackInfo pk, const PersistBwdParams p) {
  using namespace tc;
  constexpr int KB = 8;             // 512 gate columns per K-split
  constexpr int WS = 128 / NS;      // units this CTA finalises
  constexpr int UG = WS / 32;       // groups of 8 units per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                 // KB x 16 KB, resident
  uint8_t* sA = smem + KB * 16384;                    // PB_STAGES x 16 KB ring
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + PB_STAGES * 16384);
  uint64_t* empty = full + PB_STAGES;
  uint64_t* wbar = empty + PB_STAGES;
  uint64_t* tfull = wbar + 1;                         // [RB][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2 * RB);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = blockIdx.x % (NS * NS);
  const int nq = rank / NS, ks = rank % NS;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < PB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(wbar, 1);
    for (int i = 0; i < 2 * RB; ++i) mbar_init(&tfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256 * RB);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Row block r of this CTA is alive for steps [0, t_end[r]); going backward it starts at t_end[r] - 1.  Its step index
  // s = t_end[r] - 1 - t counts from there; s = 0 has no recurrent term (no row of the block is alive at t + 1).
  // batch_sizes is non-increasing, so t_end[0] >= t_end[1]; a row block past the end of the batch has t_end = 0.
  int m_blk[RB], t_end[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    m_blk[r] = p.m_base + (int)(blockIdx.x / (NS * NS)) * RB + r;
    int te = 0;
    while (te < p.T && m_blk[r] * BM < pk.off[te + 1] - pk.off[te]) ++te;
    t_end[r] = te;
  }

  if (warp == 0) {
    // ============ TMA producer (lane 0 issues; lanes 0..NS-1 poll the CTAs that publish this K range) ============
    int stage = 0;
    uint32_t phase = 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, (uint32_t)(KB * 16384));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * 16384, &tmB, wbar, 512 * ks + kb * BK, 128 * nq);
    }
    for (int t = t_end[0] - 2; t >= 0; --t) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t + 1 >= t_end[r]) continue;  // s = 0 (or row block not alive): nothing to contract
        const int s = t_end[r] - 1 - t;
        // gate columns 512ks + 64kb.. are units 128ks + 16kb..: finalised by CTA (nq' = ks, ks' = kb * NS / 8)
        const int* f = p.dgflag + ((int64_t)m_blk[r] * p.T + (t + 1)) * PB_FLAGS_PER_STEP + ks * NS;
        int kb = 0;
        const long long t0 = clock64();
        while (kb < KB) {
          const int v = lane < NS ? ld_acquire_gpu(f + lane) : 0;
          const unsigned waiting = __ballot_sync(0xffffffffu, lane < NS && v < 1);
          const int ready = waiting ? (KB / NS) * (__ffs((int)waiting) - 1) : KB;
          if (ready <= kb) {
            if (clock64() - t0 > SNT_MBAR_TIMEOUT_CYCLES) {
              if (lane == 0) printf("snt: lstm bwd persistent flag timeout block %d step %d\n", (int)blockIdx.x, t);
              __trap();
            }
            continue;
          }
          if (lane == 0) {
            if (kb == 0 && r == 0) DBG_STAMP(s, 0);
            for (int k = kb; k < ready; ++k) {
              mbar_wait(&empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full[stage], 16384);
              tma_load_2d(sA + stage * 16384, &tmA, &full[stage], 512 * ks + k * BK, pk.off[t + 1] + m_blk[r] * BM);
              if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
            }
            if (ready == KB && r == 0) DBG_STAMP(s, 1);
          }
          kb = ready;
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(BM, 128, false, false);
      mbar_wait(wbar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_end[0] - 2; t >= 0; --t) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (t + 1 >= t_end[r]) continue;
          const int s = t_end[r] - 1 - t;
          // accumulator (r, s & 1) is free: its previous reader (step s-2's phase 1) finished before dG'_{t+1} could exist
          const uint32_t tmem_d = tmem_base + (uint32_t)(r * 256 + (s & 1) * 128);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&full[stage], phase);
            if (kb == 0 && r == 0) DBG_STAMP(s, 2);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * 16384);
            const uint32_t b_addr = smem_u32(sW + kb * 16384);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
          }
          if (r == 0) DBG_STAMP(s, 3);
          umma_commit(&tfull[r * 2 + (s & 1)]);
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    // phase 1 (TMEM -> scratch) uses the TMEM mapping: warp -> (lane quadrant q, 32-column chunk c), thread = one row.
    // phase 2 (cell backward) re-maps threads so that global accesses coalesce: warp wi owns rows 8wi..8wi+7, the 4
    // lanes of a row own interleaved unit pairs {8k + 2j, 8k + 2j + 1}, so one store instruction writes 64 contiguous
    // bytes of dG' per row (and one scratch load reads 256 contiguous bytes).
    const int q = warp & 3, c = (warp - 4) >> 2;
    const int row_l1 = q * 32 + lane;
    const int wi = warp - 4, j = lane & 3;
    const int row_l = wi * 8 + (lane >> 2);
    const int H = p.H, H4 = 4 * p.H;
    const int u0 = 128 * nq + WS * ks;  // first unit this CTA finalises
    constexpr int KP = WS / 8;          // 8-unit groups in this CTA's slice
    const bool stamp = threadIdx.x == 128;
    float dcreg[RB][UG][8];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int g = 0; g < UG; ++g)
#pragma unroll
        for (int u = 0; u < 8; ++u) dcreg[r][g][u] = 0.f;
    // inputs of one item (step t of row block mb) for this thread's row and the 4 unit pairs k = 4g..4g+3
    float2 dh2[4], c2[4], p2[4];
    uint4 a4[4];
    auto load_inputs = [&](int t, int mb, int g) {
      const int bs = pk.off[t + 1] - pk.off[t];
      const int row = mb * BM + row_l;
      if (row < bs) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int u = u0 + 8 * (4 * g + kk) + 2 * j;
          const int64_t o1 = ((int64_t)pk.off[t] + row) * H + u;
          dh2[kk] = *reinterpret_cast<const float2*>(p.d_hs + o1);
          c2[kk] = *reinterpret_cast<const float2*>(p.cs + o1);
          p2[kk] = t > 0 ? *reinterpret_cast<const float2*>(p.cs + ((int64_t)pk.off[t - 1] + row) * H + u)
                         : make_float2(0.f, 0.f);
          a4[kk] = *reinterpret_cast<const uint4*>(p.act + ((int64_t)pk.off[t] + row) * H4 + 4 * u);
        }
      }
    };
    if (t_end[0] > 0) load_inputs(t_end[0] - 1, m_blk[0], 0);
    for (int t = t_end[0] - 1; t >= 0; --t) {
      const int bs = pk.off[t + 1] - pk.off[t];
      const int bs_next = t + 1 < p.T ? pk.off[t + 2] - pk.off[t + 1] : 0;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t >= t_end[r]) continue;
        const int s = t_end[r] - 1 - t;
        const int row = m_blk[r] * BM + row_l;
        const bool ok = row < bs;
        const bool has_next = row < bs_next;  // rows still alive at step t+1 carry recurrent gradient
        const int par = s & 1;
        const int* pf = p.pflag + ((int64_t)m_blk[r] * p.T + t) * PB_FLAGS_PER_STEP;
        float* my_part = p.part + ((int64_t)m_blk[r] * (NS * NS) + rank) * 2 * PB_PART_FLOATS;
        if (s > 0) {
          // ---- phase 1: this K-split's partial tile, chunk c (32 units of the 128-unit tile) -> scratch ----
          mbar_wait(&tfull[r * 2 + par], (uint32_t)(((s >> 1) & 1) ^ (par ^ 1)));
          if (stamp && r == 0) DBG_STAMP(s, 4);
          tcgen05_fence_after();
          uint32_t rr[32];
          tmem_ld32(tmem_base + (uint32_t)(r * 256 + par * 128 + c * 32) + ((uint32_t)(q * 32) << 16), rr);
          tmem_ld_wait();
          tcgen05_fence_before();
          {
            const int x0 = 32 * c;  // tile column -> (finalizer slice, 8-unit group inside the slice)
            float* base = my_part + (int64_t)par * PB_PART_FLOATS +
                          (((int64_t)(x0 / WS) * KP + (x0 % WS) / 8) * 128 + row_l1) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              float4* dst = reinterpret_cast<float4*>(base + (int64_t)kk * 128 * 8);
              __stcg(dst, make_float4(__uint_as_float(rr[8 * kk]), __uint_as_float(rr[8 * kk + 1]),
                                      __uint_as_float(rr[8 * kk + 2]), __uint_as_float(rr[8 * kk + 3])));
              __stcg(dst + 1, make_float4(__uint_as_float(rr[8 * kk + 4]), __uint_as_float(rr[8 * kk + 5]),
                                          __uint_as_float(rr[8 * kk + 6]), __uint_as_float(rr[8 * kk + 7])));
            }
          }
          if (stamp && r == 0) DBG_STAMP(s, 5);
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(32 * PF_EPI_WARPS) : "memory");
          if (threadIdx.x == 128) atomicAdd(const_cast<int*>(pf) + rank, 1);
          if (stamp && r == 0) DBG_STAMP(s, 6);
          // ---- wait for the NS partials of this N tile (CTAs nq*NS .. nq*NS+NS-1) ----
          if (warp == 4) {
            const long long t0 = clock64();
            for (;;) {
              const int v = lane < NS ? ld_acquire_gpu(pf + nq * NS + lane) : 1;
              if (__all_sync(0xffffffffu, v >= 1)) break;
              if (clock64() - t0 > SNT_MBAR_TIMEOUT_CYCLES) {
                if (lane == 0) printf("snt: lstm bwd partial timeout block %d step %d\n", (int)blockIdx.x, t);
                __trap();
              }
            }
          }
          asm volatile("bar.sync 2, %0;" ::"n"(32 * PF_EPI_WARPS) : "memory");
          if (stamp && r == 0) DBG_STAMP(s, 7);
        }
        // ---- phase 2: final dL/dh for this CTA's units, cell backward, publish dG'_t ----
#pragma unroll
        for (int g = 0; g < UG; ++g) {
          if (g > 0) load_inputs(t, m_blk[r], g);
          if (ok) {
            float2 rec[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) rec[kk] = make_float2(0.f, 0.f);
            if (s > 0 && has_next) {
              float2 x[NS][4];
#pragma unroll
              for (int k = 0; k < NS; ++k) {
                const float* src = p.part + (((int64_t)m_blk[r] * (NS * NS) + nq * NS + k) * 2 + par) * PB_PART_FLOATS +
                                   (((int64_t)ks * KP + 4 * g) * 128 + row_l) * 8 + 2 * j;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) x[k][kk] = __ldcg(reinterpret_cast<const float2*>(src + (int64_t)kk * 128 * 8));
              }
#pragma unroll
              for (int k = 0; k < NS; ++k)  // fixed order: deterministic
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { rec[kk].x += x[k][kk].x; rec[kk].y += x[k][kk].y; }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a[4] = {a4[kk].x, a4[kk].y, a4[kk].z, a4[kk].w};
              const float dhv[2] = {dh2[kk].x + rec[kk].x, dh2[kk].y + rec[kk].y};
              const float cv[2] = {c2[kk].x, c2[kk].y}, pv[2] = {p2[kk].x, p2[kk].y};
              uint32_t go[4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float2 if_ = unpack_bf2(a[2 * e]), go_ = unpack_bf2(a[2 * e + 1]);
                const float i_ = if_.x, f_ = if_.y, g_ = go_.x, o_ = go_.y;
                const float tc_ = tanh_(cv[e]);
                const float dh = dhv[e];
                const float dc = dcreg[r][g][2 * kk + e] + dh * o_ * (1.f - tc_ * tc_);
                go[2 * e] = pack_bf2(dc * g_ * i_ * (1.f - i_), dc * pv[e] * f_ * (1.f - f_));
                go[2 * e + 1] = pack_bf2(dc * i_ * (1.f - g_ * g_), dh * tc_ * o_ * (1.f - o_));
                dcreg[r][g][2 * kk + e] = dc * f_;
              }
              const int u = u0 + 8 * (4 * g + kk) + 2 * j;
              *reinterpret_cast<uint4*>(p.dg + ((int64_t)pk.off[t] + row) * H4 + 4 * u) = make_uint4(go[0], go[1], go[2], go[3]);
            }
          }
        }
        if (stamp && r == 0) DBG_STAMP(s, 8);
        if (t > 0) {
          publish_fence();
          asm volatile("bar.sync 1, %0;" ::"n"(32 * PF_EPI_WARPS) : "memory");
          if (threadIdx.x == 128)
            atomicAdd(p.dgflag + ((int64_t)m_blk[r] * p.T + t) * PB_FLAGS_PER_STEP + rank, 1);
          if (stamp && r == 0) DBG_STAMP(s, 9);
        }
        // inputs of the next item in the walk: the next alive row block of this step, else row block 0 of step t - 1
        {
          bool found = false;
#pragma unroll
          for (int r2 = r + 1; r2 < RB; ++r2)
            if (!found && t < t_end[r2]) { load_inputs(t, m_blk[r2], 0); found = true; }
          if (!found && t > 0) load_inputs(t - 1, m_blk[0], 0);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256 * RB);
  }
}

// PHASE-SPLIT variant (the default for RB = 2).  In the kernel above one group of 16 epilogue warps walks phase 1 and phase 2
// of every item in turn, and with two row blocks per CTA those ~10k busy cycles per item - mostly fences and L2 round
// trips - add up to more than the step's dependency chain (measured: BPTT 188 -> 300 us).  Here the two phases belong to
// two warp groups that meet only through the global partial counters the finalisers poll anyway: P1 (4 warps, one per TMEM
// lane quadrant) drains accumulators to the scratch and publishes "partial written"; P2 (16 warps, the coalesced row
// mapping) sums the partials, runs the cell backward and publishes dG'.  P1 of the next item overlaps P2 of this one.
// XRB = row blocks per CTA.  A step of this recurrence is a chain of L2 round trips (two counter exchanges, two fenced
// publications) during which the CTA's tensor core, TMA ring and epilogue warps mostly wait.  With RB = 2 a CTA owns the
// same weight slice for TWO consecutive 128-row blocks and every role walks the items (t, row block 0), (t, row block 1),
// (t-1, row block 0), ... in that one order: while row block 0's step is on the wire, row block 1's is computed.  The
// recurrence then needs half the CTAs (64 instead of 128 at B = 1024) for about the same wall time, and the step executor
// runs the dW_out contraction on the SMs that are left (csrc/step.cu).  Item i depends only on items < i of the CTAs of
// the same group, and every role of every CTA walks the items in the same order: no circular wait.
template <int NS, int RB>
__global__ void __launch_bounds__(PS_THREADS, 1)
lstm_bwd_persistent_ps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ PackInfo pk, const PersistBwdParams p) {
  using namespace tc;
  constexpr int KB = 8;             // 512 gate columns per K-split
  constexpr int WS = 128 / NS;      // units this CTA finalises
  constexpr int UG = WS / 32;       // groups of 8 units per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                 // KB x 16 KB, resident
  uint8_t* sA = smem + KB * 16384;                    // PB_STAGES x 16 KB ring
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + PB_STAGES * 16384);
  uint64_t* empty = full + PB_STAGES;
  uint64_t* wbar = empty + PB_STAGES;
  uint64_t* tfull = wbar + 1;                         // [RB][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2 * RB);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = blockIdx.x % (NS * NS);
  const int nq = rank / NS, ks = rank % NS;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < PB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(wbar, 1);
    for (int i = 0; i < 2 * RB; ++i) mbar_init(&tfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256 * RB);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Row block r of this CTA is alive for steps [0, t_end[r]); going backward it starts at t_end[r] - 1.  Its step index
  // s = t_end[r] - 1 - t counts from there; s = 0 has no recurrent term (no row of the block is alive at t + 1).
  // batch_sizes is non-increasing, so t_end[0] >= t_end[1]; a row block past the end of the batch has t_end = 0.
  int m_blk[RB], t_end[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    m_blk[r] = p.m_base + (int)(blockIdx.x / (NS * NS)) * RB + r;
    int te = 0;
    while (te < p.T && m_blk[r] * BM < pk.off[te + 1] - pk.off[te]) ++te;
    t_end[r] = te;
  }

  if (warp == 0) {
    // ============ TMA producer (lane 0 issues; lanes 0..NS-1 poll the CTAs that publish this K range) ============
    int stage = 0;
    uint32_t phase = 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, (uint32_t)(KB * 16384));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * 16384, &tmB, wbar, 512 * ks + kb * BK, 128 * nq);
    }
    for (int t = t_end[0] - 2; t >= 0; --t) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t + 1 >= t_end[r]) continue;  // s = 0 (or row block not alive): nothing to contract
        const int s = t_end[r] - 1 - t;
        // gate columns 512ks + 64kb.. are units 128ks + 16kb..: finalised by CTA (nq' = ks, ks' = kb * NS / 8)
        const int* f = p.dgflag + ((int64_t)m_blk[r] * p.T + (t + 1)) * PB_FLAGS_PER_STEP + ks * NS;
        int kb = 0;
        const long long t0 = clock64();
        while (kb < KB) {
          const int v = lane < NS ? ld_acquire_gpu(f + lane) : 0;
          const unsigned waiting = __ballot_sync(0xffffffffu, lane < NS && v < 1);
          const int ready = waiting ? (KB / NS) * (__ffs((int)waiting) - 1) : KB;
          if (ready <= kb) {
            if (clock64() - t0 > SNT_MBAR_TIMEOUT_CYCLES) {
              if (lane == 0) printf("snt: lstm bwd persistent flag timeout block %d step %d\n", (int)blockIdx.x, t);
              __trap();
            }
            continue;
          }
          if (lane == 0) {
            if (kb == 0 && r == 0) DBG_STAMP(s, 0);
            for (int k = kb; k < ready; ++k) {
              mbar_wait(&empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full[stage], 16384);
              tma_load_2d(sA + stage * 16384, &tmA, &full[stage], 512 * ks + k * BK, pk.off[t + 1] + m_blk[r] * BM);
              if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
            }
            if (ready == KB && r == 0) DBG_STAMP(s, 1);
          }
          kb = ready;
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(BM, 128, false, false);
      mbar_wait(wbar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_end[0] - 2; t >= 0; --t) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (t + 1 >= t_end[r]) continue;
          const int s = t_end[r] - 1 - t;
          // accumulator (r, s & 1) is free: its previous reader (step s-2's phase 1) finished before dG'_{t+1} could exist
          const uint32_t tmem_d = tmem_base + (uint32_t)(r * 256 + (s & 1) * 128);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&full[stage], phase);
            if (kb == 0 && r == 0) DBG_STAMP(s, 2);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * 16384);
            const uint32_t b_addr = smem_u32(sW + kb * 16384);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
          }
          if (r == 0) DBG_STAMP(s, 3);
          umma_commit(&tfull[r * 2 + (s & 1)]);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= P1: TMEM -> scratch, one warp per lane quadrant, thread = one row of the tile =================
    const int q = warp & 3;
    const int row_l1 = q * 32 + lane;
    constexpr int KP = WS / 8;          // 8-unit groups in a finaliser's slice
    const bool stamp = threadIdx.x == 128;
    for (int t = t_end[0] - 2; t >= 0; --t) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t + 1 >= t_end[r]) continue;
        const int s = t_end[r] - 1 - t;
        const int par = s & 1;
        int* pf = p.pflag + ((int64_t)m_blk[r] * p.T + t) * PB_FLAGS_PER_STEP;
        float* my_part = p.part + ((int64_t)m_blk[r] * (NS * NS) + rank) * 2 * PB_PART_FLOATS;
        mbar_wait(&tfull[r * 2 + par], (uint32_t)(((s >> 1) & 1) ^ (par ^ 1)));
        if (stamp && r == 0) DBG_STAMP(s, 4);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {  // 32-column chunks of the 128-unit tile
          uint32_t rr[32];
          tmem_ld32(tmem_base + (uint32_t)(r * 256 + par * 128 + c * 32) + ((uint32_t)(q * 32) << 16), rr);
          tmem_ld_wait();
          const int x0 = 32 * c;  // tile column -> (finalizer slice, 8-unit group inside the slice)
          float* base = my_part + (int64_t)par * PB_PART_FLOATS +
                        (((int64_t)(x0 / WS) * KP + (x0 % WS) / 8) * 128 + row_l1) * 8;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float4* dst = reinterpret_cast<float4*>(base + (int64_t)kk * 128 * 8);
            __stcg(dst, make_float4(__uint_as_float(rr[8 * kk]), __uint_as_float(rr[8 * kk + 1]),
                                    __uint_as_float(rr[8 * kk + 2]), __uint_as_float(rr[8 * kk + 3])));
            __stcg(dst + 1, make_float4(__uint_as_float(rr[8 * kk + 4]), __uint_as_float(rr[8 * kk + 5]),
                                        __uint_as_float(rr[8 * kk + 6]), __uint_as_float(rr[8 * kk + 7])));
          }
        }
        tcgen05_fence_before();
        if (stamp && r == 0) DBG_STAMP(s, 5);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 128) atomicAdd(pf + rank, 1);
        if (stamp && r == 0) DBG_STAMP(s, 6);
      }
    }
  } else if (warp >= 8) {
    // ================= P2: cell backward.  Threads are mapped so that global accesses coalesce: warp wi owns rows
    // 8wi..8wi+7, the 4 lanes of a row own interleaved unit pairs {8k + 2j, 8k + 2j + 1}, so one store instruction writes
    // 64 contiguous bytes of dG' per row (and one scratch load reads 256 contiguous bytes). =================
    const int wi = warp - 8, j = lane & 3;
    const int row_l = wi * 8 + (lane >> 2);
    const int H = p.H, H4 = 4 * p.H;
    const int u0 = 128 * nq + WS * ks;  // first unit this CTA finalises
    constexpr int KP = WS / 8;          // 8-unit groups in this CTA's slice
    const bool stamp = threadIdx.x == 256;
    float dcreg[RB][UG][8];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int g = 0; g < UG; ++g)
#pragma unroll
        for (int u = 0; u < 8; ++u) dcreg[r][g][u] = 0.f;
    // inputs of one item (step t of row block mb) for this thread's row and the 4 unit pairs k = 4g..4g+3
    float2 dh2[4], c2[4], p2[4];
    uint4 a4[4];
    auto load_inputs = [&](int t, int mb, int g) {
      const int bs = pk.off[t + 1] - pk.off[t];
      const int row = mb * BM + row_l;
      if (row < bs) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int u = u0 + 8 * (4 * g + kk) + 2 * j;
          const int64_t o1 = ((int64_t)pk.off[t] + row) * H + u;
          dh2[kk] = *reinterpret_cast<const float2*>(p.d_hs + o1);
          c2[kk] = *reinterpret_cast<const float2*>(p.cs + o1);
          p2[kk] = t > 0 ? *reinterpret_cast<const float2*>(p.cs + ((int64_t)pk.off[t - 1] + row) * H + u)
                         : make_float2(0.f, 0.f);
          a4[kk] = *reinterpret_cast<const uint4*>(p.act + ((int64_t)pk.off[t] + row) * H4 + 4 * u);
        }
      }
    };
    if (t_end[0] > 0) load_inputs(t_end[0] - 1, m_blk[0], 0);
    for (int t = t_end[0] - 1; t >= 0; --t) {
      const int bs = pk.off[t + 1] - pk.off[t];
      const int bs_next = t + 1 < p.T ? pk.off[t + 2] - pk.off[t + 1] : 0;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t >= t_end[r]) continue;
        const int s = t_end[r] - 1 - t;
        const int row = m_blk[r] * BM + row_l;
        const bool ok = row < bs;
        const bool has_next = row < bs_next;  // rows still alive at step t+1 carry recurrent gradient
        const int par = s & 1;
        const int* pf = p.pflag + ((int64_t)m_blk[r] * p.T + t) * PB_FLAGS_PER_STEP;
        if (s > 0) {
          // ---- wait for the NS partials of this N tile (P1 of CTAs nq*NS .. nq*NS+NS-1, this CTA's among them) ----
          if (warp == 8) {
            const long long t0 = clock64();
            for (;;) {
              const int v = lane < NS ? ld_acquire_gpu(pf + nq * NS + lane) : 1;
              if (__all_sync(0xffffffffu, v >= 1)) break;
              if (clock64() - t0 > SNT_MBAR_TIMEOUT_CYCLES) {
                if (lane == 0) printf("snt: lstm bwd partial timeout block %d step %d\n", (int)blockIdx.x, t);
                __trap();
              }
            }
          }
          asm volatile("bar.sync 2, 512;" ::: "memory");
          if (stamp && r == 0) DBG_STAMP(s, 7);
        }
        // ---- final dL/dh for this CTA's units, cell backward, publish dG'_t ----
#pragma unroll
        for (int g = 0; g < UG; ++g) {
          if (g > 0) load_inputs(t, m_blk[r], g);
          if (ok) {
            float2 rec[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) rec[kk] = make_float2(0.f, 0.f);
            if (s > 0 && has_next) {
              float2 x[NS][4];
#pragma unroll
              for (int k = 0; k < NS; ++k) {
                const float* src = p.part + (((int64_t)m_blk[r] * (NS * NS) + nq * NS + k) * 2 + par) * PB_PART_FLOATS +
                                   (((int64_t)ks * KP + 4 * g) * 128 + row_l) * 8 + 2 * j;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) x[k][kk] = __ldcg(reinterpret_cast<const float2*>(src + (int64_t)kk * 128 * 8));
              }
#pragma unroll
              for (int k = 0; k < NS; ++k)  // fixed order: deterministic
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { rec[kk].x += x[k][kk].x; rec[kk].y += x[k][kk].y; }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a[4] = {a4[kk].x, a4[kk].y, a4[kk].z, a4[kk].w};
              const float dhv[2] = {dh2[kk].x + rec[kk].x, dh2[kk].y + rec[kk].y};
              const float cv[2] = {c2[kk].x, c2[kk].y}, pv[2] = {p2[kk].x, p2[kk].y};
              uint32_t go[4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float2 if_ = unpack_bf2(a[2 * e]), go_ = unpack_bf2(a[2 * e + 1]);
                const float i_ = if_.x, f_ = if_.y, g_ = go_.x, o_ = go_.y;
                const float tc_ = tanh_(cv[e]);
                const float dh = dhv[e];
                const float dc = dcreg[r][g][2 * kk + e] + dh * o_ * (1.f - tc_ * tc_);
                go[2 * e] = pack_bf2(dc * g_ * i_ * (1.f - i_), dc * pv[e] * f_ * (1.f - f_));
                go[2 * e + 1] = pack_bf2(dc * i_ * (1.f - g_ * g_), dh * tc_ * o_ * (1.f - o_));
                dcreg[r][g][2 * kk + e] = dc * f_;
              }
              const int u = u0 + 8 * (4 * g + kk) + 2 * j;
              *reinterpret_cast<uint4*>(p.dg + ((int64_t)pk.off[t] + row) * H4 + 4 * u) = make_uint4(go[0], go[1], go[2], go[3]);
            }
          }
        }
        if (stamp && r == 0) DBG_STAMP(s, 8);
        if (t > 0) {
          publish_fence();
          asm volatile("bar.sync 3, 512;" ::: "memory");
          if (threadIdx.x == 256)
            atomicAdd(p.dgflag + ((int64_t)m_blk[r] * p.T + t) * PB_FLAGS_PER_STEP + rank, 1);
          if (stamp && r == 0) DBG_STAMP(s, 9);
        }
        // inputs of the next item in the walk: the next alive row block of this step, else row block 0 of step t - 1
        {
          bool found = false;
#pragma unroll
          for (int r2 = r + 1; r2 < RB; ++r2)
            if (!found && t < t_end[r2]) { load_inputs(t, m_blk[r2], 0); found = true; }
          if (!found && t > 0) load_inputs(t - 1, m_blk[0], 0);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256 * RB);
  }
}

// ---- backward step, split-K variant: the recurrent contraction dG'_{t+1} . W_hh' runs as a split-K GEMM over all SMs
// (partials in an L2-resident scratch), and this coalesced pointwise kernel sums the partials in a fixed order and
// runs the cell backward.  One thread = one packed row x 8 hidden units.
__global__ void __launch_bounds__(256)
lstm_bwd_point_kernel(int bs, int bs_next, int H, const float* __restrict__ d_hs, const bf* __restrict__ act,
                      const float* __restrict__ cs, const float* __restrict__ c_prev,
                      const float* __restrict__ partial, int splits, int64_t split_stride,
                      float* __restrict__ dc_state, bf* __restrict__ dg) {
  const int groups = H >> 3;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)bs * groups) return;
  const int row = (int)(idx / groups), ju = (int)(idx % groups) * 8;
  const bool has_next = row < bs_next;
  const int64_t o1 = (int64_t)row * H + ju;
  float dhv[8], cv[8], pv[8], sv[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(d_hs + o1), b = *reinterpret_cast<const float4*>(d_hs + o1 + 4);
    dhv[0] = a.x; dhv[1] = a.y; dhv[2] = a.z; dhv[3] = a.w; dhv[4] = b.x; dhv[5] = b.y; dhv[6] = b.z; dhv[7] = b.w;
    const float4 c0 = *reinterpret_cast<const float4*>(cs + o1), c1 = *reinterpret_cast<const float4*>(cs + o1 + 4);
    cv[0] = c0.x; cv[1] = c0.y; cv[2] = c0.z; cv[3] = c0.w; cv[4] = c1.x; cv[5] = c1.y; cv[6] = c1.z; cv[7] = c1.w;
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) { pv[u] = 0.f; sv[u] = 0.f; }
  if (c_prev) {
    const float4 a = *reinterpret_cast<const float4*>(c_prev + o1), b = *reinterpret_cast<const float4*>(c_prev + o1 + 4);
    pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = b.x; pv[5] = b.y; pv[6] = b.z; pv[7] = b.w;
  }
  if (has_next) {
    const float4 a = *reinterpret_cast<const float4*>(dc_state + o1), b = *reinterpret_cast<const float4*>(dc_state + o1 + 4);
    sv[0] = a.x; sv[1] = a.y; sv[2] = a.z; sv[3] = a.w; sv[4] = b.x; sv[5] = b.y; sv[6] = b.z; sv[7] = b.w;
    for (int k = 0; k < splits; ++k) {  // fixed order: deterministic
      const float* p = partial + (int64_t)k * split_stride + o1;
      const float4 x = __ldcg(reinterpret_cast<const float4*>(p)), y = __ldcg(reinterpret_cast<const float4*>(p + 4));
      dhv[0] += x.x; dhv[1] += x.y; dhv[2] += x.z; dhv[3] += x.w; dhv[4] += y.x; dhv[5] += y.y; dhv[6] += y.z; dhv[7] += y.w;
    }
  }
  const uint4* a4 = reinterpret_cast<const uint4*>(act + (int64_t)row * 4 * H + 4 * ju);
  uint32_t a[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) { const uint4 v = a4[q]; a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w; }
  float dcn[8];
  uint32_t go[16];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float2 if_ = unpack_bf2(a[2 * u]), go_ = unpack_bf2(a[2 * u + 1]);
    const float i_ = if_.x, f_ = if_.y, g_ = go_.x, o_ = go_.y;
    const float tc_ = tanh_(cv[u]);
    const float dh = dhv[u];
    const float dc = sv[u] + dh * o_ * (1.f - tc_ * tc_);
    go[2 * u] = pack_bf2(dc * g_ * i_ * (1.f - i_), dc * pv[u] * f_ * (1.f - f_));
    go[2 * u + 1] = pack_bf2(dc * i_ * (1.f - g_ * g_), dh * tc_ * o_ * (1.f - o_));
    dcn[u] = dc * f_;
  }
  float4* sd = reinterpret_cast<float4*>(dc_state + o1);
  sd[0] = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
  sd[1] = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
  uint4* gd = reinterpret_cast<uint4*>(dg + (int64_t)row * 4 * H + 4 * ju);
#pragma unroll
  for (int q = 0; q < 4; ++q) gd[q] = make_uint4(go[4 * q], go[4 * q + 1], go[4 * q + 2], go[4 * q + 3]);
}

// ---------------------------------------------------------------------------------------------------------------------
struct LstmWs {
  float* bsum; bf* w_ih; bf* w_hh; bf* w_hh_t; bf* gx; float* dc_state; float* cpart; float* tmp; float* sws; int* flags; float* part; bool ok;
};
static int64_t csb_partials(int64_t R, int64_t C) { return ((R + 255) / 256) * C; }
static LstmWs carve(void* ws, int64_t ws_bytes, int64_t N, int64_t B, int64_t In, int64_t H) {
  Workspace w(ws, ws_bytes);
  LstmWs r;
  r.bsum = w.take<float>(4 * H);
  r.w_ih = w.take<bf>(4 * H * In);
  r.w_hh = w.take<bf>(4 * H * H);
  r.gx = w.take<bf>(N * 4 * H);
  r.dc_state = w.take<float>(B * H);
  r.cpart = w.take<float>(csb_partials(N, 4 * H));
  r.tmp = w.take<float>(4 * H);
  r.sws = w.take<float>(MAX_SPLITS * 4 * H * (In > H ? In : H));
  r.flags = w.take<int>(((B + 127) / 128) * SNT_MAX_T * PB_FLAGS_PER_STEP * 2);
  r.w_hh_t = w.take<bf>(4 * H * H);
  r.part = w.take<float>(((B + 127) / 128) * PB_FLAGS_PER_STEP * 2 * PB_PART_FLOATS);
  r.ok = w.ok();
  return r;
}
int64_t lstm_ws_bytes(int64_t N, int64_t B, int64_t In, int64_t H) {
  return 2 * ws_bytes_for(4 * H, 4) + ws_bytes_for(4 * H * In, 2) + ws_bytes_for(4 * H * H, 2) +
         ws_bytes_for(N * 4 * H, 2) + ws_bytes_for(B * H, 4) + ws_bytes_for(csb_partials(N, 4 * H), 4) +
         ws_bytes_for(MAX_SPLITS * 4 * H * (In > H ? In : H), 4) + ws_bytes_for(((B + 127) / 128) * SNT_MAX_T * PB_FLAGS_PER_STEP * 2, 4) + ws_bytes_for(4 * H * H, 2) +
         ws_bytes_for(((B + 127) / 128) * PB_FLAGS_PER_STEP * 2 * PB_PART_FLOATS, 4);
}

// the persistent kernels need every CTA co-resident (cooperative launch) and ~200 KB of shared memory
static void persist_caps(int* coop, int* max_smem) {
  static int c = -1, m = 0;
  if (c < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&c, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&m, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  *coop = c;
  *max_smem = m;
}

// Launch a persistent recurrence kernel.  All CTAs must be co-resident (the row-block counters are spin-waited), so
// the launch is cooperative; with cluster > 1 the num_n CTAs of a row block are grouped into clusters that share
// their TMA loads by multicast.  Returns the cluster size used through *cluster_used (0: could not launch).
template <class Params>
static int launch_persistent(void (*k1)(CUtensorMap, CUtensorMap, PackInfo, Params),
                             void (*k4)(CUtensorMap, CUtensorMap, PackInfo, Params), int ctas, int num_n, size_t smem,
                             const CUtensorMap& ta, const CUtensorMap& tb, const PackInfo& pk, const Params& pp,
                             cudaStream_t st, int threads = PF_THREADS) {
  static int mode = -1;  // per instantiation: 4 = clusters of 4 fit, 1 = no clusters
  // Measured on B200 (profiles/r01_lstm_persistent_timeline.txt): clusters of 4 with multicast loads do not shorten the
  // per-step exchange (L2 already serves the 4 unicast requests of a line from one fill), and cooperative + cluster
  // launches fail under ncu, so the default is plain cooperative; SNT_PERSIST_CLUSTER=4 opts in.
  const char* env = getenv("SNT_PERSIST_CLUSTER");
  int want = env ? atoi(env) : 1;
  if (num_n % 4 != 0 || want != 4 || k4 == nullptr) want = 1;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int cl = (want == 4 && mode != 1) ? 4 : 1;
    auto kern = cl == 4 ? k4 : k1;
    SNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cl;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl > 1 ? 2 : 1;
    if (cl > 1 && mode < 0) {  // first use: do that many clusters fit at once?
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters * cl < ctas) {
        cudaGetLastError();
        mode = 1;
        continue;
      }
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, pk, pp);
    if (e == cudaSuccess) {
#ifdef SNT_LSTM_DBG
      if (mode < 0) fprintf(stderr, "[lstm dbg] persistent launch: cluster %d\n", cl);
#endif
      if (mode < 0) mode = cl;
      count_launch();
      return SNT_OK;
    }
    if (cl > 1) {  // cluster + cooperative refused on this driver: fall back to plain cooperative
      cudaGetLastError();
      mode = 1;
      continue;
    }
    return check_cuda(e, "persistent LSTM launch");
  }
  set_error("persistent LSTM launch failed");
  return SNT_EINVAL;
}

// One launch prepares everything a recurrence call needs: bf16 unit-interleaved W_ih' / W_hh' rows, the summed and
// interleaved bias, the transposed W_hh'^T of the persistent BPTT kernel, and the cleared regions (h_{-1}, arrival
// counters).  Sections are laid out back to back over blockIdx.x; a pointer that is NULL drops its section.
struct PrepJob {
  const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
  int H, In;
  bf* o_ih; bf* o_hh; float* o_bias; bf* o_hh_t;
  int ld_ih, ld_hh;                     // row pitches of o_ih / o_hh (In / H, or In + H for the concatenated [W_ih'|W_hh'])
  uint4* zero[2]; int64_t zero_n16[2];  // regions to clear, in 16-byte units
  int nA, nB, nC, nD;                   // blocks per section
};
__global__ void __launch_bounds__(256)
lstm_prep_kernel(const PrepJob j) {
  __shared__ float tile[32][33];
  int b = blockIdx.x;
  const int H = j.H;
  if (b < j.nA) {  // row rp = 4*unit + gate of the interleaved weights
    const int rp = b, src = (rp & 3) * H + (rp >> 2);
    if (j.o_ih)
      for (int c = threadIdx.x; c < j.In; c += 256)
        j.o_ih[(int64_t)rp * j.ld_ih + c] = __float2bfloat16_rn(j.w_ih[(int64_t)src * j.In + c]);
    if (j.o_hh)
      for (int c = threadIdx.x; c < H; c += 256)
        j.o_hh[(int64_t)rp * j.ld_hh + c] = __float2bfloat16_rn(j.w_hh[(int64_t)src * H + c]);
    return;
  }
  b -= j.nA;
  if (b < j.nB) {  // 32 x 32 tile of W_hh -> W^T with interleaved contraction index: out[n][4j + g]
    const int tiles_n = (H + 31) / 32;
    const int n0 = (b % tiles_n) * 32, r0 = (b / tiles_n) * 32;  // r = gate-major source row
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, n = n0 + tx;
      tile[i][tx] = (r < 4 * H && n < H) ? j.w_hh[(int64_t)r * H + n] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int n = n0 + i, r = r0 + tx;
      if (n < H && r < 4 * H) {
        const int g = r / H, u = r - g * H;
        j.o_hh_t[(int64_t)n * 4 * H + 4 * u + g] = __float2bfloat16_rn(tile[tx][i]);
      }
    }
    return;
  }
  b -= j.nB;
  if (b < j.nC) {
    const int rp = b * 256 + threadIdx.x;
    if (rp < 4 * H) {
      const int src = (rp & 3) * H + (rp >> 2);
      j.o_bias[rp] = j.b_ih[src] + j.b_hh[src];
    }
    return;
  }
  b -= j.nC;
  for (int k = 0; k < 2; ++k)
    for (int64_t i = (int64_t)b * 256 + threadIdx.x; i < j.zero_n16[k]; i += (int64_t)j.nD * 256)
      j.zero[k][i] = make_uint4(0u, 0u, 0u, 0u);
}
// zero0/zero1: optional regions to clear (byte counts are rounded up to 16: callers pass padded workspace regions)
static int prep_launch(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int64_t In, int64_t H,
                       bf* o_ih, bf* o_hh, float* o_bias, bf* o_hh_t, void* zero0, int64_t zero0_bytes, void* zero1,
                       int64_t zero1_bytes, cudaStream_t st, int64_t ld_ih = 0, int64_t ld_hh = 0) {
  PrepJob j;
  j.ld_ih = (int)(ld_ih ? ld_ih : In); j.ld_hh = (int)(ld_hh ? ld_hh : H);
  j.w_ih = w_ih; j.w_hh = w_hh; j.b_ih = b_ih; j.b_hh = b_hh; j.H = (int)H; j.In = (int)In;
  j.o_ih = o_ih; j.o_hh = o_hh; j.o_bias = (b_ih && b_hh) ? o_bias : nullptr; j.o_hh_t = o_hh_t;
  j.zero[0] = (uint4*)zero0; j.zero_n16[0] = zero0 ? (zero0_bytes + 15) / 16 : 0;
  j.zero[1] = (uint4*)zero1; j.zero_n16[1] = zero1 ? (zero1_bytes + 15) / 16 : 0;
  j.nA = (o_ih || o_hh) ? (int)(4 * H) : 0;
  j.nB = o_hh_t ? (int)(((H + 31) / 32) * ((4 * H + 31) / 32)) : 0;
  j.nC = j.o_bias ? (int)((4 * H + 255) / 256) : 0;
  const int64_t z = j.zero_n16[0] + j.zero_n16[1];
  j.nD = z > 0 ? (int)(z / 1024 + 1 > 128 ? 128 : z / 1024 + 1) : 0;
  const int total = j.nA + j.nB + j.nC + j.nD;
  if (total == 0) return SNT_OK;
  lstm_prep_kernel<<<(unsigned)total, 256, 0, st>>>(j);
  SNT_LAUNCH_CHECK("lstm_prep_kernel");
  return SNT_OK;
}
#define SNT_REQ8(v, what)                                                                               \
  do {                                                                                                  \
    if ((v) % 8 != 0) {                                                                                 \
      set_error("bf16 mode: %s=%lld must be a multiple of 8 (TMA row pitch is 16 bytes)", what,         \
                (long long)(v));                                                                        \
      return SNT_EUNSUPPORTED;                                                                          \
    }                                                                                                   \
  } while (0)

// `gates` (caller-owned, N*4H*4 bytes) holds, in bf16 mode: [0, N*4H) bf16 saved activations (interleaved columns),
// [N*4H, 2*N*4H) bf16 pre-activation gradients written by lstm_bwd.
int64_t lstm_prepared_flag_ints(int64_t B) { return ((B + 127) / 128) * SNT_MAX_T * PB_FLAGS_PER_STEP * 2; }
int lstm_prepare(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int64_t In, int64_t H,
                 int64_t B, const LstmPrepared& o, void* hprev, cudaStream_t st) {
  SNT_REQ8(In, "In");
  SNT_REQ8(H, "H");
  SNT_REQUIRE(o.w_ih && o.w_hh && o.w_hh_t && o.bsum && o.flags_fwd && o.flags_bwd && hprev, "lstm_prepare: NULL buffer");
  SNT_REQUIRE(o.flags_bwd == o.flags_fwd + lstm_prepared_flag_ints(B), "lstm_prepare: counter regions must be adjacent");
  return prep_launch(w_ih, w_hh, b_ih, b_hh, In, H, o.w_ih, o.w_hh, o.bsum, o.w_hh_t, hprev, (int64_t)sizeof(bf) * B * H,
                     o.flags_fwd, (int64_t)sizeof(int) * 2 * lstm_prepared_flag_ints(B), st);
}

int lstm_fwd(const PackInfo& pk, const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
             const float* b_ih, const float* b_hh, float* gates, float* cs, void* hs, void* hprev, void* ws,
             int64_t ws_bytes, cudaStream_t st, const LstmPrepared* prep) {
  SNT_REQ8(In, "In");
  SNT_REQ8(H, "H");
  const int T = pk.T;
  const int64_t N = pk.off[T], B = pk.off[1];
  LstmWs w = carve(ws, ws_bytes, N, B, In, H);
  if (!w.ok) { set_error("bf16 lstm_fwd: workspace too small"); return SNT_EWORKSPACE; }
  bf* hs_b = (bf*)hs;
  bf* hp_b = (bf*)hprev;
  bf* act = (bf*)gates;
  if (prep) {  // prepared by the caller (lstm_prepare): weights, bias, h_{-1} = 0, cleared counters
    w.w_ih = prep->w_ih; w.w_hh = prep->w_hh; w.bsum = prep->bsum; w.flags = prep->flags_fwd;
  } else {
    // one launch: interleaved bf16 weights + bias, h_{-1} = 0 (first B rows of hprev), cleared arrival counters
    SNT_CHECK(prep_launch(w_ih, w_hh, b_ih, b_hh, In, H, w.w_ih, w.w_hh, w.bsum, nullptr, hp_b,
                          (int64_t)sizeof(bf) * B * H, w.flags,
                          w.flags ? (int64_t)sizeof(int) * ((B + tc::BM - 1) / tc::BM) * T * PF_FLAGS_PER_STEP : 0, st));
  }
  // the input projection of every timestep as ONE tensor-core contraction: Gx' = x . W_ih'^T + (b_ih + b_hh)'
  SNT_CHECK(tc::gemm_tc(false, false, N, 4 * H, In, 1.f, (const bf*)x, In, w.w_ih, In, 0.f, nullptr, w.gx, 4 * H,
                        w.bsum, 1, nullptr, st));
  CUtensorMap ta, tb;
  SNT_CHECK(tc::make_operand_tmap(&ta, hp_b, false, N, H, H, tc::BM));
  SNT_CHECK(tc::make_operand_tmap(&tb, w.w_hh, false, 4 * H, H, H, 128));
  {
    // persistent path: every (row block, column block) CTA must be co-resident (cooperative launch) and the W_hh'
    // slice must fit in shared memory next to the h ring
    const int num_m0 = (int)((B + tc::BM - 1) / tc::BM), num_n = (int)((4 * H + 127) / 128);
    const int KB = (int)((H + tc::BK - 1) / tc::BK);
    const size_t smem = (size_t)(KB + PF_STAGES) * 16384 + 1024 + 256;
    int coop = 0, max_smem = 0;
    persist_caps(&coop, &max_smem);
    // Sequences are independent, so a batch wider than one co-resident group of row blocks (floor(SMs / num_n): 9 at
    // H = 512, i.e. 1152 sequences) simply takes several launches, each over its own group of 128-row blocks.
    const int group = num_n > 0 ? tc::sm_count() / num_n : 0;
    if (coop == 1 && !getenv("SNT_NO_PERSISTENT") && group >= 1 &&
        smem <= (size_t)max_smem && KB <= PF_FLAGS_PER_STEP && w.flags != nullptr) {
      PersistFwdParams pp;
      pp.H = (int)H; pp.T = T; pp.num_n = num_n; pp.gx = w.gx; pp.cs = cs; pp.hs = hs_b; pp.hprev = hp_b; pp.act = act;
      pp.flags = w.flags; pp.m_base = 0;
#ifdef SNT_LSTM_DBG
      static long long* dbg_dev = nullptr;
      if (!dbg_dev) cudaMalloc(&dbg_dev, sizeof(long long) * SNT_MAX_T * 12);
      cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * SNT_MAX_T * 12, st);
      pp.dbg = dbg_dev;
#endif
      for (int m0 = 0; m0 < num_m0; m0 += group) {
        pp.m_base = m0;
        const int nb = num_m0 - m0 < group ? num_m0 - m0 : group;
        SNT_CHECK(launch_persistent<PersistFwdParams>(lstm_fwd_persistent_kernel<1>, lstm_fwd_persistent_kernel<4>,
                                                      nb * num_n, num_n, smem, ta, tb, pk, pp, st));
      }
#ifdef SNT_LSTM_DBG
      {
        static int printed = 0;
        if (printed++ == 3) {
          long long h[SNT_MAX_T * 12];
          cudaStreamSynchronize(st);
          cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
          const long long t0 = h[5];
          fprintf(stderr, "cols 8: after reader fence, 9: after writer proxy fence\n");
          fprintf(stderr, "[lstm dbg] block %d: step: flag0_seen tma_issued first_full mma_issued tfull_seen h_stored fenced flag_set (cycles since step-0 start)\n", SNT_DBG_BLOCK);
          for (int t = 0; t < T; ++t) {
            fprintf(stderr, "[lstm dbg] %2d:", t);
            for (int k = 0; k < 12; ++k) fprintf(stderr, " %8lld", h[t * 12 + k] > 0 ? h[t * 12 + k] - t0 : -1);
            fprintf(stderr, "\n");
          }
        }
      }
#endif
      return SNT_OK;
    }
  }
  for (int t = 0; t < T; ++t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    tc::TileSched ts;
    ts.num_m = (bs + tc::BM - 1) / tc::BM;
    ts.num_n = (int)((4 * H + 127) / 128);
    ts.splits = 1;
    ts.n_fastest = 0;
    ts.kblocks = (int)((H + tc::BK - 1) / tc::BK);
    ts.kblocks_per_split = ts.kblocks;
    ts.a_row0 = pk.off[t];
    ts.b_row0 = 0;
    LstmFwdEpi e;
    e.bs = bs; e.bs_next = bs_next; e.H = (int)H;
    e.gx = w.gx + (int64_t)pk.off[t] * 4 * H;
    e.c_prev = t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr;
    e.cs = cs + (int64_t)pk.off[t] * H;
    e.hs = hs_b + (int64_t)pk.off[t] * H;
    e.hprev_next = bs_next > 0 ? hp_b + (int64_t)pk.off[t + 1] * H : nullptr;
    e.bias = nullptr; e.ldh = (int)H; e.ldn = (int)H;
    e.act = act + (int64_t)pk.off[t] * 4 * H;
    SNT_CHECK((tc::launch_gemm_tc<128, false, false, LstmFwdEpi>(ta, tb, ts, e, st, /*pdl=*/t > 0)));
  }
  return SNT_OK;
}

// One-shot hook for the step executor: the next lstm_bwd() on this thread records `e` on its stream right BEFORE it
// launches the recurrence (after the weight preparation).  Background work of a lower-priority stream that waits for `e`
// is then released at the same moment as the cooperative BPTT kernel, whose CTAs the block scheduler places first.
static thread_local cudaEvent_t g_bptt_gate = nullptr;
void lstm_bwd_gate_event(cudaEvent_t e) { g_bptt_gate = e; }
bool lstm_bwd_is_persistent(int64_t H) { return (H == 256 || H == 512) && !getenv("SNT_NO_PERSISTENT"); }
// Row blocks per CTA of the persistent BPTT kernel.  Measured at B = 1024 / H = 512 (profiles/r02_bptt_row_blocks.txt):
// one row block per CTA on 128 SMs 188 us; two per CTA on 64 SMs 302 us in order, 277 us phase-split, 317 us with every
// role serving whichever chain is ready first.  A step moves ~290 KB per row block through the SM's L2 port (128 KB of
// dG', 64 KB of partials out and in, the cell inputs), which is already half of the 9.4 us chain, so two chains on one SM
// cannot hide in each other's latency.  Two row blocks per CTA therefore only where the batch would otherwise need
// several launches (more row blocks than fit side by side): one 277 us launch instead of two 188 us ones.
// SNT_PERSIST_RB=1|2 forces either.
static int bptt_row_blocks_per_cta(int num_m0, int group) {
  const char* e = getenv("SNT_PERSIST_RB");
  if (e && atoi(e) == 1) return 1;
  if (e && atoi(e) == 2) return num_m0 >= 2 ? 2 : 1;
  return num_m0 > group ? 2 : 1;
}
// CTAs (= SMs: one CTA per SM) the widest launch of the persistent BPTT recurrence occupies for a batch of B sequences;
// 0 when the recurrence of this hidden size runs as per-step launches.
int lstm_bwd_persistent_ctas(int64_t B, int64_t H) {
  if (!lstm_bwd_is_persistent(H)) return 0;
  const int ns = (int)(H / 128), num_m0 = (int)((B + tc::BM - 1) / tc::BM);
  const int group = tc::sm_count() / (ns * ns);
  if (group < 1) return 0;
  const int rb = bptt_row_blocks_per_cta(num_m0, group);
  const int groups = (num_m0 + rb - 1) / rb;
  return (groups < group ? groups : group) * ns * ns;
}

int lstm_bwd(const PackInfo& pk, const float* d_hs, float* gates, const float* cs, const void* hprev,
             const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh, float* d_w_ih,
             float* d_w_hh, float* d_bias, float* dx, void* ws, int64_t ws_bytes, cudaStream_t st,
             const LstmPrepared* prep, bool* wgrad_pending) {
  if (wgrad_pending) *wgrad_pending = false;
  SNT_REQ8(In, "In");
  SNT_REQ8(H, "H");
  const int T = pk.T;
  const int64_t N = pk.off[T], B = pk.off[1];
  LstmWs w = carve(ws, ws_bytes, N, B, In, H);
  if (!w.ok) { set_error("bf16 lstm_bwd: workspace too small"); return SNT_EWORKSPACE; }
  if (prep) { w.w_ih = prep->w_ih; w.w_hh = prep->w_hh; w.w_hh_t = prep->w_hh_t; w.flags = prep->flags_bwd; }
  const bf* act = (const bf*)gates;
  bf* dg = (bf*)gates + N * 4 * H;
  bool persistent = false;
  {
    const int num_m0 = (int)((B + tc::BM - 1) / tc::BM);
    const int ns = (int)(H / 128);  // K-splits = N tiles; ns*ns CTAs per row block
    const size_t smem = (size_t)(8 + PB_STAGES) * 16384 + 1024 + 256;
    const size_t nflags = (size_t)num_m0 * T * PB_FLAGS_PER_STEP;
    int coop = 0, max_smem = 0;
    persist_caps(&coop, &max_smem);
    const int group = ns > 0 ? tc::sm_count() / (ns * ns) : 0;  // row blocks that are co-resident in one launch
    const bool can_persist = coop == 1 && !getenv("SNT_NO_PERSISTENT") && (H == 256 || H == 512) && group >= 1 &&
                             smem <= (size_t)max_smem && w.flags && w.w_hh_t && w.part;
    // one launch: W_ih' (for dX), and either W_hh'^T + cleared counters (persistent) or W_hh' (per-step path)
    if (!prep)
      SNT_CHECK(prep_launch(w_ih, w_hh, nullptr, nullptr, In, H, w.w_ih, can_persist ? nullptr : w.w_hh, nullptr,
                            can_persist ? w.w_hh_t : nullptr, can_persist ? w.flags : nullptr,
                            (int64_t)sizeof(int) * 2 * nflags, nullptr, 0, st));
    if (g_bptt_gate) {  // see lstm_bwd_gate_event(): released at the moment the recurrence kernel becomes eligible
      SNT_CUDA(cudaEventRecord(g_bptt_gate, st));
      g_bptt_gate = nullptr;
    }
    if (can_persist) {
      CUtensorMap ta, tb;
      SNT_CHECK(tc::make_operand_tmap(&ta, dg, false, N, 4 * H, 4 * H, tc::BM));
      SNT_CHECK(tc::make_operand_tmap(&tb, w.w_hh_t, false, H, 4 * H, 4 * H, 128));
      PersistBwdParams pp;
      pp.H = (int)H; pp.T = T; pp.ns = ns; pp.d_hs = d_hs; pp.act = act; pp.cs = cs; pp.dg = dg;
      pp.part = w.part; pp.dgflag = w.flags; pp.pflag = w.flags + nflags;
#ifdef SNT_LSTM_DBG
      static long long* dbg_dev = nullptr;
      if (!dbg_dev) cudaMalloc(&dbg_dev, sizeof(long long) * SNT_MAX_T * 12);
      cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * SNT_MAX_T * 12, st);
      pp.dbg = dbg_dev;
#endif
      // rb row blocks per CTA (see the kernel): a launch covers up to group * rb row blocks
      const int rb = bptt_row_blocks_per_cta(num_m0, group);
      for (int m0 = 0; m0 < num_m0; m0 += group * rb) {
        pp.m_base = m0;
        const int nb = num_m0 - m0 < group * rb ? num_m0 - m0 : group * rb;
        const int ctas = ((nb + rb - 1) / rb) * ns * ns;
        // two row blocks per CTA: the phase-split kernel (SNT_BPTT_MODE=seq: one epilogue group walks both phases)
        const char* me = getenv("SNT_BPTT_MODE");
        const bool ps = rb == 2 && !(me && !strcmp(me, "seq"));
        void (*kern)(CUtensorMap, CUtensorMap, PackInfo, PersistBwdParams);
        if (ns == 4)
          kern = rb == 1 ? lstm_bwd_persistent_kernel<4, 1>
                         : ps ? lstm_bwd_persistent_ps_kernel<4, 2> : lstm_bwd_persistent_kernel<4, 2>;
        else
          kern = rb == 1 ? lstm_bwd_persistent_kernel<2, 1>
                         : ps ? lstm_bwd_persistent_ps_kernel<2, 2> : lstm_bwd_persistent_kernel<2, 2>;
        SNT_CHECK(launch_persistent<PersistBwdParams>(kern, nullptr, ctas, 1, smem, ta, tb, pk, pp, st,
                                                      ps ? PS_THREADS : PF_THREADS));
      }
#ifdef SNT_LSTM_DBG
      {
        static int printed = 0;
        if (printed++ == 3) {
          long long h[SNT_MAX_T * 12];
          cudaStreamSynchronize(st);
          cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
          const long long t0 = h[8];
          fprintf(stderr, "[lstm bwd dbg] block %d: s: ready0_seen tma_issued first_full mma_issued tfull_seen part_stored pflag_set parts_ready dg_stored dgflag_set\n", SNT_DBG_BLOCK);
          for (int t = 0; t < T; ++t) {
            fprintf(stderr, "[lstm bwd dbg] %2d:", t);
            for (int k = 0; k < 10; ++k) fprintf(stderr, " %8lld", h[t * 12 + k] > 0 ? h[t * 12 + k] - t0 : -1);
            fprintf(stderr, "\n");
          }
        }
      }
#endif
      persistent = true;
    }
  }
  // Per-step path: split-K keeps all SMs busy on the small per-step contraction; its partials never leave L2.  (Tried in
  // round 2: the cell backward as the EPILOGUE of an unsplit contraction - one launch per step instead of two.  At
  // configs[3] it was 235 us per training step SLOWER: with one tile per CTA nothing overlaps the epilogue, and its
  // 36 bytes per hidden unit of row-per-lane traffic run at a single SM's bandwidth instead of the whole GPU's.)
  const int want_splits = 4;
  for (int t = T - 1; t >= 0 && !persistent; --t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    int used = 1;
    if (bs_next > 0) {  // partial[s] = dG'_{t+1}[:, K_s] . W_hh'[K_s, :]   (B operand MN-major)
      const int64_t cap = (int64_t)MAX_SPLITS * 4 * H * (In > H ? In : H);  // floats in w.sws
      int want = want_splits;
      while (want > 1 && (int64_t)want * bs_next * H > cap) --want;
      SNT_CHECK(tc::gemm_tc(false, true, bs_next, H, 4 * H, 1.f, dg + (int64_t)pk.off[t + 1] * 4 * H, 4 * H, w.w_hh, H,
                            0.f, w.sws, nullptr, H, nullptr, want, w.sws, st, 0, nullptr, /*keep_partials=*/true,
                            &used));
    }
    const int64_t threads = (int64_t)bs * (H / 8);
    lstm_bwd_point_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        bs, bs_next, (int)H, d_hs + (int64_t)pk.off[t] * H, act + (int64_t)pk.off[t] * 4 * H,
        cs + (int64_t)pk.off[t] * H, t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr, w.sws, used,
        (int64_t)bs_next * H, w.dc_state, dg + (int64_t)pk.off[t] * 4 * H);
    SNT_LAUNCH_CHECK("lstm_bwd_point_kernel");
  }
  // Tail of the stage: dW_ih = dG'^T.X, dW_hh = dG'^T.Hprev (rows come out interleaved: un-permuted on store), the bias
  // gradient (column sums of dG') and dX = dG'.W_ih'.  Each weight gradient has too few output tiles to fill the GPU and
  // needs a K split; when both fit next to each other (H <= 512: 16 + 32 tiles of 128 x 256) they run CONCURRENTLY on the
  // two streams with grids sized to share the SMs in proportion to their work (48 + 96 CTAs at E256/H512, three K
  // slices each) instead of one after the other at 128 CTAs each.  The column sums follow dW_ih on the side stream, next
  // to dX.
  SideStream* side = side_stream();
  const int sms = tc::grid_sms();
  const int64_t rt = (4 * H + tc::BM - 1) / tc::BM;
  const int64_t t1 = rt * ((In + 255) / 256), t2 = rt * ((H + 255) / 256);
  const int64_t kb_n = (N + tc::BK - 1) / tc::BK;
  int c1 = 0, c2 = 0;  // K slices of the concurrent mode (0: sequential)
  if (side && In >= 256 && H >= 256 && t1 + t2 <= sms && kb_n >= 16 && !getenv("SNT_NO_WGRAD_OVERLAP")) {
    const double share1 = (double)In / (double)(In + H);
    c1 = (int)((double)sms * share1 / (double)t1);
    c2 = (int)((double)(sms - (c1 < 1 ? 1 : c1) * t1) / (double)t2);
    if (c1 < 1) c1 = 1;
    if (c2 < 1) c2 = 1;
    if (c1 > MAX_SPLITS) c1 = MAX_SPLITS;
    if (c2 > MAX_SPLITS) c2 = MAX_SPLITS;
    while (c1 > 1 && kb_n / c1 < 4) --c1;
    while (c2 > 1 && kb_n / c2 < 4) --c2;
    if ((int64_t)c1 * 4 * H * In + (int64_t)c2 * 4 * H * H > (int64_t)MAX_SPLITS * 4 * H * (In > H ? In : H)) c1 = c2 = 0;
  }
  if (side) {
    SNT_CUDA(cudaEventRecord(side->fork, st));
    SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
  }
  // Deferred mode (the step executor, last layer of the backward pass): only dX is on the caller's critical path - the
  // embedding gradient and the head backward wait for it, nothing waits for the weight gradients before the optimizer.
  // dX goes first on the caller's stream; both weight-gradient contractions and the bias column sums run on side streams
  // beside whatever the caller enqueues next (a dozen small launch-bound kernels), and the caller joins them later.
  SideStream* side2 = wgrad_pending ? side_stream(2) : nullptr;
  if (wgrad_pending && side && side2 && c1 > 0 && dx && !getenv("SNT_NO_WGRAD_DEFER")) {
    SNT_CHECK(tc::gemm_tc(false, true, N, In, 4 * H, 1.f, dg, 4 * H, w.w_ih, In, 0.f, dx, nullptr, In, nullptr, 1,
                          nullptr, st));
    SNT_CUDA(cudaStreamWaitEvent(side2->s, side->fork, 0));
    float* sws2 = w.sws + align_up((int64_t)c1 * 4 * H * In, 64);
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, In, N, 1.f, dg, 4 * H, (const bf*)x, In, 0.f, d_w_ih, nullptr, In,
                          nullptr, c1, w.sws, side->s, (int)H, nullptr, false, nullptr, 256));
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, H, N, 1.f, dg, 4 * H, (const bf*)hprev, H, 0.f, d_w_hh, nullptr, H,
                          nullptr, c2, sws2, side2->s, (int)H, nullptr, false, nullptr, 256));
    SNT_CHECK(colsum_bf16(dg, N, 4 * H, 4 * H, 0.f, w.tmp, w.cpart, side->s));
    unperm_vec_kernel<<<(unsigned)((4 * H + 255) / 256), 256, 0, side->s>>>(w.tmp, (int)H, d_bias);
    SNT_LAUNCH_CHECK("unperm_vec_kernel");
    SNT_CUDA(cudaEventRecord(side->aux2, side->s));
    SNT_CUDA(cudaEventRecord(side2->aux2, side2->s));
    *wgrad_pending = true;
    return SNT_OK;
  }
  if (c1 > 0) {
    float* sws2 = w.sws + align_up((int64_t)c1 * 4 * H * In, 64);
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, In, N, 1.f, dg, 4 * H, (const bf*)x, In, 0.f, d_w_ih, nullptr, In,
                          nullptr, c1, w.sws, side->s, (int)H, nullptr, false, nullptr, 256));
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, H, N, 1.f, dg, 4 * H, (const bf*)hprev, H, 0.f, d_w_hh, nullptr, H,
                          nullptr, c2, sws2, st, (int)H, nullptr, false, nullptr, 256));
  } else {
    int s1 = tc::choose_splits(4 * H, In, N, 0), s2 = tc::choose_splits(4 * H, H, N, 0);
    if (s1 > MAX_SPLITS) s1 = MAX_SPLITS;
    if (s2 > MAX_SPLITS) s2 = MAX_SPLITS;
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, In, N, 1.f, dg, 4 * H, (const bf*)x, In, 0.f, d_w_ih, nullptr, In,
                          nullptr, s1, w.sws, st, (int)H));
    SNT_CHECK(tc::gemm_tc(true, true, 4 * H, H, N, 1.f, dg, 4 * H, (const bf*)hprev, H, 0.f, d_w_hh, nullptr, H,
                          nullptr, s2, w.sws, st, (int)H));
  }
  // bias gradient = column sums of dG': bandwidth-bound, on the side stream
  {
    cudaStream_t cs = side ? side->s : st;
    SNT_CHECK(colsum_bf16(dg, N, 4 * H, 4 * H, 0.f, w.tmp, w.cpart, cs));
    unperm_vec_kernel<<<(unsigned)((4 * H + 255) / 256), 256, 0, cs>>>(w.tmp, (int)H, d_bias);
    SNT_LAUNCH_CHECK("unperm_vec_kernel");
  }
  if (side) SNT_CUDA(cudaEventRecord(side->join, side->s));
  if (dx)
    SNT_CHECK(tc::gemm_tc(false, true, N, In, 4 * H, 1.f, dg, 4 * H, w.w_ih, In, 0.f, dx, nullptr, In, nullptr, 1,
                          nullptr, st));
  if (side) SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  return SNT_OK;
}

int lstm_bwd_join(cudaStream_t st) {
  SideStream* side = side_stream();
  SideStream* side2 = side_stream(2);
  SNT_REQUIRE(side && side2, "lstm_bwd_join: no side streams");
  SNT_CUDA(cudaStreamWaitEvent(st, side->aux2, 0));
  SNT_CUDA(cudaStreamWaitEvent(st, side2->aux2, 0));
  return SNT_OK;
}

// =====================================================================================================================
// a12: greedy decode (models.py:56-67) on the same fused kernels.  Per step and layer: one tcgen05 GEMM for the input
// projection (bf16 out), one for the recurrence with the gate epilogue; then the vocab projection whose epilogue
// reduces every 128-column slab of a row to (max, first index) — logits never leave TMEM — and a finishing kernel
// that picks the first global maximum, writes the token id and gathers the next input embedding.
// =====================================================================================================================
struct ArgmaxEpi {
  static constexpr int kWarps = 8;
  static constexpr int kStages = 0;
  static constexpr int kSmemPerWarp = 4 * 32 * 4;  // this warp's 128 bias values (4 chunks of 32 columns)
  int M, V;
  const float* bias;
  float2* part;  // [slabs][M]: (max logit, index as int bits)

  // The 128 bias values of the warp's half tile: four coalesced loads per lane one tile ahead, handed to every lane
  // through shared memory (broadcast reads).  Loading bias[col] right where it is added put a dependent global load in
  // front of every FADD: ncu's source page had 40 % of the kernel's stall samples on those adds (long scoreboard).
  struct Pre { float b[4]; };
  __device__ __forceinline__ void prefetch(Pre& p, int, int n_blk, int ew, int lane) const {
    const int c0 = n_blk * 256 + (ew >> 2) * 128 + lane;
#pragma unroll
    for (int c = 0; c < 4; ++c) p.b[c] = c0 + c * 32 < V ? __ldg(bias + c0 + c * 32) : 0.f;
  }
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int, int ew, int lane,
                                       const Pre& pre, uint8_t* wsm) const {
    const int half = ew >> 2;
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    const uint32_t bsa = tc::smem_u32(wsm);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bsa + (uint32_t)(c * 128 + lane * 4)), "f"(pre.b[c]) : "memory");
    __syncwarp();
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int cofs = half * 128 + c * 32;
      const int col0 = n_blk * 256 + cofs;
      if (col0 >= V) break;  // warp-uniform
      uint32_t r[32];
      tc::tmem_ld32(tmem_rows + (uint32_t)cofs, r);
      float bv[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 u = tc::lds128(bsa + (uint32_t)(c * 128 + q * 16));
        bv[4 * q] = __uint_as_float(u.x); bv[4 * q + 1] = __uint_as_float(u.y);
        bv[4 * q + 2] = __uint_as_float(u.z); bv[4 * q + 3] = __uint_as_float(u.w);
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        if (col < V) {
          const float x = __uint_as_float(r[j]) + bv[j];
          if (x > best || bi == 0x7fffffff) { best = x; bi = col; }  // strict '>' keeps the first maximum
        }
      }
    }
    __syncwarp();  // the next tile's bias values overwrite the stage
    if (row < M) part[(int64_t)(n_blk * 2 + half) * M + row] = make_float2(best, __int_as_float(bi));
  }
};

// First global maximum over the slab partials, id out, next input = bf16(W_emb[id]).  Block = 32 rows x 8 warps: while
// scanning, lane = row and warp w takes slabs w, w + 8, ... (consecutive rows of a slab are contiguous: coalesced 256-byte
// reads, ~10 independent loads per thread, all in flight); the 8 candidates of a row are combined through shared memory
// (greater value, lower index on ties = the first maximum), then every warp copies 4 of the 32 embedding rows.  Launched
// with programmatic stream serialization: the prologue of the next step's contraction overlaps this kernel.
__global__ void __launch_bounds__(256)
argmax_finish_kernel(const float2* __restrict__ part, int slabs, int64_t B, const float* __restrict__ w_emb, int E,
                     int64_t* __restrict__ ids, int64_t ids_stride, bf* __restrict__ x_next, int64_t ldx) {
  __shared__ float sval[8][32];
  __shared__ int sidx[8][32];
  __shared__ int sbest[32];
  tc::pdl_launch_dependents();
  tc::pdl_wait();  // the slab partials come from the preceding contraction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  const int64_t row = row0 + lane;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (row < B) {
#pragma unroll 4
    for (int k = warp; k < slabs; k += 8) {  // ascending column ranges: strict '>' keeps the first maximum
      const float2 p = __ldcg(part + (int64_t)k * B + row);
      const int idx = __float_as_int(p.y);
      if (idx != 0x7fffffff && (p.x > best || bi == 0x7fffffff)) { best = p.x; bi = idx; }
    }
  }
  sval[warp][lane] = best;
  sidx[warp][lane] = bi;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float ob = sval[w][lane];
      const int oi = sidx[w][lane];
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    sbest[lane] = bi;
    if (row < B) ids[row * ids_stride] = bi;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = warp * 4 + q;
    if (row0 + r >= B) break;
    const float* src = w_emb + (int64_t)sbest[r] * E;
    bf* dst = x_next + (row0 + r) * ldx;
    for (int e = lane * 4; e < E; e += 128) {  // E % 8 == 0
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + e));
      uint2 o;
      o.x = pack_bf2(v.x, v.y);
      o.y = pack_bf2(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + e) = o;
    }
  }
}

// fp32 [rows, cols] -> bf16 rows of pitch ld (the [x | h] operand buffers of the decode loop)
__global__ void __launch_bounds__(256)
cast_rows_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, bf* __restrict__ dst, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols, c = i % cols;
  dst[r * ld + c] = __float2bfloat16_rn(src[i]);
}

// Per layer the step's two contractions (input projection and recurrence) are ONE GEMM over the concatenated operand
// [x_s | h_{s-1}] (K = in + H) against [W_ih' | W_hh']: the gate epilogue writes h_s straight into the h half of the
// next step's operand buffer (and into the x half of the next layer's), argmax_finish writes the next token's embedding
// into the x half, and the vocab projection reads h_s in place through a strided tensor map - no copies, no Gx' buffer.
int64_t greedy_ws_bytes(int64_t B, int64_t E, int64_t H, int64_t V, int L) {
  const int64_t slabs = ((V + 255) / 256) * 2;
  int64_t b = ws_bytes_for(V * H, 2) + ws_bytes_for(slabs * B, 8);
  for (int k = 0; k < L; ++k) {
    const int64_t in = k == 0 ? E : H;
    b += 2 * ws_bytes_for(B * (in + H), 2) + ws_bytes_for(B * H, 4) + ws_bytes_for(4 * H, 4) +
         ws_bytes_for(4 * H * (in + H), 2);
  }
  return b;
}

int greedy_decode(const float* features, const float* w_emb, int L, const float* const* w_ih,
                  const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                  const float* w_out, const float* b_out, const float* h0, const float* c0, int64_t B, int64_t E,
                  int64_t H, int64_t V, int steps, int64_t* ids, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQ8(E, "E");
  SNT_REQ8(H, "H");
  SNT_REQUIRE(B < (1LL << 31) && V < (1LL << 31), "greedy_decode: extent too large");
  const int slabs = (int)((V + 255) / 256) * 2;
  Workspace w(ws, ws_bytes);
  bf* wout = w.take<bf>(V * H);
  float2* part = w.take<float2>((int64_t)slabs * B);
  bf *cat[SNT_MAX_LAYERS][2], *wcat[SNT_MAX_LAYERS];
  float *c[SNT_MAX_LAYERS], *bsum[SNT_MAX_LAYERS];
  int64_t kin[SNT_MAX_LAYERS];
  for (int k = 0; k < L; ++k) {
    kin[k] = k == 0 ? E : H;
    cat[k][0] = w.take<bf>(B * (kin[k] + H));
    cat[k][1] = w.take<bf>(B * (kin[k] + H));
    c[k] = w.take<float>(B * H);
    bsum[k] = w.take<float>(4 * H);
    wcat[k] = w.take<bf>(4 * H * (kin[k] + H));
  }
  if (!w.ok()) { set_error("bf16 greedy_decode: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(w_out, wout, V * H, st));
  CUtensorMap ta[SNT_MAX_LAYERS][2], tw[SNT_MAX_LAYERS], tout_a[2], tout_b;
  for (int k = 0; k < L; ++k) {
    const int64_t ld = kin[k] + H;
    // [W_ih' | W_hh'] interleaved rows + summed bias; the step-0 operand buffer starts as zeros (h_{-1} = 0)
    SNT_CHECK(prep_launch(w_ih[k], w_hh[k], b_ih[k], b_hh[k], kin[k], H, wcat[k], wcat[k] + kin[k], bsum[k], nullptr,
                          cat[k][0], (int64_t)sizeof(bf) * B * ld, nullptr, 0, st, ld, ld));
    if (h0) {
      cast_rows_kernel<<<(unsigned)((B * H + 255) / 256), 256, 0, st>>>(h0 + (int64_t)k * B * H, B, H, cat[k][0] + kin[k], ld);
      SNT_LAUNCH_CHECK("cast_rows_kernel");
      SNT_CUDA(cudaMemcpyAsync(c[k], c0 + (int64_t)k * B * H, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
    } else {
      SNT_CUDA(cudaMemsetAsync(c[k], 0, sizeof(float) * B * H, st));
    }
    for (int q = 0; q < 2; ++q) SNT_CHECK(tc::make_operand_tmap(&ta[k][q], cat[k][q], false, B, ld, ld, tc::BM));
    SNT_CHECK(tc::make_operand_tmap(&tw[k], wcat[k], false, 4 * H, ld, ld, 128));
  }
  cast_rows_kernel<<<(unsigned)((B * E + 255) / 256), 256, 0, st>>>(features, B, E, cat[0][0], E + H);
  SNT_LAUNCH_CHECK("cast_rows_kernel");
  {
    const int64_t ld = kin[L - 1] + H;
    for (int q = 0; q < 2; ++q)
      SNT_CHECK(tc::make_operand_tmap(&tout_a[q], cat[L - 1][q] + kin[L - 1], false, B, H, ld, tc::BM));
  }
  SNT_CHECK(tc::make_operand_tmap(&tout_b, wout, false, V, H, H, 256));
  for (int s = 0; s < steps; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;  // cat[k][cur] = [input_s | h_{s-1}]; h_s goes to cat[k][nxt]
    for (int k = 0; k < L; ++k) {
      const int64_t ld = kin[k] + H;
      tc::TileSched ts;
      ts.num_m = (int)((B + tc::BM - 1) / tc::BM);
      ts.num_n = (int)((4 * H + 127) / 128);
      ts.splits = 1;
      ts.n_fastest = 0;
      ts.kblocks = (int)((ld + tc::BK - 1) / tc::BK);
      ts.kblocks_per_split = ts.kblocks;
      ts.a_row0 = 0;
      ts.b_row0 = 0;
      LstmFwdEpi e;
      e.bs = (int)B; e.H = (int)H; e.gx = nullptr; e.bias = bsum[k]; e.c_prev = c[k]; e.cs = c[k];
      e.hs = cat[k][nxt] + kin[k]; e.ldh = (int)ld;
      e.hprev_next = k + 1 < L ? cat[k + 1][cur] : nullptr;  // this step's input of the next layer
      e.bs_next = k + 1 < L ? (int)B : 0;
      e.ldn = k + 1 < L ? (int)(kin[k + 1] + H) : 0;
      e.act = nullptr;
      SNT_CHECK((tc::launch_gemm_tc<128, false, false, LstmFwdEpi>(ta[k][cur], tw[k], ts, e, st, true)));
    }
    tc::TileSched ts;
    ts.num_m = (int)((B + tc::BM - 1) / tc::BM);
    ts.num_n = (int)((V + 255) / 256);
    ts.splits = 1;
    ts.n_fastest = 0;
    ts.kblocks = (int)((H + tc::BK - 1) / tc::BK);
    ts.kblocks_per_split = ts.kblocks;
    ts.a_row0 = 0;
    ts.b_row0 = 0;
    ArgmaxEpi e;
    e.M = (int)B; e.V = (int)V; e.bias = b_out; e.part = part;
    SNT_CHECK((tc::launch_gemm_tc<256, false, false, ArgmaxEpi>(tout_a[nxt], tout_b, ts, e, st, true)));
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((B + 31) / 32));
      cfg.blockDim = dim3(256);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      SNT_CUDA(cudaLaunchKernelEx(&cfg, argmax_finish_kernel, (const float2*)part, slabs, B, w_emb, (int)E, ids + s,
                                  (int64_t)steps, cat[0][nxt], E + H));
      count_launch();
    }
  }
  return SNT_OK;
}

}  // namespace bf16
}  // namespace snt
