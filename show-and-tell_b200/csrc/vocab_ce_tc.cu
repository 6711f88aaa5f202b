// a8+a9 on tensor cores: vocab Linear fused with log-softmax + cross-entropy (models.py:53 + train.py:53,143).
//
// forward : one tcgen05 pass logits = Hs . W_out^T whose epilogue reduces each 128-column slab of a row to an
//           online-softmax partial (max, sum-exp) straight out of TMEM — logits are never written anywhere.
//           A finishing kernel merges the partials into lse[n] and the mean NLL.
// backward: the same contraction is recomputed per chunk of rows; its epilogue forms
//           dlogits = (softmax - onehot) * scale in registers and stores it as bf16 into an L2-resident chunk
//           buffer, which two more tcgen05 contractions consume (dHs = dlogits.W_out, dW_out += dlogits^T.Hs).
#include "bf16.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

#include <stdlib.h>

namespace snt {
namespace bf16 {

typedef __nv_bfloat16 bf;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int CE_BN = 256;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- forward epilogue: per (row, 128-column slab) online-softmax partial in the log2 domain ------------------------
struct CeFwdEpi {
  static constexpr int kWarps = 8;
  static constexpr int kStages = 0;
  static constexpr int kSmemPerWarp = 0;
  int M;                 // rows
  int V;                 // valid columns
  const float* bias;     // [V]
  float2* part;          // [num_slabs][M] (max2, sum2): sum_j 2^(y_j - max2), y = logit * log2(e)
  const int64_t* targets;  // [M]
  float* tl;             // [M] out: logit of the target class, picked out of the accumulator by the one thread whose
                         // chunk holds it (rows with an out-of-range target are left alone; ce_finish flags them)

  struct Pre { int tgt; };
  __device__ __forceinline__ void prefetch(Pre& p, int m_blk, int, int ew, int lane) const {
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    int64_t t = row < M ? targets[row] : -1;
    p.tgt = (t >= 0 && t < V) ? (int)t : -1;
  }
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int split, int ew, int lane,
                                       const Pre& pre, uint8_t* wsm) const {
    const int half = ew >> 2;
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    float m = -INFINITY, s = 0.f;
#pragma unroll 1
    for (int c = 0; c < CE_BN / 64; ++c) {
      const int cofs = half * (CE_BN / 2) + c * 32;
      const int col0 = n_blk * CE_BN + cofs;
      if (col0 >= V) break;  // warp-uniform
      uint32_t r[32];
      tc::tmem_ld32(tmem_rows + (uint32_t)cofs, r);
      float4 bv[8];  // bias of the chunk, fetched while the TMEM load is in flight
      if (col0 + 32 <= V) {
#pragma unroll
        for (int q = 0; q < 8; ++q) bv[q] = __ldg(reinterpret_cast<const float4*>(bias + col0) + q);
      }
      tc::tmem_ld_wait();
      {
        const int trel = pre.tgt - col0;
        const bool hit = pre.tgt >= 0 && trel >= 0 && trel < 32;
        if (__any_sync(0xffffffffu, hit)) {  // ~10 % of the chunks: some row of the warp has its target in here
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v = (trel == j) ? __uint_as_float(r[j]) : v;
          if (hit) tl[row] = v + __ldg(bias + pre.tgt);
        }
      }
      float y[32];
      if (col0 + 32 <= V) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = bv[j >> 2];
          y[j] = (__uint_as_float(r[j]) + b.x) * LOG2E;
          y[j + 1] = (__uint_as_float(r[j + 1]) + b.y) * LOG2E;
          y[j + 2] = (__uint_as_float(r[j + 2]) + b.z) * LOG2E;
          y[j + 3] = (__uint_as_float(r[j + 3]) + b.w) * LOG2E;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          y[j] = (col0 + j < V) ? (__uint_as_float(r[j]) + bias[col0 + j]) * LOG2E : -INFINITY;
      }
      float cm = y[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) cm = fmaxf(cm, y[j]);
      const float mn = fmaxf(m, cm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += ex2(y[j] - mn);
      s = s * ex2(m - mn) + acc;  // m = -inf on the first chunk: ex2(-inf) = 0
      m = mn;
    }
    if (row < M) part[(int64_t)(n_blk * 2 + half) * M + row] = make_float2(m, s);
  }
};

// lse[n] = ln2 * (M + log2 sum_k s_k 2^(m_k - M)); nll[n] = lse - tl; loss = mean(nll).  Block = 64 rows x 8 slab
// groups: every thread merges its share of the slab partials online (coalesced across rows), the 8 groups are merged
// through shared memory in a fixed order.  Each block leaves the sum of its 64 nll values in block_sum[]; the block that
// takes the last ticket adds those in index order (deterministic) and writes the mean - no separate reduction launch.
__global__ void __launch_bounds__(512)
ce_finish_kernel(const float2* __restrict__ part, int slabs, int64_t N, const float* __restrict__ tl,
                 const int64_t* __restrict__ targets, int64_t V, float* __restrict__ lse, float* __restrict__ block_sum,
                 int* __restrict__ ticket, float inv_n, float* __restrict__ loss, int* flags) {
  __shared__ float2 red[8][64];
  __shared__ float wred[16];
  __shared__ int s_last;
  const int tx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t row = (int64_t)blockIdx.x * 64 + tx;
  float M = -INFINITY, S = 0.f;
  if (row < N) {
    for (int k = g; k < slabs; k += 8) {
      const float2 p = part[(int64_t)k * N + row];
      if (p.y > 0.f) {
        const float mn = fmaxf(M, p.x);
        S = S * exp2f(M - mn) + p.y * exp2f(p.x - mn);
        M = mn;
      }
    }
  }
  red[g][tx] = make_float2(M, S);
  __syncthreads();
  float nll = 0.f;
  if (g == 0 && row < N) {
    float Mt = -INFINITY, St = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 p = red[i][tx];
      if (p.y > 0.f) {
        const float mn = fmaxf(Mt, p.x);
        St = St * exp2f(Mt - mn) + p.y * exp2f(p.x - mn);
        Mt = mn;
      }
    }
    const float l = (Mt + log2f(St)) * LN2;
    lse[row] = l;
    const int64_t t = targets[row];
    const bool valid = t >= 0 && t < V;
    if (!valid) atomicOr(flags, 2);
    nll = l - (valid ? tl[row] : 0.f);
  }
  // block sum of the 64 nll values (threads 0..63 = warps 0 and 1), fixed order
  if (g == 0) {
    const float w = warp_sum(nll);
    if ((tx & 31) == 0) wred[tx >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    block_sum[blockIdx.x] = wred[0] + wred[1];
    __threadfence();
    s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 512) t += __ldcg(block_sum + i);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) tot += wred[i];
      loss[0] = tot * inv_n;
      *ticket = 0;  // self-resetting
    }
  }
}

// ---- backward epilogue: dlogits = (2^(y - lse2) - onehot) * scale -> bf16 chunk ------------------------------------
template <int W>
struct CeBwdEpiT {
  static constexpr int kWarps = W;
  static constexpr int kStages = W == 16 ? 3 : 0;  // 16 staging buffers (40 KB) fit next to 3 ring stages
  static constexpr int kSmemPerWarp = 32 * 64;  // swizzled 32 x 64-byte transpose stage for coalesced stores (tc::stage_addr)
  int M;                    // rows in this chunk
  int V;
  const float* bias;        // [V]
  const float* lse;         // [M] (chunk-local pointer)
  const int64_t* targets;   // [M]
  bf* out;                  // [M, ldo] bf16: softmax - onehot, UNSCALED (the 1/N and dloss factors are folded
                            // into the consumers' alpha: scaling first would round every target entry the same way)
  int64_t ldo;

  struct Pre { float l2; int tgt; };
  __device__ __forceinline__ void prefetch(Pre& p, int m_blk, int, int ew, int lane) const {
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    const bool row_ok = row < M;
    p.l2 = row_ok ? lse[row] * LOG2E : 0.f;
    p.tgt = row_ok ? (int)targets[row] : -1;
  }
  // one 32-column chunk: softmax - onehot from the accumulator registers, bf16, staged through shared memory so that each
  // store instruction writes 8 whole 64-byte row segments (4 lanes per row) instead of 32 scattered 16-byte pieces
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], const float4 (&bv)[8], int col0, int row0, int lane,
                                        float l2, int tgt, uint32_t wsa) const {
    const int trel = tgt - col0;
    if (col0 + 32 <= V) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = bv[j >> 2];
        float p0 = ex2((__uint_as_float(r[j]) + b.x) * LOG2E - l2);
        float p1 = ex2((__uint_as_float(r[j + 1]) + b.y) * LOG2E - l2);
        float p2 = ex2((__uint_as_float(r[j + 2]) + b.z) * LOG2E - l2);
        float p3 = ex2((__uint_as_float(r[j + 3]) + b.w) * LOG2E - l2);
        if (trel == j) p0 -= 1.f;
        if (trel == j + 1) p1 -= 1.f;
        if (trel == j + 2) p2 -= 1.f;
        if (trel == j + 3) p3 -= 1.f;
        __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
        pk[j / 2] = *reinterpret_cast<uint32_t*>(&lo);
        pk[j / 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        tc::sts128(tc::stage_addr(wsa, lane, q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      __syncwarp();
      uint4 v[4];  // all four loads first, then the stores: no load waits behind a store
#pragma unroll
      for (int it = 0; it < 4; ++it) v[it] = tc::lds128(tc::stage_addr(wsa, it * 8 + (lane >> 2), lane & 3));
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + (lane >> 2), cq = lane & 3;
        if (row0 + rr < M) *reinterpret_cast<uint4*>(out + (int64_t)(row0 + rr) * ldo + col0 + cq * 8) = v[it];
      }
      __syncwarp();
    } else if (row0 + lane < M) {
      bf* orow = out + (int64_t)(row0 + lane) * ldo;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        if (col < V) {
          float p = ex2((__uint_as_float(r[j]) + bias[col]) * LOG2E - l2);
          if (trel == j) p -= 1.f;
          orow[col] = __float2bfloat16_rn(p);
        }
      }
    }
  }
  __device__ __forceinline__ void load_bias(float4 (&bv)[8], int col0) const {
    if (col0 + 32 <= V) {
#pragma unroll
      for (int q = 0; q < 8; ++q) bv[q] = __ldg(reinterpret_cast<const float4*>(bias + col0) + q);
    }
  }
  // 4 chunks per warp; the TMEM load and the bias fetch of chunk c+1 are in flight while chunk c is processed
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int split, int ew, int lane,
                                       const Pre& pre, uint8_t* wsm) const {
    const int half = ew >> 2;
    const int row0 = m_blk * tc::BM + (ew & 3) * 32;
    const uint32_t wsa = tc::smem_u32(wsm);
    if (W == 16) {  // 16 warps x 2 chunks, no register double-buffering: latency is hidden by warps instead
      const int cb = half * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = n_blk * CE_BN + cb + c * 32;
        if (col0 >= V) break;  // warp-uniform
        uint32_t r[32];
        float4 b8[8];
        tc::tmem_ld32(tmem_rows + (uint32_t)(cb + c * 32), r);
        load_bias(b8, col0);
        tc::tmem_ld_wait();
        chunk(r, b8, col0, row0, lane, pre.l2, pre.tgt, wsa);
      }
      return;
    }
    const int cbase = half * (CE_BN / 2);
    const int colb = n_blk * CE_BN + cbase;
    constexpr int NCH = CE_BN / 64;
    static_assert(NCH % 2 == 0, "chunk loop is unrolled by two");
    uint32_t ra[32], rb[32];
    float4 ba[8], bb[8];
    if (colb >= V) return;  // warp-uniform
    tc::tmem_ld32(tmem_rows + (uint32_t)cbase, ra);
    load_bias(ba, colb);
#pragma unroll 1
    for (int c = 0; c < NCH; c += 2) {
      const int col0 = colb + c * 32;
      if (col0 >= V) break;  // warp-uniform
      tc::tmem_ld_wait();
      const bool more1 = col0 + 32 < V;
      if (more1) {
        tc::tmem_ld32(tmem_rows + (uint32_t)(cbase + (c + 1) * 32), rb);
        load_bias(bb, col0 + 32);
      }
      chunk(ra, ba, col0, row0, lane, pre.l2, pre.tgt, wsa);
      if (!more1) break;
      tc::tmem_ld_wait();
      if (c + 2 < NCH && col0 + 64 < V) {
        tc::tmem_ld32(tmem_rows + (uint32_t)(cbase + (c + 2) * 32), ra);
        load_bias(ba, col0 + 64);
      }
      chunk(rb, bb, col0 + 32, row0, lane, pre.l2, pre.tgt, wsa);
    }
  }
};

typedef CeBwdEpiT<8> CeBwdEpi;

// ---- column sums of a bf16 matrix (for d_b_out), deterministic two pass -----------------------------------------------
constexpr int CSB_ROWS = 256;
__global__ void colsum_bf16_final_kernel(const float* __restrict__ partial, int64_t chunks, int64_t C, float beta,
                                         float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int64_t k = 0; k < chunks; ++k) s += partial[k * C + c];
  out[c] = (beta != 0.f ? beta * out[c] : 0.f) + s;
}
// Partial sums per (256-column slab, 256-row chunk) item.  Block = 32 column groups (8 bf16 = one 16-byte load each) x 32
// row lanes, 8 independent loads in flight per thread (128 KB per block).  The grid walks the items with a stride, so the
// same kernel (same summation order, bit-identical results) runs either wide - one block per item - or NARROW: a few
// blocks on the SMs a cooperative persistent kernel leaves free (the BPTT recurrence occupies 128 of 148 SMs and is
// latency-bound, HBM idle), instead of next to the contractions that read the same matrix.
__global__ void __launch_bounds__(1024)
colsum_bf16_partial_kernel(const bf* __restrict__ in, int64_t R, int64_t C, int64_t ld, float* __restrict__ partial,
                          const float* __restrict__ roww, int slabs, int chunks) {
  __shared__ float red[32][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int item = blockIdx.x; item < slabs * chunks; item += gridDim.x) {
    const int slab = item % slabs, chunk = item / slabs;
    const int64_t c = ((int64_t)slab * 32 + tx) * 8;
    const int64_t r0 = (int64_t)chunk * CSB_ROWS, r1 = min(R, r0 + CSB_ROWS);
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    if (c + 8 <= C) {
      uint4 v[8];
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // rows r0 + ty + 32 i: all loads issued before the first use
        const int64_t r = r0 + ty + 32 * i;
        const bool ok = r < r1;
        v[i] = ok ? __ldcs(reinterpret_cast<const uint4*>(in + r * ld + c)) : make_uint4(0u, 0u, 0u, 0u);
        w[i] = ok ? (roww ? __ldg(roww + r) : 1.f) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(p[k]);
          s[2 * k] = fmaf(f.x, w[i], s[2 * k]);
          s[2 * k + 1] = fmaf(f.y, w[i], s[2 * k + 1]);
        }
      }
    } else if (c < C) {
      for (int64_t r = r0 + ty; r < r1; r += 32)
        for (int k = 0; k < 8 && c + k < C; ++k)
          s[k] = fmaf(__bfloat162float(in[r * ld + c + k]), roww ? __ldg(roww + r) : 1.f, s[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ty][tx * 8 + k] = s[k];
    __syncthreads();
    if (threadIdx.x < 256) {
      const int64_t cc = (int64_t)slab * 256 + threadIdx.x;
      if (cc < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
        partial[(int64_t)chunk * C + cc] = t;
      }
    }
    __syncthreads();
  }
}
static int64_t colsum_bf16_partials(int64_t R, int64_t C) { return ((R + CSB_ROWS - 1) / CSB_ROWS) * C; }
int colsum_bf16(const bf* in, int64_t R, int64_t C, int64_t ld, float beta, float* out, float* partial,
                cudaStream_t st, const float* roww, int max_blocks) {
  const int64_t chunks = (R + CSB_ROWS - 1) / CSB_ROWS;
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)chunks);
  const int64_t items = (int64_t)grid.x * chunks;
  const int64_t blocks = (max_blocks > 0 && items > max_blocks) ? max_blocks : items;
  colsum_bf16_partial_kernel<<<(unsigned)blocks, 1024, 0, st>>>(in, R, C, ld, partial, roww, (int)grid.x, (int)chunks);
  SNT_LAUNCH_CHECK("colsum_bf16_partial_kernel");
  colsum_bf16_final_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(partial, chunks, C, beta, out);
  SNT_LAUNCH_CHECK("colsum_bf16_final_kernel");
  return SNT_OK;
}

__global__ void scale_vec_kernel(const float* __restrict__ in, int64_t n, float scale, const float* __restrict__ dloss,
                                 float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scale * (dloss ? dloss[0] : 1.f);
}

// ---------------------------------------------------------------------------------------------------------------------
static inline int64_t pad8(int64_t x) { return (x + 7) / 8 * 8; }
// rows of dlogits alive at once: (SMs/4) row tiles make dHs (BN=128) exactly one wave and the softmax-grad pass a whole
// number of waves; capped so the bf16 chunk stays around 3/4 of the 126 MB L2 (37*128 x 10000 x 2 B = 95 MB).
// Measured (round 1, V=10000, H=512): 33-37 row tiles per chunk give 499 us for the whole backward; 25 tiles 544 us,
// 19 tiles 645 us - smaller chunks would sit deeper in L2 (ncu: the consumers re-read ~1-2x the chunk from DRAM at 37
// tiles) but lose more to partial waves and per-launch overhead than they gain.
static int64_t bwd_chunk_rows(int64_t N, int64_t V) {
  int64_t tiles = tc::sm_count() / 4;
  if (tiles < 1) tiles = 1;
  while (tiles > 1 && tiles * 128 * ((V + 7) / 8 * 8) * 2 > (int64_t)96 << 20) --tiles;
  const int64_t r = tiles * 128;
  return N < r ? N : r;
}
constexpr int MAX_SPLITS = 16;

struct CeWs {
  bf* wb; float2* part; float* tl; float* nll; bf* dl; float* cpart; float* sws; float* db;
  int slabs; int64_t R, Vp; bool ok;
};
static CeWs carve(void* ws, int64_t ws_bytes, int64_t N, int64_t H, int64_t V) {
  Workspace w(ws, ws_bytes);
  CeWs r;
  r.slabs = (int)((V + CE_BN - 1) / CE_BN) * 2;
  r.R = bwd_chunk_rows(N, V);
  r.Vp = pad8(V);
  r.wb = w.take<bf>(V * H);
  r.part = w.take<float2>((int64_t)r.slabs * N);
  r.tl = w.take<float>(N);
  r.nll = w.take<float>(N);
  r.dl = w.take<bf>(r.R * r.Vp);
  r.cpart = w.take<float>(colsum_bf16_partials(r.R, V));
  r.sws = w.take<float>(MAX_SPLITS * 1024 * H);
  r.db = w.take<float>(V);
  r.ok = w.ok();
  return r;
}
int64_t vocab_ce_ws_bytes(int64_t N, int64_t H, int64_t V) {
  const int64_t slabs = ((V + CE_BN - 1) / CE_BN) * 2;
  const int64_t R = bwd_chunk_rows(N, V);
  return ws_bytes_for(V * H, 2) + ws_bytes_for(slabs * N, 8) + 2 * ws_bytes_for(N, 4) + ws_bytes_for(R * pad8(V), 2) +
         ws_bytes_for(colsum_bf16_partials(R, V), 4) + ws_bytes_for(MAX_SPLITS * 1024 * H, 4) + ws_bytes_for(V, 4);
}

static int make_sched(int64_t M, int64_t Ncols, int64_t K, tc::TileSched* ts) {
  ts->num_m = (int)((M + tc::BM - 1) / tc::BM);
  ts->num_n = (int)((Ncols + CE_BN - 1) / CE_BN);
  ts->splits = 1;
  ts->n_fastest = 0;
  ts->kblocks = (int)((K + tc::BK - 1) / tc::BK);
  ts->kblocks_per_split = ts->kblocks;
  ts->a_row0 = 0;
  ts->b_row0 = 0;
  return SNT_OK;
}

// self-resetting ticket counter of ce_finish_kernel's last-block reduction (one per device, allocated once)
static int* ce_ticket() {
  static int* tab[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!tab[dev]) {
    if (cudaMalloc(&tab[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(tab[dev], 0, sizeof(int));
  }
  return tab[dev];
}

int vocab_ce_fwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, int64_t N,
                 int64_t H, int64_t V, float* lse, float* loss, void* ws, int64_t ws_bytes, cudaStream_t st,
                 float loss_scale) {
  if (H % 8 != 0) { set_error("bf16 mode: H=%lld must be a multiple of 8", (long long)H); return SNT_EUNSUPPORTED; }
  SNT_REQUIRE(V < (1LL << 31) && N < (1LL << 31), "vocab_ce_fwd: extent too large");
  CeWs w = carve(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_fwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* hs_b = (const bf*)hs;
  SNT_CHECK(cast_bf16(w_out, w.wb, V * H, st));
  CUtensorMap ta, tb;
  SNT_CHECK(tc::make_operand_tmap(&ta, hs_b, false, N, H, H, tc::BM));
  SNT_CHECK(tc::make_operand_tmap(&tb, w.wb, false, V, H, H, CE_BN));
  tc::TileSched ts;
  make_sched(N, V, H, &ts);
  CeFwdEpi e;
  e.M = (int)N; e.V = (int)V; e.bias = b_out; e.part = w.part; e.targets = targets; e.tl = w.tl;
  SNT_CHECK((tc::launch_gemm_tc<CE_BN, false, false, CeFwdEpi>(ta, tb, ts, e, st)));
  int* ticket = ce_ticket();
  if (!ticket) { set_error("vocab_ce_fwd: could not allocate the reduction ticket"); return SNT_EINVAL; }
  // w.nll doubles as the per-block partial sums (ceil(N/64) <= N floats)
  ce_finish_kernel<<<(unsigned)((N + 63) / 64), 512, 0, st>>>(w.part, w.slabs, N, w.tl, targets, V, lse, w.nll, ticket,
                                                             loss_scale / (float)N, loss, device_flags());
  SNT_LAUNCH_CHECK("ce_finish_kernel");
  return SNT_OK;
}

int vocab_ce_bwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, const float* lse,
                 const float* dloss, float grad_scale, int64_t N, int64_t H, int64_t V, float* d_hs,
                 float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (H % 8 != 0) { set_error("bf16 mode: H=%lld must be a multiple of 8", (long long)H); return SNT_EUNSUPPORTED; }
  CeWs w = carve(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_bwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* hs_b = (const bf*)hs;
  const float scale = grad_scale / (float)N;
  SNT_CHECK(cast_bf16(w_out, w.wb, V * H, st));
  CUtensorMap tb;
  SNT_CHECK(tc::make_operand_tmap(&tb, w.wb, false, V, H, H, CE_BN));
  // tile width for dW_out: whichever of 256 / 128 wastes less of its last wave
  int dw_bn = 0;
  if (H >= 256) {
    const int64_t sms = tc::sm_count(), mt = (V + 127) / 128;
    const int64_t t256 = mt * ((H + 255) / 256), t128 = mt * ((H + 127) / 128);
    const double c256 = (double)((t256 + sms - 1) / sms) * 2.0, c128 = (double)((t128 + sms - 1) / sms) * 1.15;
    dw_bn = c128 < c256 ? 128 : 256;
  }

  SideStream* side = side_stream();
  for (int64_t r0 = 0; r0 < N; r0 += w.R) {
    const int64_t r = N - r0 < w.R ? N - r0 : w.R;
    const float acc = r0 > 0 ? 1.f : 0.f;
    if (side && r0 > 0) SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));  // previous chunk's column sums have read dl
    CUtensorMap ta;
    SNT_CHECK(tc::make_operand_tmap(&ta, hs_b + r0 * H, false, r, H, H, tc::BM));
    tc::TileSched ts;
    make_sched(r, V, H, &ts);
    // 16 epilogue warps (2 chunks each) hide the TMEM-load / staging / store latencies better than 8 warps with register
    // double-buffering: 58 -> 51 us per chunk pass (measured); SNT_CEBWD_W8 selects the 8-warp variant.
    if (!getenv("SNT_CEBWD_W8")) {
      CeBwdEpiT<16> e;
      e.M = (int)r; e.V = (int)V; e.bias = b_out; e.lse = lse + r0; e.targets = targets + r0;
      e.out = w.dl; e.ldo = w.Vp;
      SNT_CHECK((tc::launch_gemm_tc<CE_BN, false, false, CeBwdEpiT<16>>(ta, tb, ts, e, st)));
    } else {
      CeBwdEpi e;
      e.M = (int)r; e.V = (int)V; e.bias = b_out; e.lse = lse + r0; e.targets = targets + r0;
      e.out = w.dl; e.ldo = w.Vp;
      SNT_CHECK((tc::launch_gemm_tc<CE_BN, false, false, CeBwdEpi>(ta, tb, ts, e, st)));
    }
    if (side) {
      SNT_CUDA(cudaEventRecord(side->fork, st));
      SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
      SNT_CHECK(colsum_bf16(w.dl, r, V, w.Vp, acc, w.db, w.cpart, side->s));
      SNT_CUDA(cudaEventRecord(side->join, side->s));
    }
    // dHs[r,H] = dlogits[r,V] . W_out[V,H]      (B operand MN-major)
    int sp = tc::choose_splits(r, H, V, 0);
    {  // fill the machine: (row tiles x column tiles) x splits ~ SM count
      const int64_t tiles = ((r + 127) / 128) * ((H + 127) / 128);
      const int want = (int)(tc::sm_count() / (tiles > 0 ? tiles : 1));
      if (want > sp) sp = want;
    }
    if (sp > MAX_SPLITS) sp = MAX_SPLITS;
    if ((int64_t)sp * r > MAX_SPLITS * 1024) sp = (int)(MAX_SPLITS * 1024 / r);  // partials must fit the scratch
    if (sp < 1) sp = 1;
    SNT_CHECK(tc::gemm_tc(false, true, r, H, V, scale, w.dl, w.Vp, w.wb, H, 0.f, d_hs + r0 * H, nullptr, H, nullptr, sp,
                          w.sws, st, 0, dloss));
    // dW_out[V,H] += dlogits^T[V,r] . Hs[r,H]   (both operands MN-major).  Measured at V=10000, H=512 (round 1): 128- and
    // 256-wide tiles, with or without a split-K tail for the last partial wave, all land within 2% of each other - the
    // contraction re-reads the 95 MB dlogits chunk from L2 once per column tile and is bound there, not by wave shape.
    // Column tiles of one vocabulary panel are scheduled next to each other (n_fastest): a wave then touches 37 panels
    // of dlogits instead of 79+, each panel is fetched once and its other column tiles hit L2.
    SNT_CHECK(tc::gemm_tc(true, true, V, H, r, scale, w.dl, w.Vp, hs_b + r0 * H, H, acc, d_w_out, nullptr, H, nullptr, 1,
                          nullptr, st, 0, dloss, false, nullptr, /*force_bn=*/dw_bn,
                          /*n_fastest=*/1));
    if (!side) SNT_CHECK(colsum_bf16(w.dl, r, V, w.Vp, acc, w.db, w.cpart, st));
  }
  if (side) SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  scale_vec_kernel<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(w.db, V, scale, dloss, d_b_out);
  SNT_LAUNCH_CHECK("scale_vec_kernel");
  return SNT_OK;
}


// =====================================================================================================================
// Training path with STORED softmax numerators: the logits contraction runs ONCE per step.
//
// The recompute path above pays 2*N*H*V FLOP twice (forward statistics, then the same contraction again in backward
// because softmax needs the finished row statistics).  Here the forward pass stores U = exp(logit - c_row) as bf16 for a
// per-row shift c_row that is known BEFORE the pass, so no second sweep is needed:
//     softmax = U / S,  S = sum_v U  (fp32, accumulated in the same epilogue),   lse = c + log S.
// The 1/S factor is a per-ROW scalar, so it folds into the consumers exactly:
//     dHs[n,:]  = (1/S_n) * sum_v U'[n,v] W[v,:]          -> row scale in the epilogue of the contraction
//     dW[v,:]   = sum_n U'[n,v] * (Hs[n,:]/S_n)            -> the other operand is pre-scaled (hs_scaled, bf16)
//     db[v]     = sum_n U'[n,v] / S_n                      -> row-weighted column sums
// with U'[n,t_n] = U[n,t_n] - S_n (the one-hot, patched into the stored row by the finishing kernel; the patched element
// is (p-1)*S rounded to bf16 once - the same rounding the recompute path applies to p-1).
// The shift only has to be within ~70 nats of the row maximum (bf16 has fp32's exponent range): c_row = max over the
// first 256 vocabulary columns ("pilot" tile, one tiny contraction) + 20.  c <= rowmax + 20 always holds, so S >= e^-20;
// a row whose maximum exceeded the pilot maximum by more than ~88 nats would overflow: the epilogue then raises device
// flag bit 2 and the loss comes out non-finite (loud), and SNT_CE_RECOMPUTE=1 selects the recompute path.
// =====================================================================================================================
constexpr float CE_SHIFT2 = 20.f * LOG2E;  // pilot shift in the log2 domain

struct RowMaxEpi {  // pilot: per row the maximum of y = logit*log2(e) over each 128-column half of column tile 0
  static constexpr int kWarps = 8;
  static constexpr int kStages = 0;
  static constexpr int kSmemPerWarp = 0;
  int M, V;
  const float* bias;
  float* pm;  // [2][M]
  using Pre = tc::NoPre;
  __device__ __forceinline__ void prefetch(Pre&, int, int, int, int) const {}
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int, int, int ew, int lane, const Pre&,
                                       uint8_t*) const {
    const int half = ew >> 2;
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    float best = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col0 = half * 128 + c * 32;
      if (col0 >= V) break;  // warp-uniform
      uint32_t r[32];
      tc::tmem_ld32(tmem_rows + (uint32_t)col0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < V) best = fmaxf(best, (__uint_as_float(r[j]) + __ldg(bias + col0 + j)) * LOG2E);
    }
    if (row < M) pm[(int64_t)half * M + row] = best;
  }
};

// forward epilogue of the stored-numerator path: U = 2^(y - c2) -> bf16 [M, ldo], row sums of the (unrounded) U per
// 64-column quarter of the tile, and the target logit picked out of the accumulator.  16 warps x 2 chunks of 32 columns.
struct CeStoreEpi {
  static constexpr int kWarps = 16;
  static constexpr int kStages = 3;
  static constexpr int kSmemPerWarp = 32 * 64 + 256;  // swizzled transpose stage (as CeBwdEpiT) + this warp's 64 bias values
  int M, V;
  const float* bias;        // [V]
  const float* pm;          // [2][M] pilot maxima (log2 domain)
  const int64_t* targets;   // [M]
  float* tl;                // [M] target logit
  float* part;              // [4 * num_n][M] partial row sums
  bf* out;                  // [M, ldo]
  int64_t ldo;
  int* flags;

  // everything the tile needs from global memory, fetched one tile ahead (gemm_tc_kernel): the row's shift and target,
  // and two of the 64 bias values of this warp's column quarter (lane and lane + 32), shared through shared memory
  struct Pre { float c2; int tgt; float b0, b1; };
  __device__ __forceinline__ void prefetch(Pre& p, int m_blk, int n_blk, int ew, int lane) const {
    const int row = m_blk * tc::BM + (ew & 3) * 32 + lane;
    const bool ok = row < M;
    p.c2 = ok ? fmaxf(__ldg(pm + row), __ldg(pm + M + row)) + CE_SHIFT2 : 0.f;
    const int64_t t = ok ? targets[row] : -1;
    p.tgt = (t >= 0 && t < V) ? (int)t : -1;
    const int c0 = n_blk * CE_BN + (ew >> 2) * 64 + lane;
    p.b0 = c0 < V ? __ldg(bias + c0) : 0.f;
    p.b1 = c0 + 32 < V ? __ldg(bias + c0 + 32) : 0.f;
  }
  __device__ __forceinline__ void load_bias(float4 (&bv)[8], uint32_t bsa) const {  // 32 floats, broadcast reads
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 u = tc::lds128(bsa + q * 16);
      bv[q] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
    }
  }
  __device__ __forceinline__ float chunk(const uint32_t (&r)[32], const float4 (&bv)[8], int col0, int row0, int lane,
                                         float c2, int tgt, uint32_t wsa) const {
    float s = 0.f;
    const int trel = tgt - col0;
    const bool hit = tgt >= 0 && (unsigned)trel < 32u;
    if (__any_sync(0xffffffffu, hit)) {  // some row of the warp has its target among these 32 columns
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) v = (trel == j) ? __uint_as_float(r[j]) : v;
      if (hit) tl[row0 + lane] = v + __ldg(bias + tgt);
    }
    if (col0 + 32 <= V) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = bv[j >> 2];
        const float p0 = ex2(fmaf(__uint_as_float(r[j]) + b.x, LOG2E, -c2));
        const float p1 = ex2(fmaf(__uint_as_float(r[j + 1]) + b.y, LOG2E, -c2));
        const float p2 = ex2(fmaf(__uint_as_float(r[j + 2]) + b.z, LOG2E, -c2));
        const float p3 = ex2(fmaf(__uint_as_float(r[j + 3]) + b.w, LOG2E, -c2));
        s += (p0 + p1) + (p2 + p3);
        __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
        pk[j / 2] = *reinterpret_cast<uint32_t*>(&lo);
        pk[j / 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        tc::sts128(tc::stage_addr(wsa, lane, q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      __syncwarp();
      uint4 v[4];  // all four loads first, then the stores: no load waits behind a store
#pragma unroll
      for (int it = 0; it < 4; ++it) v[it] = tc::lds128(tc::stage_addr(wsa, it * 8 + (lane >> 2), lane & 3));
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + (lane >> 2), cq = lane & 3;
        if (row0 + rr < M) *reinterpret_cast<uint4*>(out + (int64_t)(row0 + rr) * ldo + col0 + cq * 8) = v[it];
      }
      __syncwarp();
    } else if (row0 + lane < M) {
      bf* orow = out + (int64_t)(row0 + lane) * ldo;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        if (col < V) {
          const float p = ex2(fmaf(__uint_as_float(r[j]) + bias[col], LOG2E, -c2));
          s += p;
          orow[col] = __float2bfloat16_rn(p);
        }
      }
    }
    return s;
  }
  __device__ __forceinline__ void tile(uint32_t tmem_rows, int m_blk, int n_blk, int, int ew, int lane, const Pre& pre,
                                       uint8_t* wsm) const {
    const int q4 = ew >> 2;  // 64-column quarter of the tile
    const int row0 = m_blk * tc::BM + (ew & 3) * 32;
    const uint32_t wsa = tc::smem_u32(wsm), bsa = wsa + 32 * 64;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bsa + lane * 4), "f"(pre.b0) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bsa + 128 + lane * 4), "f"(pre.b1) : "memory");
    __syncwarp();
    float s = 0.f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = n_blk * CE_BN + q4 * 64 + c * 32;
      if (col0 >= V) break;  // warp-uniform
      uint32_t r[32];
      float4 b8[8];
      tc::tmem_ld32(tmem_rows + (uint32_t)(q4 * 64 + c * 32), r);
      load_bias(b8, bsa + c * 128);
      tc::tmem_ld_wait();
      s += chunk(r, b8, col0, row0, lane, pre.c2, pre.tgt, wsa);
    }
    if (row0 + lane < M) {
      part[(int64_t)(n_blk * 4 + q4) * M + row0 + lane] = s;
      if (!(s < 1e35f)) atomicOr(flags, 4);  // inf / NaN / about to overflow: the pilot shift was too far off
    }
  }
};

// Finishes the stored-numerator forward.  Block = 64 rows x 8 slab groups (512 threads), as ce_finish_kernel:
//   S = sum of the row's partial sums (fixed order), lse = (c2 + log2 S) ln 2, nll = lse - target logit, mean -> loss;
//   the one-hot is patched into the stored row (U[n,t] -= S), inv_s[n] = the row scale (1/S up to 2^-9), and the block's 64 rows of Hs are written
//   scaled by 1/S (the pre-scaled operand of the dW_out contraction).
__global__ void __launch_bounds__(512)
ce_finish_u_kernel(const float* __restrict__ part, int slabs, int64_t N, const float* __restrict__ pm,
                   const float* __restrict__ tl, const int64_t* __restrict__ targets, int64_t V, bf* __restrict__ u,
                   int64_t ldu, const bf* __restrict__ hs, int H, float* __restrict__ lse, float* __restrict__ inv_s,
                   bf* __restrict__ hs_scaled, float* __restrict__ block_sum, int* __restrict__ ticket, float inv_n,
                   float* __restrict__ loss, int* flags) {
  tc::pdl_launch_dependents();
  tc::pdl_wait();  // partial sums, target logits and the stored numerators come from the preceding contraction
  __shared__ float red[8][64];
  __shared__ float s_inv[64];
  __shared__ float wred[16];
  __shared__ int s_last;
  const int tx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t row = (int64_t)blockIdx.x * 64 + tx;
  float S = 0.f;
  if (row < N)
    for (int k = g; k < slabs; k += 8) S += part[(int64_t)k * N + row];
  red[g][tx] = S;
  __syncthreads();
  float nll = 0.f;
  if (g == 0) {
    float inv = 0.f;
    if (row < N) {
      float St = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) St += red[i][tx];
      const float c2 = fmaxf(pm[row], pm[N + row]) + CE_SHIFT2;
      const float l = (c2 + log2f(St)) * LN2;
      lse[row] = l;
      inv = 1.f / St;
      const int64_t t = targets[row];
      const bool valid = t >= 0 && t < V;
      if (!valid) atomicOr(flags, 2);
      nll = l - (valid ? tl[row] : 0.f);
      if (valid) {
        // one-hot: U[n,t] <- U[n,t] - S, rounded to bf16 once.  That element carries the dominant term of every gradient
        // (-W[t,:], -Hs[n,:], -1), so the ROW SCALE is chosen to make it exact: r = (p_t - 1) / stored value, which
        // differs from 1/S by at most 2^-9 and moves the rounding error onto the small softmax terms of the row.
        bf* e = u + row * ldu + t;
        const float ut = __bfloat162float(*e);
        const bf patched = __float2bfloat16_rn(ut - St);
        *e = patched;
        const float pv = __bfloat162float(patched);
        if (pv != 0.f) inv = (ut * inv - 1.f) / pv;
      }
      inv_s[row] = inv;
    }
    s_inv[tx] = inv;
    const float w = warp_sum(nll);
    if ((tx & 31) == 0) wred[tx >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    block_sum[blockIdx.x] = wred[0] + wred[1];
    __threadfence();
    s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  }
  // Hs rows of this block, scaled by 1/S (H % 8 == 0: 16-byte pieces)
  const int h8 = H >> 3;
  for (int i = threadIdx.x; i < 64 * h8; i += 512) {
    const int r = i / h8, c = i - r * h8;
    const int64_t grow = (int64_t)blockIdx.x * 64 + r;
    if (grow >= N) break;
    const float w = s_inv[r];
    uint4 v = *reinterpret_cast<const uint4*>(hs + grow * H + c * 8);
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(p2[k]);
      p2[k] = __floats2bfloat162_rn(f.x * w, f.y * w);
    }
    *reinterpret_cast<uint4*>(hs_scaled + grow * H + c * 8) = v;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 512) t += __ldcg(block_sum + i);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) tot += wred[i];
      loss[0] = tot * inv_n;
      *ticket = 0;  // self-resetting
    }
  }
}

struct CeTrainWs {
  float* pm; float* part; float* tl; float* bsum; float* cpart; float* sws; float* db; int slabs; int64_t Vp; bool ok;
};
// scratch for the K-split tails of the two backward contractions: 16 splits x (less than one wave of row tiles) x H
static int64_t train_sws_elems(int64_t N, int64_t H, int64_t V) {
  const int64_t num_n = (H + 255) / 256 > 0 ? (H + 255) / 256 : 1;
  const int64_t tail_rows = ((int64_t)tc::sm_count() / num_n + 2) * 128;
  const int64_t cap = N > V ? N : V;
  return 16 * (tail_rows < cap ? tail_rows : cap) * H;
}
static CeTrainWs carve_train(void* ws, int64_t ws_bytes, int64_t N, int64_t H, int64_t V) {
  Workspace w(ws, ws_bytes);
  CeTrainWs r;
  r.slabs = (int)((V + CE_BN - 1) / CE_BN) * 4;
  r.Vp = pad8(V);
  r.pm = w.take<float>(2 * N);
  r.part = w.take<float>((int64_t)r.slabs * N);
  r.tl = w.take<float>(N);
  r.bsum = w.take<float>(N);
  r.cpart = w.take<float>(colsum_bf16_partials(N, V));
  r.sws = w.take<float>(train_sws_elems(N, H, V));
  r.db = w.take<float>(V);
  r.ok = w.ok();
  return r;
}
int64_t vocab_ce_train_ws_bytes(int64_t N, int64_t H, int64_t V) {
  const int64_t slabs = ((V + CE_BN - 1) / CE_BN) * 4;
  return ws_bytes_for(2 * N, 4) + ws_bytes_for(slabs * N, 4) + 2 * ws_bytes_for(N, 4) +
         ws_bytes_for(colsum_bf16_partials(N, V), 4) + ws_bytes_for(train_sws_elems(N, H, V), 4) + ws_bytes_for(V, 4);
}

int vocab_ce_train_fwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, int64_t N,
                       int64_t H, int64_t V, float* lse, float* loss, void* u, float* inv_s, void* hs_scaled,
                       void* w_bf16, void* ws, int64_t ws_bytes, cudaStream_t st, float loss_scale, bool w_prepared) {
  if (H % 8 != 0) { set_error("bf16 mode: H=%lld must be a multiple of 8", (long long)H); return SNT_EUNSUPPORTED; }
  SNT_REQUIRE(V < (1LL << 31) && N < (1LL << 31), "vocab_ce_train_fwd: extent too large");
  CeTrainWs w = carve_train(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_train_fwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* hs_b = (const bf*)hs;
  bf* wb = (bf*)w_bf16;
  if (!w_prepared) SNT_CHECK(cast_bf16(w_out, wb, V * H, st));
  CUtensorMap ta, tb;
  SNT_CHECK(tc::make_operand_tmap(&ta, hs_b, false, N, H, H, tc::BM));
  SNT_CHECK(tc::make_operand_tmap(&tb, wb, false, V, H, H, CE_BN));
  tc::TileSched ts;
  make_sched(N, V, H, &ts);
  {  // pilot: column tile 0 only
    tc::TileSched tp = ts;
    tp.num_n = 1;
    RowMaxEpi e;
    e.M = (int)N; e.V = (int)V; e.bias = b_out; e.pm = w.pm;
    SNT_CHECK((tc::launch_gemm_tc<CE_BN, false, false, RowMaxEpi>(ta, tb, tp, e, st)));
  }
  CeStoreEpi e;
  e.M = (int)N; e.V = (int)V; e.bias = b_out; e.pm = w.pm; e.targets = targets; e.tl = w.tl; e.part = w.part;
  e.out = (bf*)u; e.ldo = w.Vp; e.flags = device_flags();
  SNT_CHECK((tc::launch_gemm_tc<CE_BN, false, false, CeStoreEpi>(ta, tb, ts, e, st)));
  int* ticket = ce_ticket();
  if (!ticket) { set_error("vocab_ce_train_fwd: could not allocate the reduction ticket"); return SNT_EINVAL; }
  SNT_CUDA(tc::launch_chained(ce_finish_u_kernel, dim3((unsigned)((N + 63) / 64)), dim3(512), 0, st, (const float*)w.part,
                              w.slabs, N, (const float*)w.pm, (const float*)w.tl, targets, V, (bf*)u, w.Vp, hs_b, (int)H,
                              lse, inv_s, (bf*)hs_scaled, w.bsum, ticket, loss_scale / (float)N, loss, device_flags()));
  SNT_LAUNCH_CHECK("ce_finish_u_kernel");
  return SNT_OK;
}

// d_b_out of the stored-numerator path on its own (see vocab_ce_train_bwd(defer_bias)): column sums of U' weighted by the
// row scales, from a grid of at most `max_blocks` blocks (0: the wide grid).  part: vocab_ce_train_bias_part_elems(N, V)
// floats, db: V floats - caller-owned, so the call may overlap later stages that reuse the stage workspace.
int64_t vocab_ce_train_bias_part_elems(int64_t N, int64_t V) { return colsum_bf16_partials(N, V); }
int vocab_ce_train_bias(const void* u, const float* inv_s, const float* dloss, float grad_scale, int64_t N, int64_t V,
                        float* d_b_out, float* part, float* db, cudaStream_t st, int max_blocks) {
  SNT_REQUIRE(u && inv_s && d_b_out && part && db, "vocab_ce_train_bias: NULL argument");
  SNT_CHECK(colsum_bf16((const bf*)u, N, V, pad8(V), 0.f, db, part, st, inv_s, max_blocks));
  scale_vec_kernel<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(db, V, grad_scale / (float)N, dloss, d_b_out);
  SNT_LAUNCH_CHECK("scale_vec_kernel");
  return SNT_OK;
}

// The dW_out half of vocab_ce_train_bwd(defer_dw = true) on any stream, from a grid of at most `max_ctas` CTAs (0: all
// SMs), with caller-owned split-K scratch of vocab_ce_train_sws_elems(N, H, V) floats: the call may overlap later stages
// that reuse the stage workspace.
int64_t vocab_ce_train_sws_elems(int64_t N, int64_t H, int64_t V) { return train_sws_elems(N, H, V); }
int vocab_ce_train_dw(const void* u, const void* hs_scaled, const float* dloss, float grad_scale, int64_t N, int64_t H,
                      int64_t V, float* d_w_out, float* sws, int64_t sws_elems, cudaStream_t st, int max_ctas) {
  SNT_REQUIRE(u && hs_scaled && d_w_out, "vocab_ce_train_dw: NULL argument");
  const int bn = H >= 256 ? 256 : 128;
  const int old_cap = tc::set_grid_cap(max_ctas);
  const int rc = tc::gemm_tc_balanced(true, true, V, H, N, grad_scale / (float)N, (const bf*)u, pad8(V),
                                      (const bf*)hs_scaled, H, d_w_out, H, sws, sws_elems, st, dloss, bn);
  tc::set_grid_cap(old_cap);
  return rc;
}

int vocab_ce_train_bwd(const void* u, const float* inv_s, const void* hs_scaled, const void* w_bf16,
                       const float* dloss, float grad_scale, int64_t N, int64_t H, int64_t V, float* d_hs,
                       float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st, bool defer_bias,
                       bool defer_dw) {
  if (H % 8 != 0) { set_error("bf16 mode: H=%lld must be a multiple of 8", (long long)H); return SNT_EUNSUPPORTED; }
  CeTrainWs w = carve_train(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_train_bwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* ub = (const bf*)u;
  const bf* wb = (const bf*)w_bf16;
  const float scale = grad_scale / (float)N;
  // d_b_out = scale * sum_n U'[n,:] / S_n: bandwidth-bound, on the side stream next to the two contractions
  // (defer_bias: the caller runs vocab_ce_train_bias later, next to a stage that leaves HBM idle)
  SideStream* side = defer_bias ? nullptr : side_stream();
  if (side) {
    SNT_CUDA(cudaEventRecord(side->fork, st));
    SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    SNT_CHECK(colsum_bf16(ub, N, V, w.Vp, 0.f, w.db, w.cpart, side->s, inv_s));
    scale_vec_kernel<<<(unsigned)((V + 255) / 256), 256, 0, side->s>>>(w.db, V, scale, dloss, d_b_out);
    SNT_LAUNCH_CHECK("scale_vec_kernel");
    SNT_CUDA(cudaEventRecord(side->join, side->s));
  }
  const int bn = H >= 256 ? 256 : 128;
  const int64_t sws_elems = train_sws_elems(N, H, V);
  PdlSuppress plain_order(side != nullptr);  // the column sums above run beside the two contractions: leave them gaps
  // dHs[N,H] = diag(scale * r) . U'[N,V] . W_out[V,H]      (B operand MN-major); the column tiles of one 128-row panel
  // run next to each other so that the panel of U is fetched from HBM once
  SNT_CHECK(tc::gemm_tc_balanced(false, true, N, H, V, scale, ub, w.Vp, wb, H, d_hs, H, w.sws, sws_elems, st, dloss, bn,
                                 inv_s));
  // dW_out[V,H] = scale * U'^T[V,N] . (r * Hs)[N,H]          (both operands MN-major)
  // (defer_dw: the caller runs vocab_ce_train_dw beside the persistent BPTT recurrence, which needs dHs only)
  if (!defer_dw)
    SNT_CHECK(tc::gemm_tc_balanced(true, true, V, H, N, scale, ub, w.Vp, (const bf*)hs_scaled, H, d_w_out, H, w.sws,
                                   sws_elems, st, dloss, bn));
  if (side) {
    SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  } else if (!defer_bias) {
    SNT_CHECK(colsum_bf16(ub, N, V, w.Vp, 0.f, w.db, w.cpart, st, inv_s));
    scale_vec_kernel<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(w.db, V, scale, dloss, d_b_out);
    SNT_LAUNCH_CHECK("scale_vec_kernel");
  }
  return SNT_OK;
}

}  // namespace bf16
}  // namespace snt
