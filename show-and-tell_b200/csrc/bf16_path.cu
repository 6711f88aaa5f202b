// SNT_PREC_BF16 pipelines: every contraction runs on tcgen05 tensor cores (gemm_tc.cuh) with bf16 operands and
// fp32 accumulation in TMEM; master weights stay fp32 and are rounded to bf16 into the workspace per call.
// Contiguous operand dimensions (In, H, K of the head) must be multiples of 8 (TMA 16-byte pitch rule).
#include "bf16.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace snt {
namespace bf16 {

typedef __nv_bfloat16 bf;

static inline int64_t pad8(int64_t x) { return (x + 7) / 8 * 8; }
constexpr int64_t CE_CHUNK_ROWS = 1024;
constexpr int MAX_SPLITS = 16;

#define SNT_REQUIRE_ALIGNED8(v, what)                                                                  \
  do {                                                                                                 \
    if ((v) % 8 != 0) {                                                                                \
      set_error("bf16 mode: %s=%lld must be a multiple of 8 (TMA row pitch is 16 bytes)", what,        \
                (long long)(v));                                                                       \
      return SNT_EUNSUPPORTED;                                                                         \
    }                                                                                                  \
  } while (0)

// src[rows, cols] fp32 (ld_src) -> dst[rows, ld_dst] bf16, columns >= cols zero
__global__ void __launch_bounds__(256)
cast2d_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src, bf* __restrict__ dst,
              int64_t ld_dst) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * ld_dst) return;
  const int64_t r = i / ld_dst, c = i % ld_dst;
  dst[i] = __float2bfloat16_rn(c < cols ? src[r * ld_src + c] : 0.f);
}
static int cast2d(const float* src, int64_t rows, int64_t cols, int64_t ld_src, bf* dst, int64_t ld_dst,
                  cudaStream_t st) {
  if (rows <= 0) return SNT_OK;
  if (cols == ld_src && cols == ld_dst) return cast_bf16(src, dst, rows * cols, st);
  cast2d_kernel<<<(unsigned)((rows * ld_dst + 255) / 256), 256, 0, st>>>(src, rows, cols, ld_src, dst, ld_dst);
  SNT_LAUNCH_CHECK("cast2d_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// head helpers
// ---------------------------------------------------------------------------------------------------------
int64_t head_extra_ws_bytes(int64_t B, int64_t K, int64_t E) {
  return ws_bytes_for(B * K, 2) + ws_bytes_for(E * K, 2) + ws_bytes_for(B * pad8(E), 2) +
         ws_bytes_for(MAX_SPLITS * E * K, 4);
}

int linear_nt(const float* a, const float* w, const float* bias, int64_t M, int64_t N, int64_t K, float* y,
              void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(K, "K");
  Workspace wk(ws, ws_bytes);
  bf* ab = wk.take<bf>(M * K);
  bf* wb = wk.take<bf>(N * K);
  if (!wk.ok()) { set_error("bf16 linear: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(a, ab, M * K, st));
  SNT_CHECK(cast_bf16(w, wb, N * K, st));
  return tc::gemm_tc(false, false, M, N, K, 1.f, ab, K, wb, K, 0.f, y, nullptr, N, bias, 1, nullptr, st);
}

int wgrad_tn(const float* dy, const float* a, int64_t M, int64_t N, int64_t K, float* dw, void* ws,
             int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(K, "K");
  const int64_t Np = pad8(N);
  Workspace wk(ws, ws_bytes);
  bf* ab = wk.take<bf>(M * K);
  wk.take<bf>(N * K);
  bf* dyb = wk.take<bf>(M * Np);
  float* sws = wk.take<float>(MAX_SPLITS * N * K);
  if (!wk.ok()) { set_error("bf16 wgrad: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(a, ab, M * K, st));
  SNT_CHECK(cast2d(dy, M, N, N, dyb, Np, st));
  int splits = tc::choose_splits(N, K, M, 0);
  if (splits > MAX_SPLITS) splits = MAX_SPLITS;
  // dw[N,K] = dy^T . a : A = dy as [K'=M rows, M'=N] (MN-major), B = a as [K'=M rows, N'=K] (MN-major)
  return tc::gemm_tc(true, true, N, K, M, 1.f, dyb, Np, ab, K, 0.f, dw, nullptr, K, nullptr, splits, sws, st);
}

// ---------------------------------------------------------------------------------------------------------
// LSTM layer
// ---------------------------------------------------------------------------------------------------------
struct LstmWs {
  float* bsum; bf* w_ih; bf* w_hh; float* dh_rec; float* dc_state; float* part; bf* dg; float* sws;
  bool ok;
};
static LstmWs carve_lstm(void* ws, int64_t ws_bytes, int64_t N, int64_t B, int64_t In, int64_t H) {
  Workspace w(ws, ws_bytes);
  LstmWs r;
  r.bsum = w.take<float>(4 * H);
  r.w_ih = w.take<bf>(4 * H * In);
  r.w_hh = w.take<bf>(4 * H * H);
  r.dh_rec = w.take<float>(B * H);
  r.dc_state = w.take<float>(B * H);
  r.part = w.take<float>(colsum_partial_count(N, 4 * H));
  r.dg = w.take<bf>(N * 4 * H);
  r.sws = w.take<float>(MAX_SPLITS * 4 * H * (In > H ? In : H));
  r.ok = w.ok();
  return r;
}
int64_t lstm_ws_bytes(int64_t N, int64_t B, int64_t In, int64_t H) {
  return ws_bytes_for(4 * H, 4) + ws_bytes_for(4 * H * In, 2) + ws_bytes_for(4 * H * H, 2) +
         2 * ws_bytes_for(B * H, 4) + ws_bytes_for(colsum_partial_count(N, 4 * H), 4) +
         ws_bytes_for(N * 4 * H, 2) + ws_bytes_for(MAX_SPLITS * 4 * H * (In > H ? In : H), 4);
}

int lstm_fwd(const PackInfo& pk, const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
             const float* b_ih, const float* b_hh, float* gates, float* cs, void* hs, void* hprev, void* ws,
             int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(In, "In");
  SNT_REQUIRE_ALIGNED8(H, "H");
  const int T = pk.T;
  const int64_t N = pk.off[T], B = pk.off[1];
  LstmWs w = carve_lstm(ws, ws_bytes, N, B, In, H);
  if (!w.ok) { set_error("bf16 lstm_fwd: workspace too small"); return SNT_EWORKSPACE; }
  bf* hs_b = (bf*)hs;
  bf* hp_b = (bf*)hprev;
  SNT_CHECK(add_vec(b_ih, b_hh, w.bsum, 4 * H, st));
  SNT_CHECK(cast_bf16(w_ih, w.w_ih, 4 * H * In, st));
  SNT_CHECK(cast_bf16(w_hh, w.w_hh, 4 * H * H, st));
  // input projection of all timesteps as one tensor-core contraction: gates = x . W_ih^T + (b_ih + b_hh)
  SNT_CHECK(tc::gemm_tc(false, false, N, 4 * H, In, 1.f, (const bf*)x, In, w.w_ih, In, 0.f, gates, nullptr, 4 * H,
                        w.bsum, 1, nullptr, st));
  SNT_CUDA(cudaMemsetAsync(hp_b, 0, sizeof(bf) * (size_t)B * H, st));
  for (int t = 0; t < T; ++t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    float* g_t = gates + (int64_t)pk.off[t] * 4 * H;
    if (t > 0)
      SNT_CHECK(tc::gemm_tc(false, false, bs, 4 * H, H, 1.f, hp_b + (int64_t)pk.off[t] * H, H, w.w_hh, H, 1.f, g_t,
                            nullptr, 4 * H, nullptr, 1, nullptr, st));
    const float* c_prev = t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr;
    SNT_CHECK(lstm_point_fwd<bf>(g_t, c_prev, cs + (int64_t)pk.off[t] * H, hs_b + (int64_t)pk.off[t] * H,
                                 bs_next > 0 ? hp_b + (int64_t)pk.off[t + 1] * H : nullptr, bs, bs_next, H, st));
  }
  return SNT_OK;
}

int lstm_bwd(const PackInfo& pk, const float* d_hs, float* gates, const float* cs, const void* hprev,
             const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh, float* d_w_ih,
             float* d_w_hh, float* d_bias, float* dx, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(In, "In");
  SNT_REQUIRE_ALIGNED8(H, "H");
  const int T = pk.T;
  const int64_t N = pk.off[T], B = pk.off[1];
  LstmWs w = carve_lstm(ws, ws_bytes, N, B, In, H);
  if (!w.ok) { set_error("bf16 lstm_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(w_ih, w.w_ih, 4 * H * In, st));
  SNT_CHECK(cast_bf16(w_hh, w.w_hh, 4 * H * H, st));
  for (int t = T - 1; t >= 0; --t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    float* g_t = gates + (int64_t)pk.off[t] * 4 * H;
    bf* dg_t = w.dg + (int64_t)pk.off[t] * 4 * H;
    const float* c_prev = t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr;
    SNT_CHECK(lstm_point_bwd(g_t, cs + (int64_t)pk.off[t] * H, c_prev, d_hs + (int64_t)pk.off[t] * H, w.dh_rec,
                             w.dc_state, bs, bs_next, H, st));
    SNT_CHECK(cast_bf16(g_t, dg_t, (int64_t)bs * 4 * H, st));
    if (t > 0)  // dh_{t-1} = dG_t . W_hh : B operand is W_hh as [K=4H, N=H] (MN-major)
      SNT_CHECK(tc::gemm_tc(false, true, bs, H, 4 * H, 1.f, dg_t, 4 * H, w.w_hh, H, 0.f, w.dh_rec, nullptr, H,
                            nullptr, 1, nullptr, st));
  }
  int s1 = tc::choose_splits(4 * H, In, N, 0), s2 = tc::choose_splits(4 * H, H, N, 0);
  if (s1 > MAX_SPLITS) s1 = MAX_SPLITS;
  if (s2 > MAX_SPLITS) s2 = MAX_SPLITS;
  SNT_CHECK(tc::gemm_tc(true, true, 4 * H, In, N, 1.f, w.dg, 4 * H, (const bf*)x, In, 0.f, d_w_ih, nullptr, In,
                        nullptr, s1, w.sws, st));
  SNT_CHECK(tc::gemm_tc(true, true, 4 * H, H, N, 1.f, w.dg, 4 * H, (const bf*)hprev, H, 0.f, d_w_hh, nullptr, H,
                        nullptr, s2, w.sws, st));
  SNT_CHECK(colsum(gates, N, 4 * H, 4 * H, 0.f, d_bias, w.part, st));
  if (dx)
    SNT_CHECK(tc::gemm_tc(false, true, N, In, 4 * H, 1.f, w.dg, 4 * H, w.w_ih, In, 0.f, dx, nullptr, In, nullptr, 1,
                          nullptr, st));
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// materialising vocab Linear
// ---------------------------------------------------------------------------------------------------------
int64_t linear_ws_bytes(int64_t N, int64_t H, int64_t V) {
  return ws_bytes_for(V * H, 2) + ws_bytes_for(N * pad8(V), 2) + ws_bytes_for(colsum_partial_count(N, V), 4);
}
int linear_fwd(const void* hs, const float* w_out, const float* b_out, int64_t N, int64_t H, int64_t V,
               float* logits, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(H, "H");
  Workspace w(ws, ws_bytes);
  bf* wb = w.take<bf>(V * H);
  if (!w.ok()) { set_error("bf16 linear_fwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(w_out, wb, V * H, st));
  return tc::gemm_tc(false, false, N, V, H, 1.f, (const bf*)hs, H, wb, H, 0.f, logits, nullptr, V, b_out, 1, nullptr,
                     st);
}
int linear_bwd(const float* dlogits, const void* hs, const float* w_out, int64_t N, int64_t H, int64_t V,
               float* d_hs, float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(H, "H");
  const int64_t Vp = pad8(V);
  Workspace w(ws, ws_bytes);
  bf* wb = w.take<bf>(V * H);
  bf* dlb = w.take<bf>(N * Vp);
  float* part = w.take<float>(colsum_partial_count(N, V));
  if (!w.ok()) { set_error("bf16 linear_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(w_out, wb, V * H, st));
  SNT_CHECK(cast2d(dlogits, N, V, V, dlb, Vp, st));
  SNT_CHECK(tc::gemm_tc(false, true, N, H, V, 1.f, dlb, Vp, wb, H, 0.f, d_hs, nullptr, H, nullptr, 1, nullptr, st));
  SNT_CHECK(tc::gemm_tc(true, true, V, H, N, 1.f, dlb, Vp, (const bf*)hs, H, 0.f, d_w_out, nullptr, H, nullptr, 1,
                        nullptr, st));
  return colsum(dlogits, N, V, V, 0.f, d_b_out, part, st);
}

// ---------------------------------------------------------------------------------------------------------
// vocab Linear + log-softmax + CE: chunks of CE_CHUNK_ROWS rows of logits live in the workspace (L2-sized)
// ---------------------------------------------------------------------------------------------------------
struct CeWs { bf* wb; float* chunk; bf* chunk_b; float* nll; float* part; bool ok; int64_t R, Vp; };
static CeWs carve_ce(void* ws, int64_t ws_bytes, int64_t N, int64_t H, int64_t V) {
  Workspace w(ws, ws_bytes);
  CeWs r;
  r.R = N < CE_CHUNK_ROWS ? N : CE_CHUNK_ROWS;
  r.Vp = pad8(V);
  r.wb = w.take<bf>(V * H);
  r.chunk = w.take<float>(r.R * V);
  r.chunk_b = w.take<bf>(r.R * r.Vp);
  r.nll = w.take<float>(N);
  r.part = w.take<float>(colsum_partial_count(r.R, V));
  r.ok = w.ok();
  return r;
}
int64_t vocab_ce_ws_bytes(int64_t N, int64_t H, int64_t V) {
  const int64_t R = N < CE_CHUNK_ROWS ? N : CE_CHUNK_ROWS;
  return ws_bytes_for(V * H, 2) + ws_bytes_for(R * V, 4) + ws_bytes_for(R * pad8(V), 2) + ws_bytes_for(N, 4) +
         ws_bytes_for(colsum_partial_count(R, V), 4);
}
int vocab_ce_fwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, int64_t N,
                 int64_t H, int64_t V, float* lse, float* loss, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(H, "H");
  CeWs w = carve_ce(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_fwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* hs_b = (const bf*)hs;
  SNT_CHECK(cast_bf16(w_out, w.wb, V * H, st));
  for (int64_t r0 = 0; r0 < N; r0 += w.R) {
    const int64_t r = N - r0 < w.R ? N - r0 : w.R;
    SNT_CHECK(tc::gemm_tc(false, false, r, V, H, 1.f, hs_b + r0 * H, H, w.wb, H, 0.f, w.chunk, nullptr, V, b_out, 1,
                          nullptr, st));
    SNT_CHECK(ce_rows_fwd(w.chunk, r, V, V, targets + r0, lse + r0, w.nll + r0, st));
  }
  return reduce_sum(w.nll, N, 1.0f / (float)N, loss, st);
}
int vocab_ce_bwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, const float* lse,
                 const float* dloss, float grad_scale, int64_t N, int64_t H, int64_t V, float* d_hs,
                 float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(H, "H");
  CeWs w = carve_ce(ws, ws_bytes, N, H, V);
  if (!w.ok) { set_error("bf16 vocab_ce_bwd: workspace too small"); return SNT_EWORKSPACE; }
  const bf* hs_b = (const bf*)hs;
  const float scale = grad_scale / (float)N;
  SNT_CHECK(cast_bf16(w_out, w.wb, V * H, st));
  for (int64_t r0 = 0; r0 < N; r0 += w.R) {
    const int64_t r = N - r0 < w.R ? N - r0 : w.R;
    const float acc = r0 > 0 ? 1.f : 0.f;
    SNT_CHECK(tc::gemm_tc(false, false, r, V, H, 1.f, hs_b + r0 * H, H, w.wb, H, 0.f, w.chunk, nullptr, V, b_out, 1,
                          nullptr, st));
    SNT_CHECK(ce_rows_bwd(w.chunk, r, V, V, targets + r0, lse + r0, dloss, scale, st));
    SNT_CHECK(cast2d(w.chunk, r, V, V, w.chunk_b, w.Vp, st));
    SNT_CHECK(tc::gemm_tc(false, true, r, H, V, 1.f, w.chunk_b, w.Vp, w.wb, H, 0.f, d_hs + r0 * H, nullptr, H, nullptr,
                          1, nullptr, st));
    SNT_CHECK(tc::gemm_tc(true, true, V, H, r, 1.f, w.chunk_b, w.Vp, hs_b + r0 * H, H, acc, d_w_out, nullptr, H,
                          nullptr, 1, nullptr, st));
    SNT_CHECK(colsum(w.chunk, r, V, V, acc, d_b_out, w.part, st));
  }
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// greedy decode
// ---------------------------------------------------------------------------------------------------------
int64_t greedy_ws_bytes(int64_t B, int64_t E, int64_t H, int64_t V, int L) {
  int64_t b = ws_bytes_for(B * E, 4) + ws_bytes_for(B * E, 2) + ws_bytes_for(V * H, 2) +
              ws_bytes_for(B * 4 * H, 4) + ws_bytes_for(B * V, 4);
  for (int k = 0; k < L; ++k) {
    const int64_t in = k == 0 ? E : H;
    b += ws_bytes_for(B * H, 2) + ws_bytes_for(B * H, 4) + ws_bytes_for(4 * H, 4) + ws_bytes_for(4 * H * in, 2) +
         ws_bytes_for(4 * H * H, 2);
  }
  return b;
}
int greedy_decode(const float* features, const float* w_emb, int L, const float* const* w_ih,
                  const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                  const float* w_out, const float* b_out, const float* h0, const float* c0, int64_t B, int64_t E,
                  int64_t H, int64_t V, int steps, int64_t* ids, void* ws, int64_t ws_bytes, cudaStream_t st) {
  SNT_REQUIRE_ALIGNED8(E, "E");
  SNT_REQUIRE_ALIGNED8(H, "H");
  Workspace w(ws, ws_bytes);
  float* x_f = w.take<float>(B * E);
  bf* x_b = w.take<bf>(B * E);
  bf* wout_b = w.take<bf>(V * H);
  float* gates = w.take<float>(B * 4 * H);
  float* logits = w.take<float>(B * V);
  bf *h[SNT_MAX_LAYERS], *wih[SNT_MAX_LAYERS], *whh[SNT_MAX_LAYERS];
  float *c[SNT_MAX_LAYERS], *bsum[SNT_MAX_LAYERS];
  for (int k = 0; k < L; ++k) {
    const int64_t in = k == 0 ? E : H;
    h[k] = w.take<bf>(B * H);
    c[k] = w.take<float>(B * H);
    bsum[k] = w.take<float>(4 * H);
    wih[k] = w.take<bf>(4 * H * in);
    whh[k] = w.take<bf>(4 * H * H);
  }
  if (!w.ok()) { set_error("bf16 greedy_decode: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(cast_bf16(w_out, wout_b, V * H, st));
  SNT_CHECK(cast_bf16(features, x_b, B * E, st));
  for (int k = 0; k < L; ++k) {
    const int64_t in = k == 0 ? E : H;
    SNT_CHECK(add_vec(b_ih[k], b_hh[k], bsum[k], 4 * H, st));
    SNT_CHECK(cast_bf16(w_ih[k], wih[k], 4 * H * in, st));
    SNT_CHECK(cast_bf16(w_hh[k], whh[k], 4 * H * H, st));
    if (h0) {
      SNT_CHECK(cast_bf16(h0 + (int64_t)k * B * H, h[k], B * H, st));
      SNT_CUDA(cudaMemcpyAsync(c[k], c0 + (int64_t)k * B * H, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
    } else {
      SNT_CUDA(cudaMemsetAsync(h[k], 0, sizeof(bf) * B * H, st));
      SNT_CUDA(cudaMemsetAsync(c[k], 0, sizeof(float) * B * H, st));
    }
  }
  for (int s = 0; s < steps; ++s) {
    const bf* inp = x_b;
    int64_t in = E;
    for (int k = 0; k < L; ++k) {
      SNT_CHECK(tc::gemm_tc(false, false, B, 4 * H, in, 1.f, inp, in, wih[k], in, 0.f, gates, nullptr, 4 * H, bsum[k],
                            1, nullptr, st));
      SNT_CHECK(tc::gemm_tc(false, false, B, 4 * H, H, 1.f, h[k], H, whh[k], H, 1.f, gates, nullptr, 4 * H, nullptr,
                            1, nullptr, st));
      SNT_CHECK(lstm_point_fwd<bf>(gates, c[k], c[k], h[k], nullptr, (int)B, 0, H, st));
      inp = h[k];
      in = H;
    }
    SNT_CHECK(tc::gemm_tc(false, false, B, V, H, 1.f, inp, H, wout_b, H, 0.f, logits, nullptr, V, b_out, 1, nullptr,
                          st));
    SNT_CHECK(argmax_gather(logits, B, V, V, w_emb, E, ids + s, steps, x_f, st));
    SNT_CHECK(cast_bf16(x_f, x_b, B * E, st));
  }
  return SNT_OK;
}

}  // namespace bf16
}  // namespace snt
