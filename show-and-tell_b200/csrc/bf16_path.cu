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
  wk.take<bf>(M * pad8(N));
  float* sws = wk.take<float>(MAX_SPLITS * N * K);  // shared with wgrad_tn's split-K scratch
  SNT_CHECK(cast_bf16(a, ab, M * K, st, /*chain=*/true));   // casts -> contraction -> reduction -> BatchNorm: one chain
  SNT_CHECK(cast_bf16(w, wb, N * K, st, /*chain=*/true));
  // few output tiles (8 x 2 at B=1024, E=256) but a long contraction: split K over the idle SMs
  int splits = sws ? tc::choose_splits(M, N, K, 0) : 1;
  while (splits > 1 && (int64_t)splits * M > (int64_t)MAX_SPLITS * K) --splits;
  return tc::gemm_tc(false, false, M, N, K, 1.f, ab, K, wb, K, 0.f, y, nullptr, N, bias, splits, sws, st);
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

}  // namespace bf16
}  // namespace snt
