// Stage functions of the C ABI (include/snt_b200.h): argument validation, workspace carving and the
// SNT_PREC_FP32 ("fp32-faithful") pipelines built from the SIMT fp32 contraction (gemm_f32.cu) and the
// pointwise / HBM-bound kernels (kernels.cu).  SNT_PREC_BF16 dispatches to the tcgen05 pipelines (bf16.cuh).
#include "kernels.cuh"
#include "x3.cuh"
#include "bf16.cuh"

using namespace snt;

namespace {

constexpr int64_t CE_CHUNK_ROWS = 1024;  // rows of logits alive at once in the fused CE (fp32 mode)

inline bool valid_prec(int prec) { return prec == SNT_PREC_FP32 || prec == SNT_PREC_BF16; }

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// small exports
// ------------------------------------------------------------------------------------------------------------
extern "C" int snt_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  SNT_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "snt_cast_bf16: bad arguments");
  return cast_bf16(src, (__nv_bfloat16*)dst, n, (cudaStream_t)stream);
}

extern "C" int snt_clamp_adam(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                              double beta2, double eps, float grad_clip, float grad_scale, int64_t step,
                              void* stream) {
  SNT_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v)), "snt_clamp_adam: bad arguments");
  return clamp_adam(p, g, m, v, n, lr, beta1, beta2, eps, grad_clip, grad_scale, step, (cudaStream_t)stream);
}

extern "C" int snt_clamp_adam_multi(int count, float* const* p, const float* const* g, float* const* m,
                                    float* const* v, const int64_t* n, double lr, double beta1, double beta2,
                                    double eps, float grad_clip, float grad_scale, int64_t step, void* stream) {
  SNT_REQUIRE(count >= 0 && (count == 0 || (p && g && m && v && n)), "snt_clamp_adam_multi: bad arguments");
  return clamp_adam_multi(count, p, g, m, v, n, lr, beta1, beta2, eps, grad_clip, grad_scale, step,
                          (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// a1/a2: encoder head
// ------------------------------------------------------------------------------------------------------------
extern "C" int64_t snt_head_workspace_bytes(int prec, int64_t B, int64_t K, int64_t E) {
  if (!valid_prec(prec) || B < 1 || K < 1 || E < 1) return -1;
  int64_t b = ws_bytes_for(B * E, 4) + ws_bytes_for(colsum_partial_count(B, E), 4);
  if (prec == SNT_PREC_BF16) b += bf16::head_extra_ws_bytes(B, K, E);
  return b;
}

extern "C" int snt_head_fwd(int prec, const float* pooled, const float* w_fc, const float* b_fc,
                            const float* gamma, const float* beta, float* running_mean, float* running_var,
                            int training, float momentum, float eps, int64_t B, int64_t K, int64_t E,
                            float* features, float* yhat, float* rstd, void* ws, int64_t ws_bytes, void* stream) {
  return snt::head_fwd(prec, pooled, w_fc, b_fc, gamma, beta, running_mean, running_var, training, momentum, eps, B, K, E,
                       features, yhat, rstd, ws, ws_bytes, (cudaStream_t)stream, nullptr);
}
int snt::head_fwd(int prec, const float* pooled, const float* w_fc, const float* b_fc, const float* gamma,
                  const float* beta, float* running_mean, float* running_var, int training, float momentum, float eps,
                  int64_t B, int64_t K, int64_t E, float* features, float* yhat, float* rstd, void* ws, int64_t ws_bytes,
                  cudaStream_t stream, __nv_bfloat16* features_bf16) {
  SNT_REQUIRE(valid_prec(prec), "snt_head_fwd: bad prec %d", prec);
  SNT_REQUIRE(B >= 1 && K >= 1 && E >= 1, "snt_head_fwd: bad sizes");
  SNT_REQUIRE(pooled && w_fc && b_fc && gamma && beta && running_mean && running_var && features && yhat && rstd,
              "snt_head_fwd: NULL tensor");
  SNT_REQUIRE(!training || B > 1, "snt_head_fwd: BatchNorm1d needs more than 1 value per channel in training");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* y = w.take<float>(B * E);
  w.take<float>(colsum_partial_count(B, E));
  if (!w.ok()) { set_error("snt_head_fwd: workspace too small"); return SNT_EWORKSPACE; }
  if (prec == SNT_PREC_BF16) {
    SNT_CHECK(bf16::linear_nt(pooled, w_fc, b_fc, B, E, K, y, (char*)ws + w.used, ws_bytes - w.used, st));
  } else {
    SNT_CHECK(gemm_f32(0, 1, B, E, K, 1.f, pooled, K, w_fc, K, 0.f, y, E, b_fc, st));
  }
  return bn_fwd(y, gamma, beta, running_mean, running_var, training, momentum, eps, B, E, features, yhat, rstd, st,
                features_bf16);
}

extern "C" int snt_head_bwd(int prec, const float* dfeatures, const float* pooled, const float* yhat,
                            const float* rstd, const float* gamma, int training, int64_t B, int64_t K, int64_t E,
                            float* d_w_fc, float* d_b_fc, float* d_gamma, float* d_beta,
                            void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_head_bwd: bad prec %d", prec);
  SNT_REQUIRE(B >= 1 && K >= 1 && E >= 1, "snt_head_bwd: bad sizes");
  SNT_REQUIRE(dfeatures && pooled && yhat && rstd && gamma && d_w_fc && d_b_fc && d_gamma && d_beta,
              "snt_head_bwd: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* dy = w.take<float>(B * E);
  float* part = w.take<float>(colsum_partial_count(B, E));
  if (!w.ok()) { set_error("snt_head_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(bn_bwd(dfeatures, yhat, rstd, gamma, training, B, E, dy, d_gamma, d_beta, st));
  // d_b_fc = column sums of dy: on the side stream, next to the weight-gradient contraction that reads the same dy
  SideStream* side = side_stream();
  if (side) {
    SNT_CUDA(cudaEventRecord(side->fork, st));
    SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    SNT_CHECK(colsum(dy, B, E, E, 0.f, d_b_fc, part, side->s));
    SNT_CUDA(cudaEventRecord(side->join, side->s));
  }
  if (prec == SNT_PREC_BF16) {
    SNT_CHECK(bf16::wgrad_tn(dy, pooled, B, E, K, d_w_fc, (char*)ws + w.used, ws_bytes - w.used, st));
  } else {
    SNT_CHECK(gemm_f32(1, 0, E, K, B, 1.f, dy, E, pooled, K, 0.f, d_w_fc, K, nullptr, st));
  }
  if (side) {
    SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return SNT_OK;
  }
  return colsum(dy, B, E, E, 0.f, d_b_fc, part, st);
}

// ------------------------------------------------------------------------------------------------------------
// a4-a6: gather + concat + pack
// ------------------------------------------------------------------------------------------------------------
extern "C" int snt_embed_pack_fwd(const float* features, const float* w_emb, const int64_t* captions,
                                  int64_t cap_stride, const int32_t* batch_sizes, int T, int64_t E, int64_t V,
                                  float* x_f32, void* x_bf16, void* stream) {
  PackInfo pk;
  SNT_CHECK(make_pack(batch_sizes, T, &pk));
  SNT_REQUIRE(E >= 1 && V >= 1 && features && w_emb, "snt_embed_pack_fwd: bad arguments");
  SNT_REQUIRE(T == 1 || (captions && cap_stride >= T - 1), "snt_embed_pack_fwd: captions narrower than T-1");
  SNT_REQUIRE(x_f32 || x_bf16, "snt_embed_pack_fwd: no output");
  return embed_pack_fwd(pk, features, w_emb, captions, cap_stride, E, V, x_f32, (__nv_bfloat16*)x_bf16,
                        (cudaStream_t)stream);
}

extern "C" int64_t snt_embed_bwd_workspace_bytes(int64_t N, int64_t V) {
  if (N < 1 || V < 1) return -1;
  return embed_bwd_ws_bytes(N, V);
}

extern "C" int snt_embed_pack_bwd(const float* dx, const int64_t* captions, int64_t cap_stride,
                                  const int32_t* batch_sizes, int T, int64_t B, int64_t E, int64_t V,
                                  float* dfeatures, float* d_w_emb, void* ws, int64_t ws_bytes, void* stream) {
  PackInfo pk;
  SNT_CHECK(make_pack(batch_sizes, T, &pk));
  SNT_REQUIRE(dx && E >= 1 && V >= 1 && B >= batch_sizes[0], "snt_embed_pack_bwd: bad arguments");
  SNT_REQUIRE(T == 1 || (captions && cap_stride >= T - 1), "snt_embed_pack_bwd: captions narrower than T-1");
  return embed_pack_bwd(pk, dx, captions, cap_stride, B, E, V, dfeatures, d_w_emb, ws, ws_bytes,
                        (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// a7: one LSTM layer over the packed sequence
// ------------------------------------------------------------------------------------------------------------
// Stream-ordered scratch for loops that contract many activation blocks against the SAME fp32 weight on the tensor cores
// (gemm_x3.cu): the weight is expanded into its bf16 slots once, each step expands only its activations.  Freed (on the
// stream) when the object goes out of scope.
namespace {
struct X3Loop {
  cudaStream_t st;
  void* base = nullptr;
  char* cur = nullptr;
  char* end = nullptr;
  explicit X3Loop(cudaStream_t s) : st(s) {}
  ~X3Loop() { x3::scratch_free(base, st); }
  int open(int64_t bytes) {
    SNT_CHECK(x3::scratch_alloc(&base, bytes, st));
    cur = (char*)base;
    end = cur + bytes;
    return SNT_OK;
  }
  void* take(int64_t bytes) {
    bytes = align_up(bytes, 256);
    if (!cur || cur + bytes > end) return nullptr;
    void* r = cur;
    cur += bytes;
    return r;
  }
  static int64_t need(int64_t elems) { return align_up(elems * 2, 256); }
};
}  // namespace

extern "C" int64_t snt_lstm_workspace_bytes(int prec, int64_t N, int64_t B, int64_t In, int64_t H) {
  if (!valid_prec(prec) || N < 1 || B < 1 || In < 1 || H < 1) return -1;
  if (prec == SNT_PREC_BF16) return bf16::lstm_ws_bytes(N, B, In, H);
  // fwd: bias sum [4H];  bwd: dh_rec [B,H], dc_state [B,H], colsum partials for [N,4H]
  return ws_bytes_for(4 * H, 4) + 2 * ws_bytes_for(B * H, 4) + ws_bytes_for(colsum_partial_count(N, 4 * H), 4);
}

extern "C" int snt_lstm_fwd(int prec, const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
                            const float* b_ih, const float* b_hh, const int32_t* batch_sizes, int T,
                            float* gates, float* cs, void* hs, void* hprev, void* ws, int64_t ws_bytes,
                            void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_lstm_fwd: bad prec %d", prec);
  PackInfo pk;
  SNT_CHECK(make_pack(batch_sizes, T, &pk));
  SNT_REQUIRE(In >= 1 && H >= 1, "snt_lstm_fwd: bad sizes");
  SNT_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && gates && cs && hs && hprev, "snt_lstm_fwd: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::lstm_fwd(pk, x, In, H, w_ih, w_hh, b_ih, b_hh, gates, cs, hs, hprev, ws, ws_bytes, st);

  const int64_t N = pk.off[T];
  Workspace w(ws, ws_bytes);
  float* bsum = w.take<float>(4 * H);
  if (!w.ok()) { set_error("snt_lstm_fwd: workspace too small"); return SNT_EWORKSPACE; }
  float* hs_f = (float*)hs;
  float* hp_f = (float*)hprev;
  SNT_CHECK(add_vec(b_ih, b_hh, bsum, 4 * H, st));
  // input projection for every timestep at once: gates = x . W_ih^T + (b_ih + b_hh)
  SNT_CHECK(gemm_f32(0, 1, N, 4 * H, In, 1.f, (const float*)x, In, w_ih, In, 0.f, gates, 4 * H, bsum, st));
  SNT_CUDA(cudaMemsetAsync(hp_f, 0, sizeof(float) * (size_t)pk.off[1] * H, st));  // h_{-1} = 0
  // tensor-core path of the recurrence: W_hh expanded once, h_{t-1} per step
  const int64_t B0 = pk.off[1];
  X3Loop xl(st);
  x3::Operand whh, hop;
  void* hbuf = nullptr;
  const bool tc_rec = T > 1 && x3::worth_it(B0, 4 * H, H);
  if (tc_rec) {
    SNT_CHECK(xl.open(X3Loop::need(x3::operand_elems(4 * H, H, true)) + X3Loop::need(x3::operand_elems(B0, H, true))));
    void* wbuf = xl.take(x3::operand_elems(4 * H, H, true) * 2);
    hbuf = xl.take(x3::operand_elems(B0, H, true) * 2);
    SNT_CHECK(x3::expand(w_hh, 4 * H, H, H, true, true, wbuf, &whh, st));
  }
  for (int t = 0; t < T; ++t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    float* g_t = gates + (int64_t)pk.off[t] * 4 * H;
    if (t > 0) {  // gates_t += h_{t-1} . W_hh^T
      if (tc_rec && bs >= 32) {
        SNT_CHECK(x3::expand(hp_f + (int64_t)pk.off[t] * H, bs, H, H, true, false, hbuf, &hop, st));
        SNT_CHECK(x3::gemm(hop, whh, bs, 4 * H, 1.f, 1.f, g_t, 4 * H, nullptr, nullptr, 0, st));
      } else {
        SNT_CHECK(gemm_f32(0, 1, bs, 4 * H, H, 1.f, hp_f + (int64_t)pk.off[t] * H, H, w_hh, H, 1.f, g_t, 4 * H,
                           nullptr, st));
      }
    }
    const float* c_prev = t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr;
    SNT_CHECK(lstm_point_fwd<float>(g_t, c_prev, cs + (int64_t)pk.off[t] * H, hs_f + (int64_t)pk.off[t] * H,
                                    bs_next > 0 ? hp_f + (int64_t)pk.off[t + 1] * H : nullptr, bs, bs_next, H, st));
  }
  return SNT_OK;
}

extern "C" int snt_lstm_bwd(int prec, const float* d_hs, float* gates, const float* cs, const void* hprev,
                            const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
                            const int32_t* batch_sizes, int T, float* d_w_ih, float* d_w_hh, float* d_bias,
                            float* dx, void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_lstm_bwd: bad prec %d", prec);
  PackInfo pk;
  SNT_CHECK(make_pack(batch_sizes, T, &pk));
  SNT_REQUIRE(In >= 1 && H >= 1, "snt_lstm_bwd: bad sizes");
  SNT_REQUIRE(d_hs && gates && cs && hprev && x && w_ih && w_hh && d_w_ih && d_w_hh && d_bias,
              "snt_lstm_bwd: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::lstm_bwd(pk, d_hs, gates, cs, hprev, x, In, H, w_ih, w_hh, d_w_ih, d_w_hh, d_bias, dx, ws,
                          ws_bytes, st);

  const int64_t N = pk.off[T];
  const int64_t B = pk.off[1];
  Workspace w(ws, ws_bytes);
  w.take<float>(4 * H);
  float* dh_rec = w.take<float>(B * H);
  float* dc_state = w.take<float>(B * H);
  float* part = w.take<float>(colsum_partial_count(N, 4 * H));
  if (!w.ok()) { set_error("snt_lstm_bwd: workspace too small"); return SNT_EWORKSPACE; }
  // tensor-core path of the reverse recurrence: W_hh (as the [4H, H] MN-major operand) expanded once, dG_t per step
  X3Loop xl(st);
  x3::Operand whh, gop;
  void* gbuf = nullptr;
  float* sws = nullptr;                      // K-split slices: [bs, H] has few output tiles
  const int64_t sws_elems = 8 * B * H;
  const bool tc_rec = T > 1 && x3::worth_it(B, H, 4 * H);
  if (tc_rec) {
    SNT_CHECK(xl.open(X3Loop::need(x3::operand_elems(4 * H, H, false)) + X3Loop::need(x3::operand_elems(B, 4 * H, true)) +
                      align_up(sws_elems * 4, 256)));
    sws = (float*)xl.take(sws_elems * 4);
    void* wbuf = xl.take(x3::operand_elems(4 * H, H, false) * 2);
    gbuf = xl.take(x3::operand_elems(B, 4 * H, true) * 2);
    SNT_CHECK(x3::expand(w_hh, 4 * H, H, H, false, true, wbuf, &whh, st));
  }
  for (int t = T - 1; t >= 0; --t) {
    const int bs = pk.off[t + 1] - pk.off[t];
    const int bs_next = t + 1 < T ? pk.off[t + 2] - pk.off[t + 1] : 0;
    float* g_t = gates + (int64_t)pk.off[t] * 4 * H;
    const float* c_prev = t > 0 ? cs + (int64_t)pk.off[t - 1] * H : nullptr;
    SNT_CHECK(lstm_point_bwd(g_t, cs + (int64_t)pk.off[t] * H, c_prev, d_hs + (int64_t)pk.off[t] * H, dh_rec,
                             dc_state, bs, bs_next, H, st));
    if (t > 0) {  // dh_{t-1} (recurrent part) = dG_t . W_hh
      if (tc_rec && bs >= 32) {
        SNT_CHECK(x3::expand(g_t, bs, 4 * H, 4 * H, true, false, gbuf, &gop, st));
        SNT_CHECK(x3::gemm(gop, whh, bs, H, 1.f, 0.f, dh_rec, H, nullptr, sws, sws_elems, st));
      } else {
        SNT_CHECK(gemm_f32(0, 0, bs, H, 4 * H, 1.f, g_t, 4 * H, w_hh, H, 0.f, dh_rec, H, nullptr, st));
      }
    }
  }
  // weight gradients over the whole packed sequence
  SNT_CHECK(gemm_f32(1, 0, 4 * H, In, N, 1.f, gates, 4 * H, (const float*)x, In, 0.f, d_w_ih, In, nullptr, st));
  SNT_CHECK(gemm_f32(1, 0, 4 * H, H, N, 1.f, gates, 4 * H, (const float*)hprev, H, 0.f, d_w_hh, H, nullptr, st));
  SNT_CHECK(colsum(gates, N, 4 * H, 4 * H, 0.f, d_bias, part, st));
  if (dx) SNT_CHECK(gemm_f32(0, 0, N, In, 4 * H, 1.f, gates, 4 * H, w_ih, In, 0.f, dx, In, nullptr, st));
  return SNT_OK;
}

// ------------------------------------------------------------------------------------------------------------
// a8: materialising vocab Linear
// ------------------------------------------------------------------------------------------------------------
extern "C" int64_t snt_linear_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V) {
  if (!valid_prec(prec) || N < 1 || H < 1 || V < 1) return -1;
  if (prec == SNT_PREC_BF16) return bf16::linear_ws_bytes(N, H, V);
  return ws_bytes_for(colsum_partial_count(N, V), 4);
}

extern "C" int snt_linear_fwd(int prec, const void* hs, const float* w_out, const float* b_out, int64_t N,
                              int64_t H, int64_t V, float* logits, void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_linear_fwd: bad prec %d", prec);
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && hs && w_out && b_out && logits, "snt_linear_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16) return bf16::linear_fwd(hs, w_out, b_out, N, H, V, logits, ws, ws_bytes, st);
  return gemm_f32(0, 1, N, V, H, 1.f, (const float*)hs, H, w_out, H, 0.f, logits, V, b_out, st);
}

extern "C" int snt_linear_bwd(int prec, const float* dlogits, const void* hs, const float* w_out, int64_t N,
                              int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out, void* ws,
                              int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_linear_bwd: bad prec %d", prec);
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && dlogits && hs && w_out && d_hs && d_w_out && d_b_out,
              "snt_linear_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::linear_bwd(dlogits, hs, w_out, N, H, V, d_hs, d_w_out, d_b_out, ws, ws_bytes, st);
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>(colsum_partial_count(N, V));
  if (!w.ok()) { set_error("snt_linear_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_CHECK(gemm_f32(0, 0, N, H, V, 1.f, dlogits, V, w_out, H, 0.f, d_hs, H, nullptr, st));
  SNT_CHECK(gemm_f32(1, 0, V, H, N, 1.f, dlogits, V, (const float*)hs, H, 0.f, d_w_out, H, nullptr, st));
  return colsum(dlogits, N, V, V, 0.f, d_b_out, part, st);
}

// ------------------------------------------------------------------------------------------------------------
// a8+a9: vocab Linear fused with log-softmax + cross-entropy.  fp32 mode: chunks of CE_CHUNK_ROWS rows of
// logits live in the workspace (L2-sized), never the whole [N,V].
// ------------------------------------------------------------------------------------------------------------
extern "C" int64_t snt_vocab_ce_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V) {
  if (!valid_prec(prec) || N < 1 || H < 1 || V < 1) return -1;
  if (prec == SNT_PREC_BF16) return bf16::vocab_ce_ws_bytes(N, H, V);
  const int64_t R = N < CE_CHUNK_ROWS ? N : CE_CHUNK_ROWS;
  return ws_bytes_for(R * V, 4) + ws_bytes_for(N, 4) + ws_bytes_for(colsum_partial_count(R, V), 4);
}

extern "C" int snt_vocab_ce_fwd(int prec, const void* hs, const float* w_out, const float* b_out,
                                const int64_t* targets, int64_t N, int64_t H, int64_t V, float* lse, float* loss,
                                void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_vocab_ce_fwd: bad prec %d", prec);
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && hs && w_out && b_out && targets && lse && loss,
              "snt_vocab_ce_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::vocab_ce_fwd(hs, w_out, b_out, targets, N, H, V, lse, loss, ws, ws_bytes, st);
  const int64_t R = N < CE_CHUNK_ROWS ? N : CE_CHUNK_ROWS;
  Workspace w(ws, ws_bytes);
  float* chunk = w.take<float>(R * V);
  float* nll = w.take<float>(N);
  if (!w.ok()) { set_error("snt_vocab_ce_fwd: workspace too small"); return SNT_EWORKSPACE; }
  const float* hs_f = (const float*)hs;
  // tensor-core path: W_out expanded into its bf16 slots once, the rows of Hs chunk by chunk (gemm_x3.cu)
  X3Loop xl(st);
  x3::Operand wx, ax;
  void* abuf = nullptr;
  const bool tc = x3::worth_it(R, V, H);
  if (tc) {
    SNT_CHECK(xl.open(X3Loop::need(x3::operand_elems(V, H, true)) + X3Loop::need(x3::operand_elems(R, H, true))));
    void* wbuf = xl.take(x3::operand_elems(V, H, true) * 2);
    abuf = xl.take(x3::operand_elems(R, H, true) * 2);
    SNT_CHECK(x3::expand(w_out, V, H, H, true, true, wbuf, &wx, st));
  }
  for (int64_t r0 = 0; r0 < N; r0 += R) {
    const int64_t r = N - r0 < R ? N - r0 : R;
    if (tc && r >= 32) {
      SNT_CHECK(x3::expand(hs_f + r0 * H, r, H, H, true, false, abuf, &ax, st));
      SNT_CHECK(x3::gemm(ax, wx, r, V, 1.f, 0.f, chunk, V, b_out, nullptr, 0, st));
    } else {
      SNT_CHECK(gemm_f32(0, 1, r, V, H, 1.f, hs_f + r0 * H, H, w_out, H, 0.f, chunk, V, b_out, st));
    }
    SNT_CHECK(ce_rows_fwd(chunk, r, V, V, targets + r0, lse + r0, nll + r0, st));
  }
  return reduce_sum(nll, N, 1.0f / (float)N, loss, st);
}

extern "C" int snt_vocab_ce_bwd(int prec, const void* hs, const float* w_out, const float* b_out,
                                const int64_t* targets, const float* lse, const float* dloss, float grad_scale,
                                int64_t N, int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out,
                                void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_vocab_ce_bwd: bad prec %d", prec);
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && hs && w_out && b_out && targets && lse && d_hs && d_w_out && d_b_out,
              "snt_vocab_ce_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::vocab_ce_bwd(hs, w_out, b_out, targets, lse, dloss, grad_scale, N, H, V, d_hs, d_w_out, d_b_out,
                              ws, ws_bytes, st);
  const int64_t R = N < CE_CHUNK_ROWS ? N : CE_CHUNK_ROWS;
  Workspace w(ws, ws_bytes);
  float* chunk = w.take<float>(R * V);
  w.take<float>(N);
  float* part = w.take<float>(colsum_partial_count(R, V));
  if (!w.ok()) { set_error("snt_vocab_ce_bwd: workspace too small"); return SNT_EWORKSPACE; }
  const float* hs_f = (const float*)hs;
  const float scale = grad_scale / (float)N;
  // tensor-core path: W_out expanded once per call in both roles (the [V,H] K-major operand of the logits recompute and
  // the [V,H] MN-major operand of dHs = dlogits . W_out); the chunk-dependent operands are expanded per chunk
  X3Loop xl(st);
  x3::Operand wx, wxt, ax, cx;
  void *abuf = nullptr, *cbuf = nullptr;
  float* sws = nullptr;                      // K-split slices of dHs (a chunk has few output tiles)
  const int64_t sws_elems = 8 * R * H;
  const bool tc = x3::worth_it(R, V, H);
  if (tc) {
    SNT_CHECK(xl.open(X3Loop::need(x3::operand_elems(V, H, true)) + X3Loop::need(x3::operand_elems(V, H, false)) +
                      X3Loop::need(x3::operand_elems(R, H, true)) + X3Loop::need(x3::operand_elems(R, V, true)) +
                      align_up(sws_elems * 4, 256)));
    sws = (float*)xl.take(sws_elems * 4);
    void* w1 = xl.take(x3::operand_elems(V, H, true) * 2);
    void* w2 = xl.take(x3::operand_elems(V, H, false) * 2);
    abuf = xl.take(x3::operand_elems(R, H, true) * 2);
    cbuf = xl.take(x3::operand_elems(R, V, true) * 2);
    SNT_CHECK(x3::expand(w_out, V, H, H, true, true, w1, &wx, st));
    SNT_CHECK(x3::expand(w_out, V, H, H, false, true, w2, &wxt, st));
  }
  for (int64_t r0 = 0; r0 < N; r0 += R) {
    const int64_t r = N - r0 < R ? N - r0 : R;
    const float acc = r0 > 0 ? 1.f : 0.f;
    if (tc && r >= 32) {
      SNT_CHECK(x3::expand(hs_f + r0 * H, r, H, H, true, false, abuf, &ax, st));
      SNT_CHECK(x3::gemm(ax, wx, r, V, 1.f, 0.f, chunk, V, b_out, nullptr, 0, st));
    } else {
      SNT_CHECK(gemm_f32(0, 1, r, V, H, 1.f, hs_f + r0 * H, H, w_out, H, 0.f, chunk, V, b_out, st));
    }
    SNT_CHECK(ce_rows_bwd(chunk, r, V, V, targets + r0, lse + r0, dloss, scale, st));
    if (tc && r >= 32) {
      SNT_CHECK(x3::expand(chunk, r, V, V, true, false, cbuf, &cx, st));
      SNT_CHECK(x3::gemm(cx, wxt, r, H, 1.f, 0.f, d_hs + r0 * H, H, nullptr, sws, sws_elems, st));
    } else {
      SNT_CHECK(gemm_f32(0, 0, r, H, V, 1.f, chunk, V, w_out, H, 0.f, d_hs + r0 * H, H, nullptr, st));
    }
    SNT_CHECK(gemm_f32(1, 0, V, H, r, 1.f, chunk, V, hs_f + r0 * H, H, acc, d_w_out, H, nullptr, st));
    SNT_CHECK(colsum(chunk, r, V, V, acc, d_b_out, part, st));
  }
  return SNT_OK;
}

// training variant with stored softmax numerators: one logits contraction per step (bf16 mode only)
extern "C" int64_t snt_vocab_ce_train_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V) {
  if (prec != SNT_PREC_BF16 || N < 1 || H < 1 || V < 1) return -1;
  return bf16::vocab_ce_train_ws_bytes(N, H, V);
}

extern "C" int snt_vocab_ce_train_fwd(int prec, const void* hs, const float* w_out, const float* b_out,
                                      const int64_t* targets, int64_t N, int64_t H, int64_t V, float* lse,
                                      float* loss, void* u, float* inv_s, void* hs_scaled, void* w_bf16, void* ws,
                                      int64_t ws_bytes, void* stream) {
  if (prec != SNT_PREC_BF16) {
    set_error("snt_vocab_ce_train_fwd: SNT_PREC_BF16 only (use snt_vocab_ce_fwd/bwd in fp32 mode)");
    return SNT_EUNSUPPORTED;
  }
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && hs && w_out && b_out && targets && lse && loss && u && inv_s &&
                  hs_scaled && w_bf16,
              "snt_vocab_ce_train_fwd: bad arguments");
  return bf16::vocab_ce_train_fwd(hs, w_out, b_out, targets, N, H, V, lse, loss, u, inv_s, hs_scaled, w_bf16, ws,
                                  ws_bytes, (cudaStream_t)stream);
}

extern "C" int snt_vocab_ce_train_bwd(int prec, const void* u, const float* inv_s, const void* hs_scaled,
                                      const void* w_bf16, const float* dloss, float grad_scale, int64_t N,
                                      int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out, void* ws,
                                      int64_t ws_bytes, void* stream) {
  if (prec != SNT_PREC_BF16) {
    set_error("snt_vocab_ce_train_bwd: SNT_PREC_BF16 only");
    return SNT_EUNSUPPORTED;
  }
  SNT_REQUIRE(N >= 1 && H >= 1 && V >= 1 && u && inv_s && hs_scaled && w_bf16 && d_hs && d_w_out && d_b_out,
              "snt_vocab_ce_train_bwd: bad arguments");
  return bf16::vocab_ce_train_bwd(u, inv_s, hs_scaled, w_bf16, dloss, grad_scale, N, H, V, d_hs, d_w_out, d_b_out,
                                  ws, ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// a12: greedy decode
// ------------------------------------------------------------------------------------------------------------
extern "C" int64_t snt_greedy_workspace_bytes(int prec, int64_t B, int64_t E, int64_t H, int64_t V, int L) {
  if (!valid_prec(prec) || B < 1 || E < 1 || H < 1 || V < 1 || L < 1 || L > SNT_MAX_LAYERS) return -1;
  if (prec == SNT_PREC_BF16) return bf16::greedy_ws_bytes(B, E, H, V, L);
  // x [B,E], per layer h,c [B,H] + bias sum [4H], gates [B,4H], logits [B,V], scratch h [B,H]
  return ws_bytes_for(B * E, 4) + L * (2 * ws_bytes_for(B * H, 4) + ws_bytes_for(4 * H, 4)) +
         ws_bytes_for(B * 4 * H, 4) + ws_bytes_for(B * V, 4);
}

extern "C" int snt_greedy_decode(int prec, const float* features, const float* w_emb, int L,
                                 const float* const* w_ih, const float* const* w_hh, const float* const* b_ih,
                                 const float* const* b_hh, const float* w_out, const float* b_out,
                                 const float* h0, const float* c0, int64_t B, int64_t E, int64_t H, int64_t V,
                                 int steps, int64_t* ids, void* ws, int64_t ws_bytes, void* stream) {
  SNT_REQUIRE(valid_prec(prec), "snt_greedy_decode: bad prec %d", prec);
  SNT_REQUIRE(B >= 1 && E >= 1 && H >= 1 && V >= 1 && L >= 1 && L <= SNT_MAX_LAYERS && steps >= 1,
              "snt_greedy_decode: bad sizes");
  SNT_REQUIRE(features && w_emb && w_ih && w_hh && b_ih && b_hh && w_out && b_out && ids,
              "snt_greedy_decode: NULL tensor");
  SNT_REQUIRE((h0 == nullptr) == (c0 == nullptr), "snt_greedy_decode: h0 and c0 must both be given or both NULL");
  for (int k = 0; k < L; ++k)
    SNT_REQUIRE(w_ih[k] && w_hh[k] && b_ih[k] && b_hh[k], "snt_greedy_decode: NULL weight for layer %d", k);
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == SNT_PREC_BF16)
    return bf16::greedy_decode(features, w_emb, L, w_ih, w_hh, b_ih, b_hh, w_out, b_out, h0, c0, B, E, H, V, steps,
                               ids, ws, ws_bytes, st);
  Workspace w(ws, ws_bytes);
  float* x = w.take<float>(B * E);
  float *h[SNT_MAX_LAYERS], *c[SNT_MAX_LAYERS], *bsum[SNT_MAX_LAYERS];
  for (int k = 0; k < L; ++k) {
    h[k] = w.take<float>(B * H);
    c[k] = w.take<float>(B * H);
    bsum[k] = w.take<float>(4 * H);
  }
  float* gates = w.take<float>(B * 4 * H);
  float* logits = w.take<float>(B * V);
  if (!w.ok()) { set_error("snt_greedy_decode: workspace too small"); return SNT_EWORKSPACE; }
  for (int k = 0; k < L; ++k) {
    SNT_CHECK(add_vec(b_ih[k], b_hh[k], bsum[k], 4 * H, st));
    if (h0) {
      SNT_CUDA(cudaMemcpyAsync(h[k], h0 + (int64_t)k * B * H, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
      SNT_CUDA(cudaMemcpyAsync(c[k], c0 + (int64_t)k * B * H, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
    } else {
      SNT_CUDA(cudaMemsetAsync(h[k], 0, sizeof(float) * B * H, st));
      SNT_CUDA(cudaMemsetAsync(c[k], 0, sizeof(float) * B * H, st));
    }
  }
  // Tensor-core path (gemm_x3.cu): every weight is expanded into its bf16 slots ONCE per decode; a step expands only its
  // activations (x_s and h_{s-1} per layer, h_s for the vocabulary projection).
  X3Loop xl(st);
  x3::Operand wih_x[SNT_MAX_LAYERS], whh_x[SNT_MAX_LAYERS], wout_x, a_in, a_h;
  void *abuf_in = nullptr, *abuf_h = nullptr;
  const bool tc = x3::worth_it(B, 4 * H, E < H ? E : H) && x3::worth_it(B, V, H);
  if (tc) {
    int64_t bytes = X3Loop::need(x3::operand_elems(V, H, true)) + X3Loop::need(x3::operand_elems(B, E > H ? E : H, true)) +
                    X3Loop::need(x3::operand_elems(B, H, true));
    for (int k = 0; k < L; ++k)
      bytes += X3Loop::need(x3::operand_elems(4 * H, k == 0 ? E : H, true)) + X3Loop::need(x3::operand_elems(4 * H, H, true));
    SNT_CHECK(xl.open(bytes));
    for (int k = 0; k < L; ++k) {
      const int64_t in_dim = k == 0 ? E : H;
      void* b1 = xl.take(x3::operand_elems(4 * H, in_dim, true) * 2);
      void* b2 = xl.take(x3::operand_elems(4 * H, H, true) * 2);
      SNT_CHECK(x3::expand(w_ih[k], 4 * H, in_dim, in_dim, true, true, b1, &wih_x[k], st));
      SNT_CHECK(x3::expand(w_hh[k], 4 * H, H, H, true, true, b2, &whh_x[k], st));
    }
    void* b3 = xl.take(x3::operand_elems(V, H, true) * 2);
    SNT_CHECK(x3::expand(w_out, V, H, H, true, true, b3, &wout_x, st));
    abuf_in = xl.take(x3::operand_elems(B, E > H ? E : H, true) * 2);
    abuf_h = xl.take(x3::operand_elems(B, H, true) * 2);
  }
  for (int s = 0; s < steps; ++s) {
    const float* inp = s == 0 ? features : x;
    int64_t in_dim = E;
    for (int k = 0; k < L; ++k) {
      if (tc) {
        SNT_CHECK(x3::expand(inp, B, in_dim, in_dim, true, false, abuf_in, &a_in, st));
        SNT_CHECK(x3::gemm(a_in, wih_x[k], B, 4 * H, 1.f, 0.f, gates, 4 * H, bsum[k], nullptr, 0, st));
        SNT_CHECK(x3::expand(h[k], B, H, H, true, false, abuf_h, &a_h, st));
        SNT_CHECK(x3::gemm(a_h, whh_x[k], B, 4 * H, 1.f, 1.f, gates, 4 * H, nullptr, nullptr, 0, st));
      } else {
        SNT_CHECK(gemm_f32(0, 1, B, 4 * H, in_dim, 1.f, inp, in_dim, w_ih[k], in_dim, 0.f, gates, 4 * H, bsum[k], st));
        SNT_CHECK(gemm_f32(0, 1, B, 4 * H, H, 1.f, h[k], H, w_hh[k], H, 1.f, gates, 4 * H, nullptr, st));
      }
      SNT_CHECK(lstm_point_fwd<float>(gates, c[k], c[k], h[k], nullptr, (int)B, 0, H, st));
      inp = h[k];
      in_dim = H;
    }
    if (tc) {
      SNT_CHECK(x3::expand(inp, B, H, H, true, false, abuf_h, &a_h, st));
      SNT_CHECK(x3::gemm(a_h, wout_x, B, V, 1.f, 0.f, logits, V, b_out, nullptr, 0, st));
    } else {
      SNT_CHECK(gemm_f32(0, 1, B, V, H, 1.f, inp, H, w_out, H, 0.f, logits, V, b_out, st));
    }
    SNT_CHECK(argmax_gather(logits, B, V, V, w_emb, E, ids + s, steps, x, st));
  }
  return SNT_OK;
}

extern "C" int snt_dp_adam_shard(const float* mc_g, float* mc_p, const float* p, float* m, float* v, int64_t lo,
                                 int64_t hi, double lr, double beta1, double beta2, double eps, float grad_clip,
                                 float grad_scale, int64_t step, int max_blocks, void* stream) {
  return dp_adam_shard(mc_g, mc_p, p, m, v, lo, hi, lr, beta1, beta2, eps, grad_clip, grad_scale, step, max_blocks,
                       (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// f4: words before the first <end> (eval.py:101-109)
// ------------------------------------------------------------------------------------------------------------
extern "C" int snt_caption_trim(const int64_t* ids, int64_t B, int steps, int64_t end_id, int64_t pad_id,
                                int32_t* lengths, int64_t* ids_out, void* stream) {
  SNT_REQUIRE(B >= 0 && steps >= 0, "snt_caption_trim: bad sizes");
  SNT_REQUIRE(B == 0 || steps == 0 || (ids && (lengths || ids_out)), "snt_caption_trim: NULL pointer");
  return caption_trim(ids, B, steps, end_id, pad_id, lengths, ids_out, (cudaStream_t)stream);
}
