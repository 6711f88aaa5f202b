// SNT_PREC_BF16 pipelines: bf16 operands on tcgen05 tensor cores (fp32 accumulation in TMEM), TMA-fed.
// Declared here, implemented in bf16_path.cu on top of the GEMM core in gemm_tc.cu.
#pragma once
#include "common.cuh"

namespace snt {
namespace bf16 {

// y[M,N] = a[M,K] . w[N,K]^T + bias   (fp32 in / fp32 out; operands rounded to bf16 in the workspace)
int64_t head_extra_ws_bytes(int64_t B, int64_t K, int64_t E);
int linear_nt(const float* a, const float* w, const float* bias, int64_t M, int64_t N, int64_t K, float* y,
              void* ws, int64_t ws_bytes, cudaStream_t st);
// dw[N,K] = dy[M,N]^T . a[M,K]
int wgrad_tn(const float* dy, const float* a, int64_t M, int64_t N, int64_t K, float* dw, void* ws,
             int64_t ws_bytes, cudaStream_t st);

int64_t lstm_ws_bytes(int64_t N, int64_t B, int64_t In, int64_t H);
// What lstm_fwd / lstm_bwd derive from the fp32 weights before their first contraction (gate-interleaved bf16 copies, the
// summed bias, the transposed W_hh' of the persistent BPTT kernel) and the regions they clear (h_{-1}, arrival
// counters).  A caller that runs whole steps (csrc/step.cu) prepares all of it in caller-owned memory with
// lstm_prepare() on a side stream, beside the encoder head, and hands it to both calls: they then launch no preparation
// kernel of their own.
struct LstmPrepared {
  __nv_bfloat16* w_ih;    // [4H, In]
  __nv_bfloat16* w_hh;    // [4H, H]
  __nv_bfloat16* w_hh_t;  // [H, 4H]
  float* bsum;            // [4H]
  int* flags_fwd;         // lstm_prepared_flag_ints(B) each, cleared by lstm_prepare
  int* flags_bwd;
};
int64_t lstm_prepared_flag_ints(int64_t B);
// also clears the first B*H elements of `hprev` (h_{-1} = 0)
int lstm_prepare(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int64_t In, int64_t H,
                 int64_t B, const LstmPrepared& out, void* hprev, cudaStream_t st);
int lstm_fwd(const PackInfo& pk, const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
             const float* b_ih, const float* b_hh, float* gates, float* cs, void* hs, void* hprev, void* ws,
             int64_t ws_bytes, cudaStream_t st, const LstmPrepared* prep = nullptr);
int lstm_bwd(const PackInfo& pk, const float* d_hs, float* gates, const float* cs, const void* hprev,
             const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh, float* d_w_ih,
             float* d_w_hh, float* d_bias, float* dx, void* ws, int64_t ws_bytes, cudaStream_t st,
             const LstmPrepared* prep = nullptr, bool* wgrad_pending = nullptr);
// wgrad_pending != NULL allows lstm_bwd to return with dX complete in stream order but the weight and bias gradients
// still in flight on side streams (*wgrad_pending = true); the caller orders `st` after them with lstm_bwd_join(st)
// before anything reads d_w_ih / d_w_hh / d_bias or reuses the stage workspace.
int lstm_bwd_join(cudaStream_t st);

// step executor hooks (lstm_tc.cu): event recorded right before the next lstm_bwd() launches its recurrence; whether the
// recurrence of this hidden size runs as the persistent cooperative kernel (128 of 148 SMs, latency-bound)
void lstm_bwd_gate_event(cudaEvent_t e);
bool lstm_bwd_is_persistent(int64_t H);
int lstm_bwd_persistent_ctas(int64_t B, int64_t H);  // SMs its widest launch occupies (0: per-step launches)

int64_t linear_ws_bytes(int64_t N, int64_t H, int64_t V);
int linear_fwd(const void* hs, const float* w_out, const float* b_out, int64_t N, int64_t H, int64_t V,
               float* logits, void* ws, int64_t ws_bytes, cudaStream_t st);
int linear_bwd(const float* dlogits, const void* hs, const float* w_out, int64_t N, int64_t H, int64_t V,
               float* d_hs, float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st);

int64_t vocab_ce_ws_bytes(int64_t N, int64_t H, int64_t V);
int vocab_ce_fwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, int64_t N,
                 int64_t H, int64_t V, float* lse, float* loss, void* ws, int64_t ws_bytes, cudaStream_t st,
                 float loss_scale = 1.f);
int vocab_ce_bwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, const float* lse,
                 const float* dloss, float grad_scale, int64_t N, int64_t H, int64_t V, float* d_hs,
                 float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st);

// Training path with stored softmax numerators (vocab_ce_tc.cu): the logits contraction runs once per step.
// u [N, pad8(V)] bf16, inv_s [N], hs_scaled [N,H] bf16 and w_bf16 [V,H] bf16 are caller-owned and carried from fwd to bwd.
int64_t vocab_ce_train_ws_bytes(int64_t N, int64_t H, int64_t V);
int vocab_ce_train_fwd(const void* hs, const float* w_out, const float* b_out, const int64_t* targets, int64_t N,
                       int64_t H, int64_t V, float* lse, float* loss, void* u, float* inv_s, void* hs_scaled,
                       void* w_bf16, void* ws, int64_t ws_bytes, cudaStream_t st, float loss_scale = 1.f,
                       bool w_prepared = false);  // w_prepared: w_bf16 already holds bf16(w_out)
int vocab_ce_train_bwd(const void* u, const float* inv_s, const void* hs_scaled, const void* w_bf16,
                       const float* dloss, float grad_scale, int64_t N, int64_t H, int64_t V, float* d_hs,
                       float* d_w_out, float* d_b_out, void* ws, int64_t ws_bytes, cudaStream_t st,
                       bool defer_bias = false, bool defer_dw = false);
// the dW_out half of vocab_ce_train_bwd(defer_dw = true), on any stream, with caller-owned split-K scratch
int64_t vocab_ce_train_sws_elems(int64_t N, int64_t H, int64_t V);
int vocab_ce_train_dw(const void* u, const void* hs_scaled, const float* dloss, float grad_scale, int64_t N, int64_t H,
                      int64_t V, float* d_w_out, float* sws, int64_t sws_elems, cudaStream_t st, int max_ctas);
// the d_b_out half of vocab_ce_train_bwd(defer_bias = true), on any stream, with caller-owned scratch
int64_t vocab_ce_train_bias_part_elems(int64_t N, int64_t V);
int vocab_ce_train_bias(const void* u, const float* inv_s, const float* dloss, float grad_scale, int64_t N, int64_t V,
                        float* d_b_out, float* part, float* db, cudaStream_t st, int max_blocks);

// out[c] = beta*out[c] + sum_r roww[r] * in[r*ld + c] for a bf16 matrix (roww = NULL: unweighted); partial needs
// ((R+255)/256)*C floats
int colsum_bf16(const __nv_bfloat16* in, int64_t R, int64_t C, int64_t ld, float beta, float* out, float* partial,
                cudaStream_t st, const float* roww = nullptr, int max_blocks = 0);

int64_t greedy_ws_bytes(int64_t B, int64_t E, int64_t H, int64_t V, int L);
int greedy_decode(const float* features, const float* w_emb, int L, const float* const* w_ih,
                  const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                  const float* w_out, const float* b_out, const float* h0, const float* c0, int64_t B, int64_t E,
                  int64_t H, int64_t V, int steps, int64_t* ids, void* ws, int64_t ws_bytes, cudaStream_t st);

}  // namespace bf16
}  // namespace snt
