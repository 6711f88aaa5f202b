// Launchers of the HBM-bound / pointwise kernels (kernels.cu), shared by the stage functions in api.cu.
#pragma once
#include "common.cuh"

namespace snt {

int embed_pack_fwd(const PackInfo& pk, const float* features, const float* w_emb, const int64_t* captions,
                   int64_t cap_stride, int64_t E, int64_t V, float* x_f32, __nv_bfloat16* x_bf16,
                   cudaStream_t st, int64_t row0 = 0, int64_t row1 = -1);  // packed rows [row0, row1) only (-1: all)

// out[c] = beta*out[c] + sum_r in[r*ld + c], deterministic two-pass; partial needs colsum_partial_count(R,C) floats
int64_t colsum_partial_count(int64_t R, int64_t C);
int colsum(const float* in, int64_t R, int64_t C, int64_t ld, float beta, float* out, float* partial,
           cudaStream_t st);

int add_vec(const float* a, const float* b, float* out, int64_t n, cudaStream_t st);
int cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st, bool chain = false);
int cast_f32(const __nv_bfloat16* src, float* dst, int64_t n, cudaStream_t st);

// one LSTM timestep, pointwise part.  gates_t: [bs,4H] pre-activations in, activations out.
template <typename ActT>
int lstm_point_fwd(float* gates_t, const float* c_prev, float* cs_t, ActT* hs_t, ActT* hprev_next, int bs,
                   int bs_next, int64_t H, cudaStream_t st);
int lstm_point_bwd(float* gates_t, const float* cs_t, const float* c_prev, const float* d_hs_t,
                   const float* dh_rec, float* dc_state, int bs, int bs_next, int64_t H, cudaStream_t st);

// rows of logits [R,V] (ld): lse[r], nll[r] = lse - logits[r,target]
int ce_rows_fwd(const float* logits, int64_t R, int64_t V, int64_t ld, const int64_t* targets, float* lse,
                float* nll, cudaStream_t st);
// in place: logits <- (exp(logits - lse) - onehot) * scale * (dloss ? *dloss : 1)
int ce_rows_bwd(float* logits, int64_t R, int64_t V, int64_t ld, const int64_t* targets, const float* lse,
                const float* dloss, float scale, cudaStream_t st);
// out[0] = scale * sum(v[0..n))   (single block, fixed-order tree: deterministic)
int reduce_sum(const float* v, int64_t n, float scale, float* out, cudaStream_t st);

// greedy step tail: ids[b*ids_stride] = first argmax_v logits[b,:]; x_next[b,:] = w_emb[id,:]
int argmax_gather(const float* logits, int64_t B, int64_t V, int64_t ld, const float* w_emb, int64_t E,
                  int64_t* ids, int64_t ids_stride, float* x_next, cudaStream_t st);

int bn_fwd(const float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
           int training, float momentum, float eps, int64_t B, int64_t E, float* out, float* yhat,
           float* rstd, cudaStream_t st, __nv_bfloat16* out_bf16 = nullptr);  // out_bf16: `out` again, rounded to bf16
// snt_head_fwd with the features also written as bf16 rows (the t = 0 rows of the packed LSTM input), api.cu
int head_fwd(int prec, const float* pooled, const float* w_fc, const float* b_fc, const float* gamma, const float* beta,
             float* running_mean, float* running_var, int training, float momentum, float eps, int64_t B, int64_t K,
             int64_t E, float* features, float* yhat, float* rstd, void* ws, int64_t ws_bytes, cudaStream_t st,
             __nv_bfloat16* features_bf16);
int bn_bwd(const float* dout, const float* yhat, const float* rstd, const float* gamma, int training,
           int64_t B, int64_t E, float* dy, float* dgamma, float* dbeta, cudaStream_t st);

int64_t embed_bwd_ws_bytes(int64_t N, int64_t V);
int embed_pack_bwd(const PackInfo& pk, const float* dx, const int64_t* captions, int64_t cap_stride,
                   int64_t B, int64_t E, int64_t V, float* dfeatures, float* d_w_emb, void* ws,
                   int64_t ws_bytes, cudaStream_t st, int phase = 0);

int clamp_adam(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
               double eps, float grad_clip, float grad_scale, int64_t step, cudaStream_t st);

int clamp_adam_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v,
                     const int64_t* n, double lr, double beta1, double beta2, double eps, float grad_clip,
                     float grad_scale, int64_t step, cudaStream_t st);

// the data-parallel exchange fused with the optimizer over multicast memory (kernels.cu): flat indices [lo, hi) of this rank
int dp_adam_shard(const float* mc_g, float* mc_p, const float* p, float* m, float* v, int64_t lo, int64_t hi, double lr,
                  double beta1, double beta2, double eps, float grad_clip, float grad_scale, int64_t step, int max_blocks,
                  cudaStream_t st);

// lengths[b] = index of the first end_id in ids[b,0..steps) (steps if none); ids_out (optional, may alias ids) = ids with
// everything from that position on replaced by pad_id
int caption_trim(const int64_t* ids, int64_t B, int steps, int64_t end_id, int64_t pad_id, int32_t* lengths,
                 int64_t* ids_out, cudaStream_t st);

}  // namespace snt
