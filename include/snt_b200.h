/* snt_b200.h — C ABI of the B200-native Show-and-Tell caption-decoder hot path.
 *
 * The reference (incredible-vision/show-and-tell) has no FFI of its own: its hot path is the Python
 * nn.Module pair in models.py calling into PyTorch.  Each entry point below replaces the PyTorch call
 * (or fused group of calls) cited next to it; `show-and-tell_b200/models.py` binds them with ctypes
 * behind the reference's EncoderCNN / DecoderRNN module API (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - All tensor pointers are DEVICE pointers unless marked [host].  The caller owns every buffer,
 *    including workspaces and tensors saved for backward; the library keeps no tensor memory.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, there are
 *    no hidden synchronisations.  Functions are re-entrant; one process drives one GPU.
 *  - Return value: 0 = ok, <0 = error (SNT_E*).  `snt_last_error()` returns a thread-local message.
 *  - Packed (time-major) layout, as torch.nn.utils.rnn.pack_padded_sequence produces it
 *    (models.py:51): `batch_sizes[t]` = #sequences longer than t (non-increasing, host int32[T]),
 *    off[t] = sum_{t'<t} batch_sizes[t'], packed row of (t, b) = off[t] + b, N = off[T].
 *  - `prec`: SNT_PREC_FP32 = the "fp32-accumulate" faithful mode: fp32 operands without rounding, fp32
 *    accumulation.  Large contractions run on the tcgen05 tensor cores with every fp32 value split exactly
 *    into three bf16 pieces and each product evaluated as its six significant partial products
 *    (csrc/gemm_x3.cu; contractions longer than 2048 are summed in slices so the accumulator's truncation
 *    stays below 6e-6); small ones on an FFMA kernel.  The large ones borrow stream-ordered scratch
 *    (cudaMallocAsync / cudaFreeAsync on the caller's stream) for the expanded operands.
 *    SNT_PREC_BF16 = bf16 operands on tcgen05 tensor cores with fp32 accumulation in TMEM.
 *    Tensors marked (act) are fp32 in FP32 mode and bf16 in BF16 mode.
 *  - There is no CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef SNT_B200_H
#define SNT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNT_ABI_VERSION 1
#define SNT_MAX_T 128 /* longest packed sequence (timesteps) accepted */
#define SNT_MAX_LAYERS 8

enum { SNT_OK = 0, SNT_EINVAL = -1, SNT_EUNSUPPORTED = -2, SNT_ECUDA = -3, SNT_EWORKSPACE = -4 };
enum { SNT_PREC_FP32 = 0, SNT_PREC_BF16 = 1 };

int snt_abi_version(void);
const char* snt_last_error(void);
/* sm_count, compute capability (major*10+minor), opt-in shared memory per block of `device`. */
int snt_device_query(int device, int* sm_count, int* cc, int64_t* smem_optin);
/* Sticky device-side flags (bit0: token id out of range in a gather).  Synchronises `stream`. */
int snt_read_flags(int* flags_out, int reset, void* stream);
/* Number of CUDA kernels this library has launched in this process (optionally resetting the counter). */
int64_t snt_launch_count(int reset);
/* Persistent tensor-core kernels normally launch one CTA per SM.  `n` > 0 leaves that many SMs free (grids shrink, tiles
 * are redistributed) so that concurrently running kernels of another stream - the NCCL gradient all-reduce of the
 * data-parallel step (train.py:43-44's DataParallel replaced by one process per GPU) - always find room and neither
 * stall nor push CTAs of these kernels into a second wave.  Returns the previous value.  Host-side, takes effect for
 * launches issued after the call. */
int snt_set_sm_reserve(int n);

/* ---- generic dense contractions (exported for the head, the drop-in Linear and unit tests) --------
 * C[M,N] = alpha * opA(A) . opB(B) + beta * C + bias[N]   (bias may be NULL), row-major C with ldc.
 * transA = 0: A is [M,K] row-major (lda);  1: A is [K,M] row-major.
 * transB = 0: B is [K,N] row-major (ldb);  1: B is [N,K] row-major  (the x.W^T case of nn.Linear). */
int snt_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                 float* C, int64_t ldc, const float* bias, void* stream);
/* Same contraction with bf16 operands (uint16 storage) on tcgen05/TMEM via TMA; C is fp32 or bf16
 * (c_is_bf16).  Leading dimensions must be multiples of 8 elements and bases 16-byte aligned. */
int snt_gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
                  const void* A, int64_t lda, const void* B, int64_t ldb, float beta,
                  void* C, int64_t ldc, int c_is_bf16, const float* bias, void* stream);
/* fp32 -> bf16 (round to nearest even), n elements. */
int snt_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---- a1/a2  EncoderCNN head: resnet.fc Linear(2048->E) + BatchNorm1d(E, momentum=0.01) ---------------
 * replaces models.py:27-28 applied to the pooled 2048-d features.  training!=0: batch statistics and
 * running-stat update (unbiased variance into running_var); else running stats.
 * saves yhat[B,E] and rstd[E] for backward.  workspace: snt_head_workspace_bytes. */
int64_t snt_head_workspace_bytes(int prec, int64_t B, int64_t K, int64_t E);
int snt_head_fwd(int prec, const float* pooled, const float* w_fc, const float* b_fc,
                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                 int training, float momentum, float eps, int64_t B, int64_t K, int64_t E,
                 float* features, float* yhat, float* rstd, void* ws, int64_t ws_bytes, void* stream);
/* autograd of the above w.r.t. the trainable head parameters (the backbone is frozen, models.py:14-15) */
int snt_head_bwd(int prec, const float* dfeatures, const float* pooled, const float* yhat,
                 const float* rstd, const float* gamma, int training, int64_t B, int64_t K, int64_t E,
                 float* d_w_fc, float* d_b_fc, float* d_gamma, float* d_beta,
                 void* ws, int64_t ws_bytes, void* stream);

/* ---- a4-a6  embedding gather + image-feature concat + pack  (models.py:49-51) ------------------------
 * x[off[t]+b,:] = features[b,:] (t==0) | w_emb[captions[b,t-1],:] (t>=1), for b < batch_sizes[t].
 * Writes x_f32 and/or x_bf16 (either may be NULL). */
int snt_embed_pack_fwd(const float* features, const float* w_emb, const int64_t* captions,
                       int64_t cap_stride, const int32_t* batch_sizes /*[host] T*/, int T,
                       int64_t E, int64_t V, float* x_f32, void* x_bf16, void* stream);
/* backward: dx[N,E] -> dfeatures[B,E] (rows >= batch_sizes[0] zero) and the dense embedding gradient
 * d_w_emb[V,E] (deterministic: rows are summed in packed-row order).  ws: snt_embed_bwd_workspace_bytes */
int64_t snt_embed_bwd_workspace_bytes(int64_t N, int64_t V);
int snt_embed_pack_bwd(const float* dx, const int64_t* captions, int64_t cap_stride,
                       const int32_t* batch_sizes /*[host] T*/, int T, int64_t B, int64_t E, int64_t V,
                       float* dfeatures, float* d_w_emb, void* ws, int64_t ws_bytes, void* stream);

/* ---- a7  one LSTM layer over the packed sequence  (models.py:52, nn.LSTM, gate rows i|f|g|o) ---------
 * h0 = c0 = 0.  x (act) [N,In].  Saves for backward: gates[N,4H] fp32 (post-activation i,f,g,o),
 * cs[N,H] fp32 (c_t), hs (act) [N,H] (h_t = the layer output), hprev (act) [N,H] (h_{t-1} per row). */
int64_t snt_lstm_workspace_bytes(int prec, int64_t N, int64_t B, int64_t In, int64_t H);
int snt_lstm_fwd(int prec, const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
                 const float* b_ih, const float* b_hh, const int32_t* batch_sizes /*[host] T*/, int T,
                 float* gates, float* cs, void* hs, void* hprev, void* ws, int64_t ws_bytes, void* stream);
/* BPTT for that layer.  d_hs[N,H] fp32 = gradient w.r.t. the layer output (read only).  `gates` is
 * overwritten with the pre-activation gradient dG[N,4H].  Outputs d_w_ih[4H,In], d_w_hh[4H,H],
 * d_bias[4H] (the same vector is the gradient of both b_ih and b_hh) and dx[N,In] fp32 (may be NULL). */
int snt_lstm_bwd(int prec, const float* d_hs, float* gates, const float* cs, const void* hprev,
                 const void* x, int64_t In, int64_t H, const float* w_ih, const float* w_hh,
                 const int32_t* batch_sizes /*[host] T*/, int T,
                 float* d_w_ih, float* d_w_hh, float* d_bias, float* dx,
                 void* ws, int64_t ws_bytes, void* stream);

/* ---- a8  vocab Linear, materialising  (models.py:53) and its autograd ---------------------------------*/
int64_t snt_linear_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V);
int snt_linear_fwd(int prec, const void* hs, const float* w_out, const float* b_out,
                   int64_t N, int64_t H, int64_t V, float* logits, void* ws, int64_t ws_bytes, void* stream);
int snt_linear_bwd(int prec, const float* dlogits, const void* hs, const float* w_out,
                   int64_t N, int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out,
                   void* ws, int64_t ws_bytes, void* stream);

/* ---- a8+a9  vocab Linear fused with log-softmax + cross-entropy  (models.py:53 + train.py:53,143) -----
 * loss = (1/N) sum_n (logsumexp_v logits[n,v] - logits[n,targets[n]]); logits[N,V] never reach HBM as
 * a whole.  Saves lse[N].  loss is a device scalar. */
int64_t snt_vocab_ce_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V);
int snt_vocab_ce_fwd(int prec, const void* hs, const float* w_out, const float* b_out,
                     const int64_t* targets, int64_t N, int64_t H, int64_t V,
                     float* lse, float* loss, void* ws, int64_t ws_bytes, void* stream);
/* backward of the fused loss: dlogits = (softmax - onehot) * grad_scale / N formed tile-wise.
 * `dloss` (device scalar, may be NULL = 1) multiplies grad_scale.  Outputs d_hs[N,H] fp32,
 * d_w_out[V,H], d_b_out[V]. */
int snt_vocab_ce_bwd(int prec, const void* hs, const float* w_out, const float* b_out,
                     const int64_t* targets, const float* lse, const float* dloss, float grad_scale,
                     int64_t N, int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out,
                     void* ws, int64_t ws_bytes, void* stream);

/* ---- a8+a9+a10, training variant: the logits contraction runs ONCE per step (SNT_PREC_BF16 only) -----------
 * Same mathematics as snt_vocab_ce_fwd + snt_vocab_ce_bwd (models.py:53 + train.py:53,143-144), for callers that
 * know at forward time that a backward follows.  The forward stores the softmax numerators
 * u[n,v] = exp(logit[n,v] - c[n]) as bf16 (c[n]: a per-row shift taken from the first 256 vocabulary columns, so no
 * second sweep over the logits is needed), patches the one-hot into the stored row and saves 1/sum_v u;  the backward
 * is then two contractions over `u`, without recomputing the logits.  Caller-owned, carried from fwd to bwd:
 *   u [N, ldu] bf16 with ldu = (V+7)/8*8,  inv_s [N] fp32,  hs_scaled [N,H] bf16,  w_bf16 [V,H] bf16.
 * A row whose largest logit exceeds the largest of its first 256 logits by more than ~88 would overflow: device flag
 * bit 2 (snt_read_flags) is raised and the loss is non-finite; use the two-call path above for such models.
 * Returns SNT_EUNSUPPORTED for SNT_PREC_FP32. */
int64_t snt_vocab_ce_train_workspace_bytes(int prec, int64_t N, int64_t H, int64_t V);
int snt_vocab_ce_train_fwd(int prec, const void* hs, const float* w_out, const float* b_out,
                           const int64_t* targets, int64_t N, int64_t H, int64_t V, float* lse, float* loss,
                           void* u, float* inv_s, void* hs_scaled, void* w_bf16,
                           void* ws, int64_t ws_bytes, void* stream);
int snt_vocab_ce_train_bwd(int prec, const void* u, const float* inv_s, const void* hs_scaled,
                           const void* w_bf16, const float* dloss, float grad_scale,
                           int64_t N, int64_t H, int64_t V, float* d_hs, float* d_w_out, float* d_b_out,
                           void* ws, int64_t ws_bytes, void* stream);

/* ---- a12  greedy decode  (models.py:56-67: fixed `steps` iterations, first-index argmax) ---------------
 * Per-layer weight pointer arrays are [host] arrays of device pointers.  h0/c0 [L,B,H] may be NULL
 * (zeros).  ids[B,steps] int64. */
int64_t snt_greedy_workspace_bytes(int prec, int64_t B, int64_t E, int64_t H, int64_t V, int L);
int snt_greedy_decode(int prec, const float* features, const float* w_emb, int L,
                      const float* const* w_ih, const float* const* w_hh,
                      const float* const* b_ih, const float* const* b_hh,
                      const float* w_out, const float* b_out, const float* h0, const float* c0,
                      int64_t B, int64_t E, int64_t H, int64_t V, int steps, int64_t* ids,
                      void* ws, int64_t ws_bytes, void* stream);

/* ---- f4  the caller-side tail of sample(): words before the first <end>  (eval.py:101-109) ---------------
 * lengths[b] = position of the first `end_id` in ids[b, 0..steps) (= steps when there is none), i.e. the
 * number of words eval.py keeps.  ids_out[B,steps] (optional; may alias ids) = ids with every position
 * >= lengths[b] replaced by pad_id.  At least one of lengths / ids_out must be given. */
int snt_caption_trim(const int64_t* ids, int64_t B, int steps, int64_t end_id, int64_t pad_id,
                     int32_t* lengths, int64_t* ids_out, void* stream);

/* ---- a11  clip_gradient (clamp to +-grad_clip) + Adam  (train.py:88-91,145-146) -------------------------
 * In-place on p, m, v; g is read only.  `step` is the 1-based count after this update.
 * grad_clip <= 0 disables the clamp.  grad_scale multiplies g first (1/world for averaged DP grads).
 * Hyper-parameters are doubles: torch.optim.Adam forms 1-beta and the bias corrections in double. */
int snt_clamp_adam(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                   double beta2, double eps, float grad_clip, float grad_scale, int64_t step, void* stream);

/* The same update for `count` parameter tensors in one launch: [host] arrays of device pointers and sizes. */
int snt_clamp_adam_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v,
                         const int64_t* n, double lr, double beta1, double beta2, double eps, float grad_clip,
                         float grad_scale, int64_t step, void* stream);

/* ---- data-parallel gradient exchange fused with clip_gradient + Adam  (replaces nn.DataParallel's gather / re-replicate,
 * train.py:43-44, and train.py:88-91,145-146, for one process per GPU on an NVLink / NVSwitch node) --------------------
 * The flat gradient buffer and the flat parameter buffer of every rank are mapped behind one multicast address each
 * (mc_g, mc_p; e.g. torch.distributed._symmetric_memory).  This rank owns the flat indices [lo, hi) (multiples of 4):
 * the sum of all ranks' gradients is read through the switch (multimem.ld_reduce), clamp + Adam run on this rank's m, v
 * and its local copy p of the parameters, and the updated parameters are stored to every rank (multimem.st).
 * The caller runs a cross-rank barrier before (every rank's gradients are written) and after (every rank's parameters
 * have arrived) on the same stream.  m and v are meaningful on the owner rank of an index only.
 * max_blocks > 0: a grid of at most that many (256-thread) blocks, for a bucket that is exchanged as background work
 * while the cooperative recurrence kernel runs on 128 of the SMs. */
int snt_dp_adam_shard(const float* mc_g, float* mc_p, const float* p, float* m, float* v, int64_t lo, int64_t hi,
                      double lr, double beta1, double beta2, double eps, float grad_clip, float grad_scale,
                      int64_t step, int max_blocks, void* stream);

/* ---- the whole teacher-forced step as ONE native call sequence  (train.py:137-146 for the models.py pair) --------
 * forward (encoder head -> gather/concat/pack -> L LSTM layers -> fused vocab-CE), backward (CE -> BPTT -> embedding
 * gradient and head backward on two streams) in the order of the stage functions above, with every activation carved
 * out of one caller-owned workspace.  No autograd graph, no per-stage host round trip: a step costs the host a few
 * function calls, so ragged batches (a different batch_sizes[] every step) run at the speed of the GPU without any
 * graph capture.  The packed geometry comes from the [host] batch_sizes of THIS call.
 *
 * Phases (bit mask for snt_step_run), to be issued in this order with the same descriptor and workspace; a
 * data-parallel caller starts its gradient all-reduce between them (one process per GPU; the reference's
 * nn.DataParallel, train.py:43-44):
 *   SNT_STEP_FWD       loss
 *   SNT_STEP_BWD_CE    d_w_out                                (ready first: overlaps all of BPTT; but see below)
 *   SNT_STEP_BWD_LSTM  d_b_out; d_w_ih / d_w_hh / d_b_ih / d_b_hh of every layer   (d_b_out = column sums of the stored
 *                      softmax numerators: one HBM-bound pass that runs beside the latency-bound BPTT recurrence)
 *   SNT_STEP_BWD_TAIL  d_w_emb and the head gradients (d_w_fc, d_b_fc, d_bn_w, d_bn_b), optional d_features
 * Gradients are WRITTEN (not accumulated) with the factor grad_scale folded in; `loss` = grad_scale * mean CE.
 * K = 0: no encoder head, `input` holds features[B,E] (then the head pointers may be NULL).
 * targets = NULL: targets are gathered on the device from `captions` (eval.py:91: pack(captions, lengths)). */
enum { SNT_STEP_FWD = 1, SNT_STEP_BWD_CE = 2, SNT_STEP_BWD_LSTM = 4, SNT_STEP_BWD_TAIL = 8, SNT_STEP_ALL = 15 };

typedef struct snt_step {
  int32_t struct_bytes;          /* = sizeof(snt_step): guards against a stale binding */
  int32_t prec;                  /* SNT_PREC_* */
  int32_t L, T;                  /* LSTM layers; timesteps of this batch */
  int32_t training;              /* head BatchNorm: batch statistics + running-stat update (models.py:28 in train()) */
  int32_t reserved;
  int64_t B, E, H, V, K;         /* captions, embed, hidden, vocab, pooled width (0: no head) */
  int64_t cap_stride;            /* row pitch of captions in elements */
  const int32_t* batch_sizes;    /* [host] T */
  const float* input;            /* pooled[B,K] (K > 0) or features[B,E] */
  const int64_t* captions;       /* [B, cap_stride] */
  const int64_t* targets;        /* [N] or NULL */
  const float *w_fc, *b_fc, *bn_w, *bn_b;
  float *bn_rm, *bn_rv;
  float bn_momentum, bn_eps;
  const float *w_emb, *w_out, *b_out;
  const float *w_ih[SNT_MAX_LAYERS], *w_hh[SNT_MAX_LAYERS], *b_ih[SNT_MAX_LAYERS], *b_hh[SNT_MAX_LAYERS];
  float *d_w_fc, *d_b_fc, *d_bn_w, *d_bn_b, *d_w_emb, *d_w_out, *d_b_out;
  float *d_w_ih[SNT_MAX_LAYERS], *d_w_hh[SNT_MAX_LAYERS], *d_b_ih[SNT_MAX_LAYERS], *d_b_hh[SNT_MAX_LAYERS];
  float* d_features;             /* optional [B,E]: gradient w.r.t. `input` when K = 0 */
  float grad_scale;
  float pad_;
  float* loss;                   /* device scalar */
  void* ws;
  int64_t ws_bytes;              /* >= snt_step_workspace_bytes(prec, L, B, N, E, H, V, K) with N >= sum(batch_sizes) */
} snt_step;

int64_t snt_step_workspace_bytes(int prec, int L, int64_t B, int64_t N, int64_t E, int64_t H, int64_t V, int64_t K);
int snt_step_run(const snt_step* step, int phases, void* stream);
/* 1 when a snt_step_run call whose mask holds BOTH SNT_STEP_BWD_CE and SNT_STEP_BWD_LSTM runs the d_w_out contraction
 * beside the BPTT recurrence instead of in front of it (bf16 mode, persistent recurrence kernel occupying at most two
 * thirds of the SMs: batch <= 1024 at H = 512): d_w_out is then final when BWD_LSTM ends, not when BWD_CE ends, and a
 * data-parallel caller exchanges it together with the LSTM gradients.  Issuing the two phases in separate calls keeps the
 * contraction in BWD_CE. */
int snt_step_overlaps_dw_out(int prec, int64_t B, int64_t H);

/* Per-stage device time of snt_step_run (diagnostics).  snt_step_profile(1) makes every later run record a CUDA-event
 * pair around each stage on the stream it is enqueued on; snt_step_profile_read synchronises the device and returns the
 * milliseconds of the last run per slot: 0 head fwd, 1 gather/pack (+ targets), 2+k LSTM fwd of layer k, 10 vocab-CE fwd,
 * 11 vocab-CE bwd, 12+k LSTM bwd of layer k, 20 embedding gradient, 21 head bwd (runs beside 20 on a second stream). */
#define SNT_STEP_PROFILE_SLOTS 22
int snt_step_profile(int enable);
int snt_step_profile_read(float* ms, int slots);

#ifdef __cplusplus
}
#endif
#endif /* SNT_B200_H */
