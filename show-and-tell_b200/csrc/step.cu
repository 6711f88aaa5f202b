// The teacher-forced training step as one native call sequence (include/snt_b200.h: snt_step_run): the stage functions
// of api.cu in the order train.py:137-146 runs them for the models.py pair, with every activation carved out of one
// caller-owned workspace.  What the Python autograd glue (ops.py) costs per step - ~20 ctypes calls, ~40 tensor
// allocations, the autograd graph: 0.75-1.0 ms of host time, more than the GPU needs for the step - shrinks to one to
// four calls, so a ragged batch (new batch_sizes[] every step) runs eagerly at GPU speed, and a data-parallel caller
// gets three natural points to start its gradient all-reduce.
//
// Two things run beside the main stream (validated bit-identical in round 2, profiles/r02_switch_sweep.txt):
//   * the token-dependent half of the embedding gradient (histogram + scan of the caption ids) is enqueued on the
//     executor's side stream during the forward pass - it needs the captions only;
//   * the head backward (5 small launches) runs on that stream next to the embedding-gradient kernels: dfeatures is
//     literally dx[:B], the t = 0 rows of the first layer's input gradient.
#include <stdlib.h>

#include "kernels.cuh"
#include "bf16.cuh"

namespace snt { namespace tc { int grid_sms(); int set_grid_cap(int n); } }
using namespace snt;

namespace {

// targets[off[t] + b] = captions[b, t]  for b < batch_sizes[t]   (pack_padded_sequence(captions, lengths)[0], eval.py:91)
__global__ void __launch_bounds__(256)
pack_targets_kernel(const __grid_constant__ PackInfo pk, const int64_t* __restrict__ captions, int64_t cap_stride,
                    int64_t* __restrict__ targets) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= pk.off[pk.T]) return;
  int lo = 0, hi = pk.T - 1;  // largest t with off[t] <= n
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (pk.off[mid] <= n) lo = mid; else hi = mid - 1;
  }
  targets[n] = captions[(int64_t)(n - pk.off[lo]) * cap_stride + lo];
}

__global__ void scale_scalar_kernel(float* x, float s) { x[0] *= s; }

inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

// Optional per-stage timing (snt_step_profile): one CUDA-event pair per stage slot, recorded on the stream the stage is
// enqueued on.  Off by default (no events, no overhead).
enum { ST_HEAD_F = 0, ST_EMBED_F = 1, ST_LSTM_F = 2, ST_CE_F = 10, ST_CE_B = 11, ST_LSTM_B = 12, ST_EMBED_B = 20,
       ST_HEAD_B = 21, ST_COUNT = SNT_STEP_PROFILE_SLOTS };
struct StageProf {
  bool on = false;
  cudaEvent_t ev[ST_COUNT][2] = {};
  bool used[ST_COUNT] = {};
};
StageProf g_prof;
struct StageTimer {
  int slot; cudaStream_t st; bool on;
  StageTimer(int slot_, cudaStream_t st_) : slot(slot_), st(st_), on(g_prof.on) {
    if (!on) return;
    for (int i = 0; i < 2; ++i)
      if (!g_prof.ev[slot][i] && cudaEventCreate(&g_prof.ev[slot][i]) != cudaSuccess) { on = false; return; }
    cudaEventRecord(g_prof.ev[slot][0], st);
  }
  ~StageTimer() {
    if (!on) return;
    cudaEventRecord(g_prof.ev[slot][1], st);
    g_prof.used[slot] = true;
  }
};

// dW_out beside the BPTT recurrence (see snt_step_run): bf16 mode, the recurrence runs as the persistent kernel and leaves
// at least a third of the SMs free, and the two side streams exist.
bool dw_out_beside_bptt(int prec, int64_t B, int64_t H) {
  if (prec != SNT_PREC_BF16 || getenv("SNT_NO_BIAS_DEFER") || getenv("SNT_NO_DW_DEFER")) return false;
  if (!side_stream(1) || !side_stream(2)) return false;
  const int ctas = bf16::lstm_bwd_persistent_ctas(B, H);
  return ctas > 0 && tc::grid_sms() - ctas >= tc::grid_sms() / 3;
}

// moves the end of a stage's span to "now" on `st` (a stage whose last kernels were left in flight on side streams and
// joined later: its span then ends at the join, i.e. it is reported conservatively, including what ran beside it)
void stage_extend(int slot, cudaStream_t st) {
  if (g_prof.on && g_prof.used[slot] && g_prof.ev[slot][1]) cudaEventRecord(g_prof.ev[slot][1], st);
}

struct Layer { float* gates; float* cs; void* hs; void* hprev; };
struct StepBufs {
  float *feats, *yhat, *rstd;
  int64_t* targets;
  void* x;
  Layer layer[SNT_MAX_LAYERS];
  float *lse, *inv_s, *d_hs, *dx[2];
  void *u, *hs_scaled, *w_bf16;
  float *bias_part, *bias_db;  // scratch of the deferred d_b_out column sums (bf16 mode)
  float* dw_sws;               // split-K scratch of the deferred dW_out contraction (bf16 mode, persistent BPTT)
  int64_t dw_sws_elems;
  bf16::LstmPrepared prep[SNT_MAX_LAYERS];  // bf16 mode: what the recurrence calls derive from the fp32 weights
  bool has_prep;
  void *scratch, *head_ws, *emb_ws;
  int64_t scratch_bytes, head_bytes, emb_bytes;
  bool ok;
};

int64_t act_bytes(int prec) { return prec == SNT_PREC_BF16 ? 2 : 4; }

int64_t stage_scratch_bytes(int prec, int L, int64_t B, int64_t N, int64_t E, int64_t H, int64_t V, int64_t K) {
  int64_t s = 0;
  if (K > 0) s = max64(s, snt_head_workspace_bytes(prec, B, K, E));
  for (int k = 0; k < L; ++k) s = max64(s, snt_lstm_workspace_bytes(prec, N, B, k == 0 ? E : H, H));
  s = max64(s, prec == SNT_PREC_BF16 ? bf16::vocab_ce_train_ws_bytes(N, H, V) : snt_vocab_ce_workspace_bytes(prec, N, H, V));
  return s;
}

// One bump allocation, identical for every phase of a step (it depends on the descriptor's sizes and N only).
StepBufs carve(int prec, int L, int64_t B, int64_t N, int64_t E, int64_t H, int64_t V, int64_t K, bool own_targets,
               void* ws, int64_t ws_bytes) {
  Workspace w(ws, ws_bytes);
  StepBufs b;
  const int64_t ab = act_bytes(prec);
  b.feats = K > 0 ? w.take<float>(B * E) : nullptr;
  b.yhat = K > 0 ? w.take<float>(B * E) : nullptr;
  b.rstd = K > 0 ? w.take<float>(E) : nullptr;
  b.targets = own_targets ? w.take<int64_t>(N) : nullptr;
  b.x = w.take<char>(N * E * ab);
  for (int k = 0; k < L; ++k) {
    b.layer[k].gates = w.take<float>(N * 4 * H);
    b.layer[k].cs = w.take<float>(N * H);
    b.layer[k].hs = w.take<char>(N * H * ab);
    b.layer[k].hprev = w.take<char>(N * H * ab);
  }
  b.lse = w.take<float>(N);
  b.inv_s = w.take<float>(N);
  b.d_hs = w.take<float>(N * H);
  const int64_t dxw = E > H ? E : H;
  b.dx[0] = w.take<float>(N * dxw);
  b.dx[1] = L > 1 ? w.take<float>(N * dxw) : nullptr;
  if (prec == SNT_PREC_BF16) {
    b.u = w.take<char>(N * ((V + 7) / 8 * 8) * 2);
    b.hs_scaled = w.take<char>(N * H * 2);
    b.w_bf16 = w.take<char>(V * H * 2);
    b.bias_part = w.take<float>(bf16::vocab_ce_train_bias_part_elems(N, V));
    b.bias_db = w.take<float>(V);
  } else {
    b.u = b.hs_scaled = b.w_bf16 = nullptr;
    b.bias_part = b.bias_db = nullptr;
  }
  b.scratch_bytes = stage_scratch_bytes(prec, L, B, N, E, H, V, K);
  b.scratch = w.take<char>(b.scratch_bytes);
  b.head_bytes = K > 0 ? snt_head_workspace_bytes(prec, B, K, E) : 0;
  b.head_ws = K > 0 ? w.take<char>(b.head_bytes) : nullptr;
  b.emb_bytes = snt_embed_bwd_workspace_bytes(N, V);
  b.emb_ws = w.take<char>(b.emb_bytes);
  b.dw_sws_elems = (prec == SNT_PREC_BF16 && bf16::lstm_bwd_is_persistent(H)) ? bf16::vocab_ce_train_sws_elems(N, H, V) : 0;
  b.dw_sws = b.dw_sws_elems > 0 ? w.take<float>(b.dw_sws_elems) : nullptr;
  b.has_prep = prec == SNT_PREC_BF16;
  for (int k = 0; k < L && b.has_prep; ++k) {
    const int64_t In = k == 0 ? E : H, fl = bf16::lstm_prepared_flag_ints(B);
    b.prep[k].w_ih = w.take<__nv_bfloat16>(4 * H * In);
    b.prep[k].w_hh = w.take<__nv_bfloat16>(4 * H * H);
    b.prep[k].w_hh_t = w.take<__nv_bfloat16>(4 * H * H);
    b.prep[k].bsum = w.take<float>(4 * H);
    b.prep[k].flags_fwd = w.take<int>(2 * fl);  // forward counters, then the BPTT's
    b.prep[k].flags_bwd = b.prep[k].flags_fwd ? b.prep[k].flags_fwd + fl : nullptr;
  }
  b.ok = w.ok();
  return b;
}

}  // namespace

extern "C" int64_t snt_step_workspace_bytes(int prec, int L, int64_t B, int64_t N, int64_t E, int64_t H, int64_t V,
                                            int64_t K) {
  if ((prec != SNT_PREC_FP32 && prec != SNT_PREC_BF16) || L < 1 || L > SNT_MAX_LAYERS || B < 1 || N < B || E < 1 ||
      H < 1 || V < 1 || K < 0)
    return -1;
  // measure the carve with a null base: every take() advances `used` only when it fits, so size it generously first
  const int64_t ab = act_bytes(prec);
  const int64_t dxw = E > H ? E : H;
  int64_t t = 0;
  auto add = [&](int64_t bytes) { t += align_up(bytes, 256); };
  if (K > 0) { add(B * E * 4); add(B * E * 4); add(E * 4); }
  add(N * 8);
  add(N * E * ab);
  for (int k = 0; k < L; ++k) { add(N * 4 * H * 4); add(N * H * 4); add(N * H * ab); add(N * H * ab); }
  add(N * 4); add(N * 4); add(N * H * 4);
  add(N * dxw * 4);
  if (L > 1) add(N * dxw * 4);
  if (prec == SNT_PREC_BF16) {
    add(N * ((V + 7) / 8 * 8) * 2); add(N * H * 2); add(V * H * 2);
    add(bf16::vocab_ce_train_bias_part_elems(N, V) * 4); add(V * 4);
  }
  add(stage_scratch_bytes(prec, L, B, N, E, H, V, K));
  if (K > 0) add(snt_head_workspace_bytes(prec, B, K, E));
  add(snt_embed_bwd_workspace_bytes(N, V));
  if (prec == SNT_PREC_BF16 && bf16::lstm_bwd_is_persistent(H)) add(bf16::vocab_ce_train_sws_elems(N, H, V) * 4);
  for (int k = 0; k < L && prec == SNT_PREC_BF16; ++k) {
    const int64_t In = k == 0 ? E : H;
    add(4 * H * In * 2); add(4 * H * H * 2); add(4 * H * H * 2); add(4 * H * 4);
    add(2 * bf16::lstm_prepared_flag_ints(B) * 4);
  }
  return t;
}

extern "C" int snt_step_overlaps_dw_out(int prec, int64_t B, int64_t H) { return dw_out_beside_bptt(prec, B, H) ? 1 : 0; }

extern "C" int snt_step_run(const snt_step* d, int phases, void* stream) {
  SNT_REQUIRE(d != nullptr && d->struct_bytes == (int32_t)sizeof(snt_step),
              "snt_step_run: descriptor size mismatch (binding built against another header?)");
  SNT_REQUIRE(d->prec == SNT_PREC_FP32 || d->prec == SNT_PREC_BF16, "snt_step_run: bad prec %d", d->prec);
  SNT_REQUIRE(d->L >= 1 && d->L <= SNT_MAX_LAYERS, "snt_step_run: L=%d outside [1,%d]", d->L, SNT_MAX_LAYERS);
  SNT_REQUIRE((phases & ~SNT_STEP_ALL) == 0 && phases != 0, "snt_step_run: bad phase mask %d", phases);
  PackInfo pk;
  SNT_CHECK(make_pack(d->batch_sizes, d->T, &pk));
  const int prec = d->prec, L = d->L, T = d->T;
  const int64_t B = d->B, E = d->E, H = d->H, V = d->V, K = d->K, N = pk.off[T];
  SNT_REQUIRE(B >= 1 && E >= 1 && H >= 1 && V >= 1 && K >= 0, "snt_step_run: bad sizes");
  SNT_REQUIRE(d->batch_sizes[0] == B, "snt_step_run: batch_sizes[0]=%d does not match B=%lld", d->batch_sizes[0],
              (long long)B);
  SNT_REQUIRE(d->input && d->w_emb && d->w_out && d->b_out && d->loss, "snt_step_run: NULL tensor");
  SNT_REQUIRE(T == 1 || (d->captions && d->cap_stride >= T - 1), "snt_step_run: captions narrower than T-1");
  SNT_REQUIRE(d->targets || (d->captions && d->cap_stride >= T),
              "snt_step_run: targets=NULL needs captions at least T wide (pack(captions, lengths), eval.py:91)");
  SNT_REQUIRE(K == 0 || (d->w_fc && d->b_fc && d->bn_w && d->bn_b && d->bn_rm && d->bn_rv),
              "snt_step_run: NULL head parameter");
  for (int k = 0; k < L; ++k)
    SNT_REQUIRE(d->w_ih[k] && d->w_hh[k] && d->b_ih[k] && d->b_hh[k], "snt_step_run: NULL LSTM weight, layer %d", k);
  StepBufs b = carve(prec, L, B, N, E, H, V, K, d->targets == nullptr, d->ws, d->ws_bytes);
  if (!b.ok) {
    set_error("snt_step_run: workspace too small (%lld bytes given, %lld needed)", (long long)d->ws_bytes,
              (long long)snt_step_workspace_bytes(prec, L, B, N, E, H, V, K));
    return SNT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SideStream* side = side_stream(1);
  const int64_t* targets = d->targets ? d->targets : b.targets;
  const float* feats = K > 0 ? b.feats : d->input;
  const bool bf = prec == SNT_PREC_BF16;
  // d_b_out beside the BPTT instead of beside the vocabulary contractions: only where the recurrence runs as the
  // persistent cooperative kernel (latency-bound, HBM idle, 20 SMs free); elsewhere the per-step contractions fill the GPU
  // (measured at configs[3], H = 1024, per-step launches: moving the pass there - narrow or as background work - cost the
  // BPTT stage as much as it saved the vocabulary stage)
  int st_prio = 0;
  const bool background = side != nullptr && cudaStreamGetPriority((cudaStream_t)stream, &st_prio) == cudaSuccess &&
                          st_prio < 0 && !getenv("SNT_NO_BACKGROUND");
  // Chained launches (common.cuh) pay where the step is a sequence of short dependent launches around the two persistent
  // recurrence kernels (configs[1]: -9 us per step).  Where the recurrence is a train of per-step launches (H > 512) the
  // step measured the same or slower with them (configs[3], profiles/r02_chained_launches.txt): plain stream order there;
  // the per-step forward chain keeps the explicit chaining it always had.
  PdlSuppress plain_order(!(bf && bf16::lstm_bwd_is_persistent(H)));
  const bool defer_bias = bf && side != nullptr && bf16::lstm_bwd_is_persistent(H) && !getenv("SNT_NO_BIAS_DEFER");
  // dW_out beside the BPTT too: the recurrence needs dHs only, and with two row blocks per CTA (csrc/lstm_tc.cu) it
  // occupies 64 of the 148 SMs at batch 1024.  The contraction runs on its own side stream from a grid capped at the SMs
  // the recurrence leaves free, released by the same gate event as the column sums.  Only when this call covers both
  // phases (d_w_out is final when BWD_LSTM ends: a caller that exchanges linear.weight between the two phases runs them
  // separately and keeps the contraction in BWD_CE), and only when enough SMs are left for it to end within the
  // recurrence.
  SideStream* side_dw = side_stream(2);
  const int bptt_ctas = bf ? bf16::lstm_bwd_persistent_ctas(B, H) : 0;
  const int free_sms = tc::grid_sms() - bptt_ctas;
  const bool defer_dw = defer_bias && side_dw != nullptr && (phases & SNT_STEP_BWD_CE) && (phases & SNT_STEP_BWD_LSTM) &&
                        b.dw_sws != nullptr && dw_out_beside_bptt(prec, B, H);

  // Weight preparation off the critical path: the bf16 gate-interleaved LSTM weights (both directions' layouts), the
  // summed bias, the cleared recurrence counters and bf16(W_out) depend on the parameters only, so they are produced on a
  // side stream while the main stream runs the encoder head and the gather; lstm_fwd / lstm_bwd / the vocabulary stage
  // then start with their first contraction.  (Per step: two preparation launches per layer and the 20 MB cast used to
  // sit in front of the recurrences and the vocabulary pass, ~30 us.)  SNT_NO_EARLY_PREP=1: each stage prepares for itself.
  const bool early_prep = bf && b.has_prep && side_dw != nullptr && (phases & SNT_STEP_FWD) && !getenv("SNT_NO_EARLY_PREP");
  // The embedding-gradient buffer (V x E floats, 10 MB at configs[1]) is zeroed beside the forward pass, on the stream
  // that builds the token plan, when this call starts a training step (forward and the first backward phase together:
  // the tail of backward follows by contract).  The tail - in this call or a later one on the same workspace - then
  // finds the marker and skips the two memsets in front of its first kernel.  SNT_NO_EMB_ZERO_EARLY=1: zero in the tail.
  static thread_local const void* emb_zeroed_ws = nullptr;   // workspace whose embedding gradient this thread's last forward phase zeroed
  static thread_local const float* emb_zeroed_dw = nullptr;
  const bool emb_zero_early = side != nullptr && (phases & SNT_STEP_FWD) && (phases & SNT_STEP_BWD_CE) &&
                              d->d_w_emb != nullptr && E % 4 == 0 && E <= 1024 && !getenv("SNT_NO_EMB_ZERO_EARLY");
  if (phases & SNT_STEP_FWD) {
    emb_zeroed_ws = emb_zero_early ? d->ws : nullptr;
    emb_zeroed_dw = emb_zero_early ? d->d_w_emb : nullptr;
  }
  // later phases of a step whose forward phase prepared early find the prepared buffers in the workspace
  const bool use_prep = bf && b.has_prep && side_dw != nullptr && !getenv("SNT_NO_EARLY_PREP");
  if (phases & SNT_STEP_FWD) {
    // The side streams' work is ENQUEUED after the main stream's first kernels (a caller that reads the loss back every
    // step has an idle GPU at this point: whatever is enqueued first starts first), but ordered behind the fork events
    // recorded here, i.e. it runs beside the head.
    if (early_prep) SNT_CUDA(cudaEventRecord(side_dw->fork, st));
    if (side) SNT_CUDA(cudaEventRecord(side->fork, st));
    // The packed rows of t >= 1 (rows [B, N)) are embeddings of caption tokens: they do not depend on the head, so with
    // early preparation they are gathered on the side stream beside the head, and the B rows of t = 0 are written as bf16
    // by the head's BatchNorm kernel itself: no gather launch is left between the head and the recurrence.  The front of
    // a step is a chain of small dependent launches at 5-12 us each (profiles/r02_step_front.txt); what counts is how
    // many of them the recurrence has to wait for.
    const bool gather_aside = early_prep && N > B && !getenv("SNT_NO_GATHER_ASIDE");
    const bool targets_aside = side != nullptr && !d->targets;  // pack(captions, lengths): needed by the vocabulary stage only
    const bool x0_by_head = gather_aside && K > 0;
    if (K > 0) {
      StageTimer tm(ST_HEAD_F, st);
      SNT_CHECK(head_fwd(prec, d->input, d->w_fc, d->b_fc, d->bn_w, d->bn_b, d->bn_rm, d->bn_rv, d->training,
                         d->bn_momentum, d->bn_eps, B, K, E, b.feats, b.yhat, b.rstd, b.scratch, b.scratch_bytes, st,
                         x0_by_head ? (__nv_bfloat16*)b.x : nullptr));
    }
    {
    StageTimer tm(ST_EMBED_F, st);
    if (!d->targets && !targets_aside) {
      pack_targets_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pk, d->captions, d->cap_stride, b.targets);
      SNT_LAUNCH_CHECK("pack_targets_kernel");
    }
    if (!x0_by_head)
      SNT_CHECK(embed_pack_fwd(pk, feats, d->w_emb, d->captions, d->cap_stride, E, V, bf ? nullptr : (float*)b.x,
                               bf ? (__nv_bfloat16*)b.x : nullptr, st, 0, gather_aside ? B : -1));
    }
    if (early_prep) {
      cudaStream_t ps = side_dw->s;
      SNT_CUDA(cudaStreamWaitEvent(ps, side_dw->fork, 0));
      for (int k = 0; k < L; ++k)
        SNT_CHECK(bf16::lstm_prepare(d->w_ih[k], d->w_hh[k], d->b_ih[k], d->b_hh[k], k == 0 ? E : H, H, B, b.prep[k],
                                     b.layer[k].hprev, ps));
      if (gather_aside)
        SNT_CHECK(embed_pack_fwd(pk, feats, d->w_emb, d->captions, d->cap_stride, E, V, nullptr, (__nv_bfloat16*)b.x, ps,
                                 B, N));
      SNT_CUDA(cudaEventRecord(side_dw->aux, ps));    // the recurrence's inputs are ready
      SNT_CHECK(cast_bf16(d->w_out, (__nv_bfloat16*)b.w_bf16, V * H, ps));
      SNT_CUDA(cudaEventRecord(side_dw->join, ps));   // and bf16(W_out)
    }
    if (side) {
      SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
      if (targets_aside) {
        pack_targets_kernel<<<(unsigned)((N + 255) / 256), 256, 0, side->s>>>(pk, d->captions, d->cap_stride, b.targets);
        SNT_LAUNCH_CHECK("pack_targets_kernel");
        SNT_CUDA(cudaEventRecord(side->aux, side->s));
      }
      // token-dependent half of the embedding gradient: needs the captions only
      // (the whole step in one call: the gradient buffer is zeroed here too, off the critical path of the tail)
      SNT_CHECK(embed_pack_bwd(pk, nullptr, d->captions, d->cap_stride, B, emb_zero_early ? E : 0, V, nullptr,
                               emb_zero_early ? d->d_w_emb : nullptr, b.emb_ws, b.emb_bytes, side->s, 1));
      SNT_CUDA(cudaEventRecord(side->join, side->s));
    }
    const void* inp = b.x;
    int64_t in_dim = E;
    if (early_prep) SNT_CUDA(cudaStreamWaitEvent(st, side_dw->aux, 0));
    for (int k = 0; k < L; ++k) {
      StageTimer tm(ST_LSTM_F + k, st);
      if (early_prep)
        SNT_CHECK(bf16::lstm_fwd(pk, inp, in_dim, H, d->w_ih[k], d->w_hh[k], d->b_ih[k], d->b_hh[k], b.layer[k].gates,
                                 b.layer[k].cs, b.layer[k].hs, b.layer[k].hprev, b.scratch, b.scratch_bytes, st,
                                 &b.prep[k]));
      else
        SNT_CHECK(snt_lstm_fwd(prec, inp, in_dim, H, d->w_ih[k], d->w_hh[k], d->b_ih[k], d->b_hh[k], d->batch_sizes, T,
                               b.layer[k].gates, b.layer[k].cs, b.layer[k].hs, b.layer[k].hprev, b.scratch,
                               b.scratch_bytes, st));
      inp = b.layer[k].hs;
      in_dim = H;
    }
    if (early_prep) SNT_CUDA(cudaStreamWaitEvent(st, side_dw->join, 0));
    if (targets_aside) SNT_CUDA(cudaStreamWaitEvent(st, side->aux, 0));
    StageTimer tm_c(ST_CE_F, st);
    if (bf) {
      SNT_CHECK(bf16::vocab_ce_train_fwd(inp, d->w_out, d->b_out, targets, N, H, V, b.lse, d->loss, b.u, b.inv_s,
                                         b.hs_scaled, b.w_bf16, b.scratch, b.scratch_bytes, st, d->grad_scale,
                                         early_prep));
    } else {
      SNT_CHECK(snt_vocab_ce_fwd(prec, inp, d->w_out, d->b_out, targets, N, H, V, b.lse, d->loss, b.scratch,
                                 b.scratch_bytes, st));
      if (d->grad_scale != 1.f) {
        scale_scalar_kernel<<<1, 1, 0, st>>>(d->loss, d->grad_scale);
        SNT_LAUNCH_CHECK("scale_scalar_kernel");
      }
    }
  }

  if (phases & SNT_STEP_BWD_CE) {
    SNT_REQUIRE(d->d_w_out && d->d_b_out, "snt_step_run: NULL output-layer gradient");
    const void* hs_last = b.layer[L - 1].hs;
    StageTimer tm(ST_CE_B, st);
    if (bf)  // with a side stream, d_b_out is deferred to the BPTT phase (below)
      SNT_CHECK(bf16::vocab_ce_train_bwd(b.u, b.inv_s, b.hs_scaled, b.w_bf16, nullptr, d->grad_scale, N, H, V, b.d_hs,
                                         d->d_w_out, d->d_b_out, b.scratch, b.scratch_bytes, st, defer_bias, defer_dw));
    else
      SNT_CHECK(snt_vocab_ce_bwd(prec, hs_last, d->w_out, d->b_out, targets, b.lse, nullptr, d->grad_scale, N, H, V,
                                 b.d_hs, d->d_w_out, d->d_b_out, b.scratch, b.scratch_bytes, st));
  }

  // The first layer's weight and bias gradients may stay in flight on side streams beyond BWD_LSTM when the tail of
  // backward follows in the same call (bf16::lstm_bwd, deferred mode): nothing before the optimizer reads them, and the
  // tail's small kernels run beside the two contractions instead of behind them.  A caller that exchanges the LSTM
  // gradients between the two phases issues them in separate calls and finds them final when BWD_LSTM returns.
  const bool tail_follows = (phases & SNT_STEP_BWD_LSTM) && (phases & SNT_STEP_BWD_TAIL);
  bool wgrad_pending = false;
  // gradient w.r.t. the input of layer k lands in dx[k & 1]; layer 0's is dx[0]
  if (phases & SNT_STEP_BWD_LSTM) {
    SNT_REQUIRE(d->d_b_out, "snt_step_run: NULL output-layer gradient");
    bool bias_pending = false;
    int bias_blocks = 0;
    if (defer_bias) {
      // d_b_out = column sums of the stored softmax numerators (one more pass over the N x V matrix): HBM-bound, it needs
      // no tensor core and only a few SMs.  Next to the two vocabulary contractions that read the same matrix it costs
      // them ~55 us of bandwidth; here it runs on the SMs the cooperative BPTT kernel leaves free (128 of 148 busy, HBM
      // idle).  Two ways to stay out of the recurrence's way.  If the caller's stream has a higher priority than the side stream
      // (parallel.DataParallelStep runs the step on a high-priority stream), the pass is BACKGROUND work: the usual wide
      // grid of short blocks, released at the moment the cooperative kernel becomes eligible (gate event), so the block
      // scheduler places the recurrence first and the column sums soak up whatever is left - two or more blocks per
      // free SM.  On a default-priority stream: a grid narrow enough never to take an SM the recurrence needs.
      if (background) {
        bf16::lstm_bwd_gate_event(side->fork);   // recorded by snt_lstm_bwd of the top layer, after its weight preparation
      } else {
        SNT_CUDA(cudaEventRecord(side->fork, st));
      }
      bias_pending = true;
      bias_blocks = background ? 0 : (defer_dw ? 8 : (free_sms < 8 ? 8 : free_sms));
    }
    const int dw_ctas = background ? free_sms : free_sms - 8;
    const float* d_out = b.d_hs;
    for (int k = L - 1; k >= 0; --k) {
      SNT_REQUIRE(d->d_w_ih[k] && d->d_w_hh[k] && d->d_b_ih[k] && d->d_b_hh[k],
                  "snt_step_run: NULL LSTM gradient, layer %d", k);
      const void* inp = k == 0 ? b.x : b.layer[k - 1].hs;
      const int64_t in_dim = k == 0 ? E : H;
      float* dx = b.dx[k & 1];
      StageTimer tm(ST_LSTM_B + k, st);
      if (use_prep)
        SNT_CHECK(bf16::lstm_bwd(pk, d_out, b.layer[k].gates, b.layer[k].cs, b.layer[k].hprev, inp, in_dim, H, d->w_ih[k],
                                 d->w_hh[k], d->d_w_ih[k], d->d_w_hh[k], d->d_b_ih[k], dx, b.scratch, b.scratch_bytes, st,
                                 &b.prep[k], (k == 0 && tail_follows) ? &wgrad_pending : nullptr));
      else
        SNT_CHECK(snt_lstm_bwd(prec, d_out, b.layer[k].gates, b.layer[k].cs, b.layer[k].hprev, inp, in_dim, H, d->w_ih[k],
                               d->w_hh[k], d->batch_sizes, T, d->d_w_ih[k], d->d_w_hh[k], d->d_b_ih[k], dx, b.scratch,
                               b.scratch_bytes, st));
      // b_ih and b_hh enter the gates as a sum: they receive the same gradient
      if (!(k == 0 && wgrad_pending))
        SNT_CUDA(cudaMemcpyAsync(d->d_b_hh[k], d->d_b_ih[k], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, st));
      if (bias_pending) {  // enqueued after the top layer's launches: its gate event has been recorded by now
        bias_pending = false;
        SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
        SNT_CHECK(bf16::vocab_ce_train_bias(b.u, b.inv_s, nullptr, d->grad_scale, N, V, d->d_b_out, b.bias_part,
                                            b.bias_db, side->s, bias_blocks));
        SNT_CUDA(cudaEventRecord(side->aux, side->s));
        if (defer_dw) {
          SNT_CUDA(cudaStreamWaitEvent(side_dw->s, side->fork, 0));
          SNT_CHECK(bf16::vocab_ce_train_dw(b.u, b.hs_scaled, nullptr, d->grad_scale, N, H, V, d->d_w_out, b.dw_sws,
                                            b.dw_sws_elems, side_dw->s, dw_ctas));
          SNT_CUDA(cudaEventRecord(side_dw->join, side_dw->s));
        }
      }
      d_out = dx;
    }
    if (defer_bias) SNT_CUDA(cudaStreamWaitEvent(st, side->aux, 0));  // d_b_out is final when this phase ends
    if (defer_dw) SNT_CUDA(cudaStreamWaitEvent(st, side_dw->join, 0));  // and so is d_w_out
  }

  if (phases & SNT_STEP_BWD_TAIL) {
    SNT_REQUIRE(d->d_w_emb, "snt_step_run: NULL embedding gradient");
    SNT_REQUIRE(K == 0 || (d->d_w_fc && d->d_b_fc && d->d_bn_w && d->d_bn_b), "snt_step_run: NULL head gradient");
    const float* dx0 = b.dx[0];
    const bool fork_head = K > 0 && side != nullptr;
    const bool emb_zeroed = side != nullptr && emb_zeroed_ws == d->ws && emb_zeroed_dw == d->d_w_emb;
    emb_zeroed_ws = nullptr;
    emb_zeroed_dw = nullptr;
    if (side) SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));  // the embedding plan of the forward phase
    if (fork_head) {
      // dfeatures = dx0[:B] (the t = 0 rows); the head backward does not depend on the embedding-gradient kernels
      SNT_CUDA(cudaEventRecord(side->fork, st));
      SNT_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
      {
        StageTimer tm(ST_HEAD_B, side->s);
        SNT_CHECK(snt_head_bwd(prec, dx0, d->input, b.yhat, b.rstd, d->bn_w, d->training, B, K, E, d->d_w_fc, d->d_b_fc,
                               d->d_bn_w, d->d_bn_b, b.head_ws, b.head_bytes, side->s));
      }
      SNT_CUDA(cudaEventRecord(side->join, side->s));
    }
    {
    StageTimer tm(ST_EMBED_B, st);
    if (side) {
      SNT_CHECK(embed_pack_bwd(pk, dx0, d->captions, d->cap_stride, B, E, V, K == 0 ? d->d_features : nullptr,
                               d->d_w_emb, b.emb_ws, b.emb_bytes, st, emb_zeroed ? 3 : 2));
    } else {
      SNT_CHECK(embed_pack_bwd(pk, dx0, d->captions, d->cap_stride, B, E, V, K == 0 ? d->d_features : nullptr,
                               d->d_w_emb, b.emb_ws, b.emb_bytes, st, 0));
    }
    }
    if (wgrad_pending) {
      SNT_CHECK(bf16::lstm_bwd_join(st));
      stage_extend(ST_LSTM_B, st);  // layer 0's stage ends where its weight gradients are final
      SNT_CUDA(cudaMemcpyAsync(d->d_b_hh[0], d->d_b_ih[0], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, st));
    }
    if (fork_head) {
      SNT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    } else if (K > 0) {
      StageTimer tm(ST_HEAD_B, st);
      SNT_CHECK(snt_head_bwd(prec, dx0, d->input, b.yhat, b.rstd, d->bn_w, d->training, B, K, E, d->d_w_fc, d->d_b_fc,
                             d->d_bn_w, d->d_bn_b, b.head_ws, b.head_bytes, st));
    }
  }
  return SNT_OK;
}

// ---- per-stage timing of snt_step_run (diagnostics; used by bench.py for the roofline of the dominant stage) -----------
extern "C" int snt_step_profile(int enable) {
  g_prof.on = enable != 0;
  for (int i = 0; i < ST_COUNT; ++i) g_prof.used[i] = false;
  return SNT_OK;
}

// ms[slot] = device time of the LAST recorded run of each stage (0 for a stage that did not run).  Synchronises the device.
extern "C" int snt_step_profile_read(float* ms, int slots) {
  SNT_REQUIRE(ms != nullptr && slots >= ST_COUNT, "snt_step_profile_read: need %d slots", (int)ST_COUNT);
  SNT_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < slots; ++i) ms[i] = 0.f;
  for (int i = 0; i < ST_COUNT; ++i) {
    if (!g_prof.used[i]) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.ev[i][0], g_prof.ev[i][1]) == cudaSuccess) ms[i] = t;
    else cudaGetLastError();
  }
  return SNT_OK;
}
