// HBM-bound and pointwise kernels of the caption-decoder path: gather/concat/pack, LSTM gate math,
// cross-entropy rows, argmax, BatchNorm, embedding-gradient scatter, clamp+Adam.  sm_100a.
#include "kernels.cuh"

namespace snt {

static inline unsigned nblocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

__device__ __forceinline__ int find_step(const PackInfo& p, int row) {
  int lo = 0, hi = p.T;  // invariant: off[lo] <= row < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (p.off[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------------------
// a4-a6: x[off[t]+b] = features[b] (t==0) | w_emb[captions[b,t-1]]   (models.py:49-51).  One warp per row.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_pack_fwd_kernel(const __grid_constant__ PackInfo pk, const float* __restrict__ features,
                      const float* __restrict__ w_emb, const int64_t* __restrict__ captions,
                      int64_t cap_stride, int E, int64_t V, float* __restrict__ x_f32,
                      __nv_bfloat16* __restrict__ x_bf16, int* flags, int row0, int row1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = row0 + blockIdx.x * 8 + warp;
  if (row >= row1) return;
  const int t = find_step(pk, row);
  const int b = row - pk.off[t];
  const float* src;
  if (t == 0) {
    src = features + (int64_t)b * E;
  } else {
    int64_t tok = captions[(int64_t)b * cap_stride + (t - 1)];
    if (tok < 0 || tok >= V) {
      if (lane == 0) atomicOr(flags, 1);
      src = nullptr;
    } else {
      src = w_emb + tok * E;
    }
  }
  const int64_t o = (int64_t)row * E;
  if ((E & 3) == 0) {
    for (int e = lane * 4; e < E; e += 128) {
      float4 v = src ? *reinterpret_cast<const float4*>(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (x_f32) *reinterpret_cast<float4*>(x_f32 + o + e) = v;
      if (x_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk2;
        pk2.x = *reinterpret_cast<uint32_t*>(&lo);
        pk2.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(x_bf16 + o + e) = pk2;
      }
    }
  } else {
    for (int e = lane; e < E; e += 32) {
      float v = src ? src[e] : 0.f;
      if (x_f32) x_f32[o + e] = v;
      if (x_bf16) x_bf16[o + e] = __float2bfloat16_rn(v);
    }
  }
}

int embed_pack_fwd(const PackInfo& pk, const float* features, const float* w_emb, const int64_t* captions,
                   int64_t cap_stride, int64_t E, int64_t V, float* x_f32, __nv_bfloat16* x_bf16,
                   cudaStream_t st, int64_t row0, int64_t row1) {
  const int N = pk.off[pk.T];
  if (row1 < 0 || row1 > N) row1 = N;
  if (row0 >= row1) return SNT_OK;
  embed_pack_fwd_kernel<<<nblocks(row1 - row0, 8), 256, 0, st>>>(pk, features, w_emb, captions, cap_stride, (int)E, V,
                                                                x_f32, x_bf16, device_flags(), (int)row0, (int)row1);
  SNT_LAUNCH_CHECK("embed_pack_fwd_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// column sums, deterministic two pass: partial[chunk][c] then fixed-order sum over chunks
// ---------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS_PER_CHUNK = 256;
int64_t colsum_partial_count(int64_t R, int64_t C) { return ((R + CS_ROWS_PER_CHUNK - 1) / CS_ROWS_PER_CHUNK) * C; }

__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ in, int64_t R, int64_t C, int64_t ld, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS_PER_CHUNK;
  const int64_t r1 = min(R, r0 + CS_ROWS_PER_CHUNK);
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += in[r * ld + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int64_t chunks, int64_t C, float beta,
                                    float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int64_t k = 0; k < chunks; ++k) s += partial[k * C + c];
  out[c] = (beta != 0.f ? beta * out[c] : 0.f) + s;
}
int colsum(const float* in, int64_t R, int64_t C, int64_t ld, float beta, float* out, float* partial,
           cudaStream_t st) {
  if (C <= 0) return SNT_OK;
  const int64_t chunks = (R + CS_ROWS_PER_CHUNK - 1) / CS_ROWS_PER_CHUNK;
  if (chunks > 0) {
    SNT_REQUIRE(partial != nullptr, "colsum: no workspace");
    dim3 grid(nblocks(C, 32), (unsigned)chunks);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(in, R, C, ld, partial);
    SNT_LAUNCH_CHECK("colsum_partial_kernel");
  }
  colsum_final_kernel<<<nblocks(C, 256), 256, 0, st>>>(partial, chunks, C, beta, out);
  SNT_LAUNCH_CHECK("colsum_final_kernel");
  return SNT_OK;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
int add_vec(const float* a, const float* b, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return SNT_OK;
  add_vec_kernel<<<nblocks(n, 256), 256, 0, st>>>(a, b, out, n);
  SNT_LAUNCH_CHECK("add_vec_kernel");
  return SNT_OK;
}

// CHAIN: the cast is a link of a chain of launches (common.cuh) - the encoder head's operands in front of its contraction
template <bool CHAIN>
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  if (CHAIN) {
    pdl_launch_dependents();
    pdl_wait();
  }
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n && ((reinterpret_cast<uintptr_t>(src + i) & 15) == 0) &&
      ((reinterpret_cast<uintptr_t>(dst + i) & 7) == 0)) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 p;
    p.x = *reinterpret_cast<uint32_t*>(&lo);
    p.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + i) = p;
  } else {
    for (int k = 0; k < 4 && i + k < n; ++k) dst[i + k] = __float2bfloat16_rn(src[i + k]);
  }
}
int cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st, bool chain) {
  if (n <= 0) return SNT_OK;
  if (chain)
    SNT_CUDA(launch_chained(cast_bf16_kernel<true>, dim3(nblocks((n + 3) / 4, 256)), dim3(256), 0, st, src, dst, n));
  else
    cast_bf16_kernel<false><<<nblocks((n + 3) / 4, 256), 256, 0, st>>>(src, dst, n);
  SNT_LAUNCH_CHECK("cast_bf16_kernel");
  return SNT_OK;
}
__global__ void cast_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}
int cast_f32(const __nv_bfloat16* src, float* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return SNT_OK;
  cast_f32_kernel<<<nblocks(n, 256), 256, 0, st>>>(src, dst, n);
  SNT_LAUNCH_CHECK("cast_f32_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a7 pointwise: gate rows i|f|g|o (torch/nn/modules/rnn.py), c = f*c' + i*g, h = o*tanh(c)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_act(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename ActT>
__global__ void __launch_bounds__(256)
lstm_point_fwd_kernel(float* __restrict__ gates_t, const float* c_prev /* may alias cs_t */, float* cs_t,
                      ActT* __restrict__ hs_t, ActT* __restrict__ hprev_next, int bs, int bs_next, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)bs * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float* g = gates_t + (int64_t)b * 4 * H;
  const float i_ = sigmoidf_(g[j]);
  const float f_ = sigmoidf_(g[H + j]);
  const float g_ = tanhf(g[2 * H + j]);
  const float o_ = sigmoidf_(g[3 * H + j]);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float c = f_ * cp + i_ * g_;
  const float h = o_ * tanhf(c);
  g[j] = i_; g[H + j] = f_; g[2 * H + j] = g_; g[3 * H + j] = o_;
  cs_t[idx] = c;
  store_act(hs_t + idx, h);
  if (hprev_next != nullptr && b < bs_next) store_act(hprev_next + idx, h);
}
template <typename ActT>
int lstm_point_fwd(float* gates_t, const float* c_prev, float* cs_t, ActT* hs_t, ActT* hprev_next, int bs,
                   int bs_next, int64_t H, cudaStream_t st) {
  if (bs <= 0) return SNT_OK;
  lstm_point_fwd_kernel<ActT><<<nblocks((int64_t)bs * H, 256), 256, 0, st>>>(gates_t, c_prev, cs_t, hs_t,
                                                                            hprev_next, bs, bs_next, (int)H);
  SNT_LAUNCH_CHECK("lstm_point_fwd_kernel");
  return SNT_OK;
}
template int lstm_point_fwd<float>(float*, const float*, float*, float*, float*, int, int, int64_t, cudaStream_t);
template int lstm_point_fwd<__nv_bfloat16>(float*, const float*, float*, __nv_bfloat16*, __nv_bfloat16*, int,
                                           int, int64_t, cudaStream_t);

__global__ void __launch_bounds__(256)
lstm_point_bwd_kernel(float* __restrict__ gates_t, const float* __restrict__ cs_t,
                      const float* __restrict__ c_prev, const float* __restrict__ d_hs_t,
                      const float* __restrict__ dh_rec, float* __restrict__ dc_state, int bs, int bs_next, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)bs * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float* g = gates_t + (int64_t)b * 4 * H;
  const float i_ = g[j], f_ = g[H + j], g_ = g[2 * H + j], o_ = g[3 * H + j];
  const float tc = tanhf(cs_t[idx]);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const bool has_next = b < bs_next;
  const float dh = d_hs_t[idx] + (has_next ? dh_rec[idx] : 0.f);
  const float dc = (has_next ? dc_state[idx] : 0.f) + dh * o_ * (1.f - tc * tc);
  g[j] = dc * g_ * i_ * (1.f - i_);
  g[H + j] = dc * cp * f_ * (1.f - f_);
  g[2 * H + j] = dc * i_ * (1.f - g_ * g_);
  g[3 * H + j] = dh * tc * o_ * (1.f - o_);
  dc_state[idx] = dc * f_;
}
int lstm_point_bwd(float* gates_t, const float* cs_t, const float* c_prev, const float* d_hs_t,
                   const float* dh_rec, float* dc_state, int bs, int bs_next, int64_t H, cudaStream_t st) {
  if (bs <= 0) return SNT_OK;
  lstm_point_bwd_kernel<<<nblocks((int64_t)bs * H, 256), 256, 0, st>>>(gates_t, cs_t, c_prev, d_hs_t, dh_rec,
                                                                      dc_state, bs, bs_next, (int)H);
  SNT_LAUNCH_CHECK("lstm_point_bwd_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a9: cross-entropy over rows of a logits chunk (train.py:53,143).  One block per row.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, sh[i]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += sh[i];
  return r;
}

__global__ void __launch_bounds__(256)
ce_rows_fwd_kernel(const float* __restrict__ logits, int64_t V, int64_t ld, const int64_t* __restrict__ targets,
                   float* __restrict__ lse, float* __restrict__ nll, int* flags) {
  __shared__ float sh[8];
  const int64_t r = blockIdx.x;
  const float* row = logits + r * ld;
  float m = -INFINITY;
  for (int64_t v = threadIdx.x; v < V; v += 256) m = fmaxf(m, row[v]);
  m = block_reduce_max(m, sh);
  float s = 0.f;
  for (int64_t v = threadIdx.x; v < V; v += 256) s += expf(row[v] - m);
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) {
    const float l = m + logf(s);
    lse[r] = l;
    int64_t t = targets[r];
    if (t < 0 || t >= V) { atomicOr(flags, 2); nll[r] = 0.f; }
    else nll[r] = l - row[t];
  }
}
int ce_rows_fwd(const float* logits, int64_t R, int64_t V, int64_t ld, const int64_t* targets, float* lse,
                float* nll, cudaStream_t st) {
  if (R <= 0) return SNT_OK;
  ce_rows_fwd_kernel<<<(unsigned)R, 256, 0, st>>>(logits, V, ld, targets, lse, nll, device_flags());
  SNT_LAUNCH_CHECK("ce_rows_fwd_kernel");
  return SNT_OK;
}

__global__ void __launch_bounds__(256)
ce_rows_bwd_kernel(float* __restrict__ logits, int64_t V, int64_t ld, const int64_t* __restrict__ targets,
                   const float* __restrict__ lse, const float* __restrict__ dloss, float scale) {
  const int64_t r = blockIdx.y;
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= V) return;
  const float sc = scale * (dloss ? dloss[0] : 1.f);
  float* p = logits + r * ld + v;
  float d = expf(*p - lse[r]);
  if (v == targets[r]) d -= 1.f;
  *p = d * sc;
}
int ce_rows_bwd(float* logits, int64_t R, int64_t V, int64_t ld, const int64_t* targets, const float* lse,
                const float* dloss, float scale, cudaStream_t st) {
  if (R <= 0) return SNT_OK;
  SNT_REQUIRE(R <= 65535, "ce_rows_bwd: chunk too tall");
  dim3 grid(nblocks(V, 256), (unsigned)R);
  ce_rows_bwd_kernel<<<grid, 256, 0, st>>>(logits, V, ld, targets, lse, dloss, scale);
  SNT_LAUNCH_CHECK("ce_rows_bwd_kernel");
  return SNT_OK;
}

__global__ void __launch_bounds__(1024)
reduce_sum_kernel(const float* __restrict__ v, int64_t n, float scale, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += v[i];
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s * scale;
}
int reduce_sum(const float* v, int64_t n, float scale, float* out, cudaStream_t st) {
  reduce_sum_kernel<<<1, 1024, 0, st>>>(v, n, scale, out);
  SNT_LAUNCH_CHECK("reduce_sum_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a12 tail: first-index argmax over the vocab row + gather of the next input (models.py:63-65)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
argmax_gather_kernel(const float* __restrict__ logits, int64_t V, int64_t ld, const float* __restrict__ w_emb,
                     int E, int64_t* __restrict__ ids, int64_t ids_stride, float* __restrict__ x_next) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int best;
  const int64_t b = blockIdx.x;
  const float* row = logits + b * ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int64_t v = threadIdx.x; v < V; v += 256) {
    float x = row[v];
    if (x > m || mi == 0x7fffffff) { m = x; mi = (int)v; }  // strictly greater keeps the first index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = m; si[w] = mi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = sv[0];
    int bi = si[0];
    for (int i = 1; i < 8; ++i)
      if (sv[i] > bm || (sv[i] == bm && si[i] < bi)) { bm = sv[i]; bi = si[i]; }
    best = bi;
    ids[b * ids_stride] = bi;
  }
  __syncthreads();
  if (x_next != nullptr) {
    const float* src = w_emb + (int64_t)best * E;
    for (int e = threadIdx.x; e < E; e += 256) x_next[b * E + e] = src[e];
  }
}
int argmax_gather(const float* logits, int64_t B, int64_t V, int64_t ld, const float* w_emb, int64_t E,
                  int64_t* ids, int64_t ids_stride, float* x_next, cudaStream_t st) {
  if (B <= 0) return SNT_OK;
  argmax_gather_kernel<<<(unsigned)B, 256, 0, st>>>(logits, V, ld, w_emb, (int)E, ids, ids_stride, x_next);
  SNT_LAUNCH_CHECK("argmax_gather_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a2: BatchNorm1d over the batch (models.py:17,28; momentum 0.01).
// Block = 8 features x 128 row-lanes (E/8 blocks: 32 at E=256).  A warp covers 4 rows x 8 features = four whole 32-byte
// sectors per load; each thread keeps its first BN_CACHE rows in registers, so for B <= 1024 the batch is read from
// global memory once (statistics are still two-pass: mean first, then centred squares).  Column sums are reduced over
// the 128 row-lanes through shared memory in a fixed order.
// ---------------------------------------------------------------------------------------------------------
constexpr int BN_FEATS = 8, BN_LANES = 128, BN_CACHE = 8;
__device__ __forceinline__ float bn_block_sum(float v, float (*red)[BN_FEATS], int tx, int ty) {
  // a warp holds 4 rows x 8 features: fold the 4 rows with two shuffles, then the 32 warps through shared memory
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  __syncthreads();  // previous use of red[] is over
  if ((threadIdx.x & 31) < BN_FEATS) red[threadIdx.x >> 5][tx] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < BN_LANES / 4; ++i) t += red[i][tx];  // same order in every thread: deterministic
  return t;
}
__global__ void __launch_bounds__(BN_FEATS * BN_LANES)
bn_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
              float* __restrict__ running_mean, float* __restrict__ running_var, int training, float momentum,
              float eps, int B, int E, float* __restrict__ out, float* __restrict__ yhat,
              float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ out_bf16) {
  pdl_launch_dependents();
  pdl_wait();  // launched chained (common.cuh): nothing global before this point
  __shared__ float red[BN_LANES / 4][BN_FEATS];
  const int tx = threadIdx.x & (BN_FEATS - 1), ty = threadIdx.x / BN_FEATS;
  const int e = blockIdx.x * BN_FEATS + tx;
  const bool ok = e < E;
  float v[BN_CACHE];
#pragma unroll
  for (int k = 0; k < BN_CACHE; ++k) {
    const int r = ty + BN_LANES * k;
    v[k] = (ok && r < B) ? y[(int64_t)r * E + e] : 0.f;
  }
  float mu, rs;
  if (training) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < BN_CACHE; ++k) s += v[k];
    if (ok) for (int r = ty + BN_LANES * BN_CACHE; r < B; r += BN_LANES) s += y[(int64_t)r * E + e];
    mu = bn_block_sum(s, red, tx, ty) / (float)B;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < BN_CACHE; ++k) {
      const float d = v[k] - mu;
      if (ty + BN_LANES * k < B) q += d * d;
    }
    if (ok) for (int r = ty + BN_LANES * BN_CACHE; r < B; r += BN_LANES) { const float d = y[(int64_t)r * E + e] - mu; q += d * d; }
    const float t = bn_block_sum(q, red, tx, ty);
    const float var = t / (float)B;
    rs = rsqrtf(var + eps);
    if (ty == 0 && ok) {
      const float unbiased = B > 1 ? t / (float)(B - 1) : var;
      running_mean[e] = (1.f - momentum) * running_mean[e] + momentum * mu;
      running_var[e] = (1.f - momentum) * running_var[e] + momentum * unbiased;
    }
  } else {
    mu = ok ? running_mean[e] : 0.f;
    rs = ok ? rsqrtf(running_var[e] + eps) : 0.f;
  }
  if (!ok) return;
  const float ga = gamma[e], be = beta[e];
  if (ty == 0) rstd_out[e] = rs;
#pragma unroll
  for (int k = 0; k < BN_CACHE; ++k) {
    const int r = ty + BN_LANES * k;
    if (r < B) {
      const float yh = (v[k] - mu) * rs;
      yhat[(int64_t)r * E + e] = yh;
      const float o = yh * ga + be;
      out[(int64_t)r * E + e] = o;
      if (out_bf16) out_bf16[(int64_t)r * E + e] = __float2bfloat16_rn(o);
    }
  }
  for (int r = ty + BN_LANES * BN_CACHE; r < B; r += BN_LANES) {
    const float yh = (y[(int64_t)r * E + e] - mu) * rs;
    yhat[(int64_t)r * E + e] = yh;
    const float o = yh * ga + be;
    out[(int64_t)r * E + e] = o;
    if (out_bf16) out_bf16[(int64_t)r * E + e] = __float2bfloat16_rn(o);
  }
}
int bn_fwd(const float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
           int training, float momentum, float eps, int64_t B, int64_t E, float* out, float* yhat,
           float* rstd, cudaStream_t st, __nv_bfloat16* out_bf16) {
  SNT_CUDA(launch_chained(bn_fwd_kernel, dim3(nblocks(E, BN_FEATS)), dim3(BN_FEATS * BN_LANES), 0, st, y, gamma, beta,
                          running_mean, running_var, training, momentum, eps, (int)B, (int)E, out, yhat, rstd, out_bf16));
  SNT_LAUNCH_CHECK("bn_fwd_kernel");
  return SNT_OK;
}

__global__ void __launch_bounds__(BN_FEATS * BN_LANES)
bn_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ yhat, const float* __restrict__ rstd,
              const float* __restrict__ gamma, int training, int B, int E, float* __restrict__ dy,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[BN_LANES / 4][BN_FEATS];
  const int tx = threadIdx.x & (BN_FEATS - 1), ty = threadIdx.x / BN_FEATS;
  const int e = blockIdx.x * BN_FEATS + tx;
  const bool ok = e < E;
  float g[BN_CACHE], h[BN_CACHE];
#pragma unroll
  for (int k = 0; k < BN_CACHE; ++k) {
    const int r = ty + BN_LANES * k;
    const bool in = ok && r < B;
    g[k] = in ? dout[(int64_t)r * E + e] : 0.f;
    h[k] = in ? yhat[(int64_t)r * E + e] : 0.f;
  }
  float s = 0.f, d = 0.f;
#pragma unroll
  for (int k = 0; k < BN_CACHE; ++k) { s += g[k]; d += g[k] * h[k]; }
  if (ok)
    for (int r = ty + BN_LANES * BN_CACHE; r < B; r += BN_LANES) {
      const float gg = dout[(int64_t)r * E + e];
      s += gg;
      d += gg * yhat[(int64_t)r * E + e];
    }
  const float ssum = bn_block_sum(s, red, tx, ty);
  const float sdot = bn_block_sum(d, red, tx, ty);
  if (!ok) return;
  if (ty == 0) { dbeta[e] = ssum; dgamma[e] = sdot; }
  const float ga = gamma[e], rs = rstd[e];
  const float ms = ssum / (float)B, md = sdot / (float)B;
#pragma unroll
  for (int k = 0; k < BN_CACHE; ++k) {
    const int r = ty + BN_LANES * k;
    if (r < B) dy[(int64_t)r * E + e] = training ? ga * rs * (g[k] - ms - h[k] * md) : ga * rs * g[k];
  }
  for (int r = ty + BN_LANES * BN_CACHE; r < B; r += BN_LANES) {
    const int64_t i = (int64_t)r * E + e;
    const float gg = dout[i];
    dy[i] = training ? ga * rs * (gg - ms - yhat[i] * md) : ga * rs * gg;
  }
}
int bn_bwd(const float* dout, const float* yhat, const float* rstd, const float* gamma, int training,
           int64_t B, int64_t E, float* dy, float* dgamma, float* dbeta, cudaStream_t st) {
  bn_bwd_kernel<<<nblocks(E, BN_FEATS), BN_FEATS * BN_LANES, 0, st>>>(dout, yhat, rstd, gamma, training, (int)B,
                                                                     (int)E, dy, dgamma, dbeta);
  SNT_LAUNCH_CHECK("bn_bwd_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a10 (embedding part): dx rows t>=1 scatter-add into d_w_emb[V,E], deterministically.
// counting sort by token (stable rank = number of earlier packed rows with the same token), then one block
// per vocab row sums its rows in packed-row order.  Rows t==0 go to dfeatures.
// ---------------------------------------------------------------------------------------------------------
// Deterministic scatter-add of the token rows into d_w_emb[V,E], fully parallel:
//   emb_tok_kernel     token of every packed row t >= 1 and a histogram (integer atomics: order-independent)
//   emb_scan_kernel    exclusive scan of the histogram -> start[0..V]; tokens that occur more than once are appended
//                      to a work list
//   emb_place_kernel   a row whose token occurs once is copied straight to d_w_emb (warp per row); other rows claim a
//                      slot of their token's segment (arbitrary order)
//   emb_small_kernel   segments of 2..32 rows: one warp per token ranks the row indices with shuffles and adds the rows
//                      in ascending order
//   emb_sort/chunk/final  longer segments (see below)
__global__ void __launch_bounds__(256)
emb_tok_kernel(const __grid_constant__ PackInfo pk, const int64_t* __restrict__ captions, int64_t cap_stride,
               int64_t V, int* __restrict__ tok, int* __restrict__ count, int* flags) {
  const int n1 = pk.off[pk.T] - pk.off[1];  // rows with t >= 1
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n1) return;
  const int row = pk.off[1] + i;
  const int t = find_step(pk, row);
  const int b = row - pk.off[t];
  const int64_t tk = captions[(int64_t)b * cap_stride + (t - 1)];
  if (tk < 0 || tk >= V) { atomicOr(flags, 1); tok[i] = -1; return; }
  tok[i] = (int)tk;
  atomicAdd(&count[tk], 1);
}
// single block: warp-shuffle scan of 1024-wide slabs.  Tokens that occur more than once go to one of two work lists
// (warp-aggregated appends; list order is irrelevant, tokens are independent): small[0] / multi[0] = entries,
// small[1..] = tokens with 2..EMB_SMALL_MAX rows (one warp each), multi[1..] = longer segments (one block each).
constexpr int EMB_SMALL_MAX = 32;
__global__ void __launch_bounds__(1024)
emb_scan_kernel(const int* __restrict__ count, int V, int* __restrict__ start, int* __restrict__ cursor,
                int* __restrict__ multi, int* __restrict__ small, int* flags) {
  __shared__ int wsum[32];
  __shared__ int carry_s, n_small_s, n_multi_s;  // list lengths live in shared memory: a global atomic round trip per
                                                 // warp and slab (~2 us each) used to dominate this kernel
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_s = 0; n_small_s = 0; n_multi_s = 0; }
  int pre[16];  // the first 16 slabs' counts are fetched up front: one global-load latency instead of one per slab
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = k * 1024 + threadIdx.x;
    pre[k] = i < V ? count[i] : 0;
  }
  __syncthreads();
  int slab = 0;
#pragma unroll 1
  for (int base = 0; base < V; base += 1024, ++slab) {
    const int i = base + threadIdx.x;
    int c;
    if (slab < 16) {
      c = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) c = (k == slab) ? pre[k] : c;
    } else {
      c = i < V ? count[i] : 0;
    }
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int sv = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, sv, o);
        if (lane >= o) sv += y;
      }
      wsum[lane] = sv;  // inclusive scan of the warp totals
    }
    __syncthreads();
    const int excl = carry_s + (w > 0 ? wsum[w - 1] : 0) + x - c;
    if (i < V) {
      start[i] = excl;
      cursor[i] = excl;
    }
    const bool is_small = c > 1 && c <= EMB_SMALL_MAX, is_big = c > EMB_SMALL_MAX;
    const unsigned ms = __ballot_sync(0xffffffffu, is_small), mb = __ballot_sync(0xffffffffu, is_big);
    const unsigned lt = (1u << lane) - 1u;
    if (ms) {
      int b0 = 0;
      if (lane == 0) b0 = atomicAdd(&n_small_s, __popc(ms));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (is_small) small[1 + b0 + __popc(ms & lt)] = i;
    }
    if (mb) {
      int b0 = 0;
      if (lane == 0) b0 = atomicAdd(&n_multi_s, __popc(mb));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (is_big) multi[1 + b0 + __popc(mb & lt)] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) { start[V] = carry_s; small[0] = n_small_s; multi[0] = n_multi_s; }
}
__global__ void __launch_bounds__(256)
emb_place_kernel(const int* __restrict__ tok, int n1, const int* __restrict__ count, int* __restrict__ cursor,
                 int* __restrict__ perm0, const float* __restrict__ dx1, int E, float* __restrict__ d_w_emb) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per row
  const int lane = threadIdx.x & 31;
  if (i >= n1) return;
  const int t = tok[i];
  if (t < 0) return;
  if (count[t] == 1) {
    const float* src = dx1 + (int64_t)i * E;
    float* dst = d_w_emb + (int64_t)t * E;
    for (int e = lane * 4; e < E; e += 128) *reinterpret_cast<float4*>(dst + e) = *reinterpret_cast<const float4*>(src + e);
  } else if (lane == 0) {
    perm0[atomicAdd(&cursor[t], 1)] = i;
  }
}
// segments of 2..32 rows: ONE WARP per token, no block-wide synchronisation.  Lane i holds row index i of the segment,
// ranks it against the others with shuffles, and the warp then adds the rows in ascending row order (deterministic).
template <int SLABS>
__global__ void __launch_bounds__(256)
emb_small_kernel(const float* __restrict__ dx1, const int* __restrict__ start, const int* __restrict__ perm0,
                 const int* __restrict__ small, int E, float* __restrict__ d_w_emb) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  const int n_small = small[0];
  for (int li = gw; li < n_small; li += nw) {
    const int v = small[1 + li];
    const int s0 = start[v];
    const int n = start[v + 1] - s0;  // 2..32
    const int mine = lane < n ? perm0[s0 + lane] : 0x7fffffff;
    int rank = 0;
    for (int b = 0; b < n; ++b) rank += (__shfl_sync(0xffffffffu, mine, b) < mine);  // row indices are distinct
    float4 acc[SLABS];
#pragma unroll
    for (int k = 0; k < SLABS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    // lane r of `sorted` holds the r-th smallest row index
    int sorted = 0;
    for (int r = 0; r < n; ++r) {
      const unsigned who = __ballot_sync(0xffffffffu, lane < n && rank == r);
      const int idx = __shfl_sync(0xffffffffu, mine, __ffs((int)who) - 1);
      if (lane == r) sorted = idx;
    }
    for (int r0 = 0; r0 < n; r0 += 4) {  // 4 rows in flight (a dependent round of scattered loads costs ~2 us)
      float4 a[4][SLABS];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = __shfl_sync(0xffffffffu, sorted, (r0 + u) & 31);
        const float* src = dx1 + (int64_t)idx * E + lane * 4;
#pragma unroll
        for (int k = 0; k < SLABS; ++k)
          a[u][k] = (r0 + u < n && k * 128 + lane * 4 < E) ? *reinterpret_cast<const float4*>(src + k * 128)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)  // ascending row order
#pragma unroll
        for (int k = 0; k < SLABS; ++k) {
          acc[k].x += a[u][k].x; acc[k].y += a[u][k].y; acc[k].z += a[u][k].z; acc[k].w += a[u][k].w;
        }
    }
#pragma unroll
    for (int k = 0; k < SLABS; ++k)
      if (k * 128 + lane * 4 < E) *reinterpret_cast<float4*>(d_w_emb + (int64_t)v * E + k * 128 + lane * 4) = acc[k];
  }
}
// Longer segments, three short kernels so that one frequent token (e.g. <start>: one row per caption) is spread over
// the machine instead of over one block.  A dependent round of scattered global loads costs ~2 us here, so every kernel
// issues all of a thread's loads in one round:
//   emb_sort_kernel   persistent blocks walk the work list and rank the segment's row indices into ascending order with
//                     a bitmap over all packed rows in shared memory (set bits, prefix popcount: O(n + N/32)) -> gsorted;
//                     reserve ceil(n/8) chunk slots
//   emb_chunk_kernel  one warp per chunk of 8 consecutive sorted rows, all 8 rows in flight -> partial[slot][E]
//   emb_final_kernel  one block per token: 32 warps add contiguous runs of chunk partials, then the 32 warp sums are
//                     added in warp order -> d_w_emb[v]
// Every sum runs in a fixed order: deterministic.
constexpr int EMB_CHUNK = 8;
constexpr int EMB_BITMAP_WORDS = 8192;  // rows covered by the shared-memory bitmap: 262144
struct EmbChunkLists {
  int* counter;    // [1] chunk slots reserved so far
  int* chunk_pos;  // [slots] first position in gsorted
  int* chunk_cnt;  // [slots] rows in the chunk (1..8)
  int* tok_base;   // [n_multi] first slot of work-list entry li
};
__global__ void __launch_bounds__(1024)
emb_sort_kernel(const int* __restrict__ start, const int* __restrict__ perm0, int* __restrict__ gsorted,
                const int* __restrict__ multi, int n1, EmbChunkLists cl) {
  __shared__ unsigned bitmap[EMB_BITMAP_WORDS];
  __shared__ int wpre[EMB_BITMAP_WORDS / 8];  // exclusive prefix of popcounts, one entry per 8 words
  __shared__ int wsum[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int words = (n1 + 31) >> 5;
  const int n_multi = multi[0];
  for (int li = blockIdx.x; li < n_multi; li += gridDim.x) {
    const int v = multi[1 + li];
    const int s0 = start[v];
    const int n = start[v + 1] - s0;
    __syncthreads();  // previous token's shared data no longer in use
    if (words <= EMB_BITMAP_WORDS) {
      for (int i = threadIdx.x; i < words; i += 1024) bitmap[i] = 0u;
      __syncthreads();
      for (int a = threadIdx.x; a < n; a += 1024) {
        const int idx = perm0[s0 + a];
        atomicOr(&bitmap[idx >> 5], 1u << (idx & 31));
      }
      __syncthreads();
      // thread t owns words 8t..8t+7: block-wide exclusive scan of their popcounts
      int mine = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = threadIdx.x * 8 + k;
        mine += i < words ? __popc(bitmap[i]) : 0;
      }
      int x = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) wsum[w] = x;
      __syncthreads();
      if (w == 0) {
        int sv = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, sv, o);
          if (lane >= o) sv += y;
        }
        wsum[lane] = sv;
      }
      __syncthreads();
      wpre[threadIdx.x] = (w > 0 ? wsum[w - 1] : 0) + x - mine;
      __syncthreads();
      for (int a = threadIdx.x; a < n; a += 1024) {
        const int idx = perm0[s0 + a];
        const int wd = idx >> 5;
        int rank = wpre[wd >> 3];
        for (int k = wd & ~7; k < wd; ++k) rank += __popc(bitmap[k]);
        rank += __popc(bitmap[wd] & ((1u << (idx & 31)) - 1u));
        gsorted[s0 + rank] = idx;
      }
    } else {  // more packed rows than the bitmap covers: rank straight from global memory, O(n^2)
      for (int a = threadIdx.x; a < n; a += 1024) {
        const int mine = perm0[s0 + a];
        int rank = 0;
        for (int b = 0; b < n; ++b) rank += (perm0[s0 + b] < mine);
        gsorted[s0 + rank] = mine;
      }
    }
    const int nc = (n + EMB_CHUNK - 1) / EMB_CHUNK;
    if (threadIdx.x == 0) {
      s_base = atomicAdd(cl.counter, nc);  // slot order across tokens is irrelevant: tokens are independent
      cl.tok_base[li] = s_base;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nc; k += 1024) {
      cl.chunk_pos[s_base + k] = s0 + k * EMB_CHUNK;
      cl.chunk_cnt[s_base + k] = min(EMB_CHUNK, n - k * EMB_CHUNK);
    }
  }
}
template <int SLABS>
__global__ void __launch_bounds__(256)
emb_chunk_kernel(const float* __restrict__ dx1, const int* __restrict__ gsorted, EmbChunkLists cl, int E,
                 float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  const int n_chunks = cl.counter[0];
  for (int ci = gw; ci < n_chunks; ci += nw) {
    const int pos = cl.chunk_pos[ci], cnt = cl.chunk_cnt[ci];
    const int mine = lane < cnt ? gsorted[pos + lane] : 0;
    float4 a[EMB_CHUNK][SLABS];
#pragma unroll
    for (int u = 0; u < EMB_CHUNK; ++u) {  // all rows of the chunk in flight at once
      const int idx = __shfl_sync(0xffffffffu, mine, u);
      const float* src = dx1 + (int64_t)idx * E + lane * 4;
#pragma unroll
      for (int k = 0; k < SLABS; ++k)
        a[u][k] = (u < cnt && k * 128 + lane * 4 < E) ? *reinterpret_cast<const float4*>(src + k * 128)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 acc[SLABS];
#pragma unroll
    for (int k = 0; k < SLABS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < EMB_CHUNK; ++u)  // ascending row order
#pragma unroll
      for (int k = 0; k < SLABS; ++k) {
        acc[k].x += a[u][k].x; acc[k].y += a[u][k].y; acc[k].z += a[u][k].z; acc[k].w += a[u][k].w;
      }
#pragma unroll
    for (int k = 0; k < SLABS; ++k)
      if (k * 128 + lane * 4 < E) *reinterpret_cast<float4*>(partial + (int64_t)ci * E + k * 128 + lane * 4) = acc[k];
  }
}
template <int SLABS>
__global__ void __launch_bounds__(1024)
emb_final_kernel(const float* __restrict__ partial, const int* __restrict__ start, const int* __restrict__ multi,
                 EmbChunkLists cl, int E, float* __restrict__ d_w_emb) {
  extern __shared__ float part_s[];  // [32][E]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_multi = multi[0];
  for (int li = blockIdx.x; li < n_multi; li += gridDim.x) {
    const int v = multi[1 + li];
    const int nc = (start[v + 1] - start[v] + EMB_CHUNK - 1) / EMB_CHUNK;
    const int per = (nc + 31) / 32;  // warp w adds chunk partials [w*per, (w+1)*per) in slot order
    const int c0 = w * per, c1 = min(nc, c0 + per);
    const float* src = partial + (int64_t)cl.tok_base[li] * E + lane * 4;
    float4 acc[SLABS];
#pragma unroll
    for (int k = 0; k < SLABS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c < c1; c += 4) {
      float4 a[4][SLABS];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < SLABS; ++k)
          a[u][k] = (c + u < c1 && k * 128 + lane * 4 < E)
                        ? *reinterpret_cast<const float4*>(src + (int64_t)(c + u) * E + k * 128)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < SLABS; ++k) {
          acc[k].x += a[u][k].x; acc[k].y += a[u][k].y; acc[k].z += a[u][k].z; acc[k].w += a[u][k].w;
        }
    }
    __syncthreads();  // previous token's part_s no longer in use
#pragma unroll
    for (int k = 0; k < SLABS; ++k)
      if (k * 128 + lane * 4 < E) *reinterpret_cast<float4*>(part_s + w * E + k * 128 + lane * 4) = acc[k];
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += 1024) {
      float t = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) t += part_s[k * E + e];
      d_w_emb[(int64_t)v * E + e] = t;
    }
  }
}
__global__ void __launch_bounds__(256)
dfeatures_kernel(const float* __restrict__ dx, int bs0, int64_t B, int64_t E, float* __restrict__ dfeat) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= B * E) return;
  dfeat[i] = (i / E) < bs0 ? dx[i] : 0.f;
}

static int64_t emb_chunk_slots(int64_t N) { return N / EMB_CHUNK + N / (EMB_SMALL_MAX + 1) + 2; }  // sum ceil(n_i/8)
int64_t embed_bwd_ws_bytes(int64_t N, int64_t V) {
  // E is not known here: the chunk partials are sized for the largest supported row (E = 1024)
  return ws_bytes_for(N, 4) * 3 + ws_bytes_for(V + 1, 4) * 6 + ws_bytes_for(emb_chunk_slots(N), 4) * 2 +
         ws_bytes_for(emb_chunk_slots(N) * 1024, 4) + ws_bytes_for(4, 4);
}
// phase 0: everything.  phase 1: only the token-dependent half (histogram + scan: needs captions, not dx) - it may run
// early, on another stream; given d_w_emb (and E) it also zeroes the gradient buffer and the chunk counter there.
// phase 2: the rest, for a workspace that already went through phase 1 with the same captions / geometry.  phase 3:
// as 2, after a phase 1 that did the zeroing.  Phase 0 issues exactly the launches of phase 1 + phase 2.
int embed_pack_bwd(const PackInfo& pk, const float* dx, const int64_t* captions, int64_t cap_stride,
                   int64_t B, int64_t E, int64_t V, float* dfeatures, float* d_w_emb, void* ws,
                   int64_t ws_bytes, cudaStream_t st, int phase) {
  const int N = pk.off[pk.T];
  const int n1 = N - pk.off[1];
  if (phase != 1 && dfeatures) {
    dfeatures_kernel<<<nblocks(B * E, 256), 256, 0, st>>>(dx, pk.off[1], B, E, dfeatures);
    SNT_LAUNCH_CHECK("dfeatures_kernel");
  }
  if (phase != 1 && !d_w_emb) return SNT_OK;
  Workspace w(ws, ws_bytes);
  int* tok = w.take<int>(N);
  int* perm0 = w.take<int>(N);
  int* gsorted = w.take<int>(N);
  int* count = w.take<int>(V + 1);
  int* start = w.take<int>(V + 1);
  int* cursor = w.take<int>(V + 1);
  int* multi = w.take<int>(V + 1);
  int* small = w.take<int>(V + 1);
  EmbChunkLists cl;
  cl.tok_base = w.take<int>(V + 1);
  cl.chunk_pos = w.take<int>(emb_chunk_slots(N));
  cl.chunk_cnt = w.take<int>(emb_chunk_slots(N));
  float* partial = w.take<float>(emb_chunk_slots(N) * 1024);
  cl.counter = w.take<int>(4);
  if (!w.ok()) { set_error("embed_pack_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_REQUIRE((phase == 1 && !d_w_emb) || (E % 4 == 0 && E <= 1024), "embed_pack_bwd: E must be a multiple of 4 and <= 1024");
  SNT_REQUIRE(V < (1LL << 31), "embed_pack_bwd: V too large");
  const bool zero_here = phase == 0 || phase == 2 || (phase == 1 && d_w_emb != nullptr);
  if (zero_here) SNT_CUDA(cudaMemsetAsync(d_w_emb, 0, sizeof(float) * (size_t)V * E, st));
  if (n1 <= 0) return SNT_OK;
  if (phase < 2) SNT_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(V + 1), st));
  if (zero_here) SNT_CUDA(cudaMemsetAsync(cl.counter, 0, sizeof(int) * 4, st));
  if (phase < 2) {
    emb_tok_kernel<<<nblocks(n1, 256), 256, 0, st>>>(pk, captions, cap_stride, V, tok, count, device_flags());
    SNT_LAUNCH_CHECK("emb_tok_kernel");
    emb_scan_kernel<<<1, 1024, 0, st>>>(count, (int)V, start, cursor, multi, small, device_flags());
    SNT_LAUNCH_CHECK("emb_scan_kernel");
  }
  if (phase == 1) return SNT_OK;
  const float* dx1 = dx + (int64_t)pk.off[1] * E;
  emb_place_kernel<<<nblocks(n1, 8), 256, 0, st>>>(tok, n1, count, cursor, perm0, dx1, (int)E, d_w_emb);
  SNT_LAUNCH_CHECK("emb_place_kernel");
  const int slabs = (int)((E + 127) / 128);
  const int sgrid = 4 * 148;  // persistent warps walk the work lists
  if (slabs <= 2) emb_small_kernel<2><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  else if (slabs <= 4) emb_small_kernel<4><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  else emb_small_kernel<8><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  SNT_LAUNCH_CHECK("emb_small_kernel");
  emb_sort_kernel<<<148, 1024, 0, st>>>(start, perm0, gsorted, multi, n1, cl);
  SNT_LAUNCH_CHECK("emb_sort_kernel");
  if (slabs <= 2) emb_chunk_kernel<2><<<sgrid, 256, 0, st>>>(dx1, gsorted, cl, (int)E, partial);
  else if (slabs <= 4) emb_chunk_kernel<4><<<sgrid, 256, 0, st>>>(dx1, gsorted, cl, (int)E, partial);
  else emb_chunk_kernel<8><<<sgrid, 256, 0, st>>>(dx1, gsorted, cl, (int)E, partial);
  SNT_LAUNCH_CHECK("emb_chunk_kernel");
  {
    const size_t fsm = sizeof(float) * 32 * (size_t)E;  // <= 128 KB at E = 1024
    static bool attr_set = false;
    if (!attr_set) {
      const int max_sm = (int)(sizeof(float) * 32 * 1024);
      SNT_CUDA(cudaFuncSetAttribute(emb_final_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
      SNT_CUDA(cudaFuncSetAttribute(emb_final_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
      SNT_CUDA(cudaFuncSetAttribute(emb_final_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
      attr_set = true;
    }
    if (slabs <= 2) emb_final_kernel<2><<<148, 1024, fsm, st>>>(partial, start, multi, cl, (int)E, d_w_emb);
    else if (slabs <= 4) emb_final_kernel<4><<<148, 1024, fsm, st>>>(partial, start, multi, cl, (int)E, d_w_emb);
    else emb_final_kernel<8><<<148, 1024, fsm, st>>>(partial, start, multi, cl, (int)E, d_w_emb);
  }
  SNT_LAUNCH_CHECK("emb_final_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a11: clip_gradient (clamp) + Adam (train.py:88-91,145-146; torch.optim.Adam defaults)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float step_size, float beta1,
                                      float beta2, float omb1, float omb2, float eps, float rsqrt_bc2, float grad_clip,
                                      float grad_scale) {
  float gi = g * grad_scale;
  if (grad_clip > 0.f) gi = fminf(fmaxf(gi, -grad_clip), grad_clip);
  m = beta1 * m + omb1 * gi;          // exp_avg.lerp_(grad, 1 - beta1)
  v = beta2 * v + omb2 * gi * gi;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * rsqrt_bc2 + eps;    // sqrt(v) / sqrt(bias_correction2) + eps
  p = p - step_size * (m / denom);                   // step_size = lr / bias_correction1
}
__global__ void __launch_bounds__(256)
clamp_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, int64_t n, float step_size, float beta1, float beta2, float omb1,
                  float omb2, float eps, float rsqrt_bc2, float grad_clip, float grad_scale, int vec) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  if (vec && i + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
           vv = *reinterpret_cast<float4*>(v + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    adam1(pp.x, gg.x, mm.x, vv.x, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.y, gg.y, mm.y, vv.y, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.z, gg.z, mm.z, vv.z, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.w, gg.w, mm.w, vv.w, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (int k = 0; k < 4 && i + k < n; ++k)
      adam1(p[i + k], g[i + k], m[i + k], v[i + k], step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip,
            grad_scale);
  }
}
// multi-tensor variant: one launch updates up to ADAM_MAX_TENSORS parameter tensors (pointer table by value)
constexpr int ADAM_MAX_TENSORS = 24;
struct AdamTable {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long n[ADAM_MAX_TENSORS];
  int block0[ADAM_MAX_TENSORS + 1];  // first block of each tensor (1024 elements per block)
  int count;
};
__global__ void __launch_bounds__(256)
clamp_adam_multi_kernel(const __grid_constant__ AdamTable tb, float step_size, float beta1, float beta2, float omb1,
                        float omb2, float eps, float rsqrt_bc2, float grad_clip, float grad_scale) {
  int k = 0;
  while (k + 1 < tb.count && (int)blockIdx.x >= tb.block0[k + 1]) ++k;
  const long long n = tb.n[k];
  const long long i = ((long long)(blockIdx.x - tb.block0[k]) * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  float* p = tb.p[k];
  const float* g = tb.g[k];
  float* m = tb.m[k];
  float* v = tb.v[k];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec && i + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
           vv = *reinterpret_cast<float4*>(v + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    adam1(pp.x, gg.x, mm.x, vv.x, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.y, gg.y, mm.y, vv.y, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.z, gg.z, mm.z, vv.z, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    adam1(pp.w, gg.w, mm.w, vv.w, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (int q = 0; q < 4 && i + q < n; ++q)
      adam1(p[i + q], g[i + q], m[i + q], v[i + q], step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip,
            grad_scale);
  }
}
int clamp_adam_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v,
                     const int64_t* n, double lr, double beta1, double beta2, double eps, float grad_clip,
                     float grad_scale, int64_t step, cudaStream_t st) {
  SNT_REQUIRE(step >= 1, "clamp_adam_multi: step must be >= 1");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  int k = 0;  // next tensor to place: carried across launches (empty tensors are skipped without using a table slot)
  while (k < count) {
    AdamTable tb;
    tb.count = 0;
    int blocks = 0;
    for (; k < count && tb.count < ADAM_MAX_TENSORS; ++k) {
      if (n[k] <= 0) continue;
      SNT_REQUIRE(p[k] && g[k] && m[k] && v[k], "clamp_adam_multi: NULL tensor %d", k);
      const int c = tb.count++;
      tb.p[c] = p[k]; tb.g[c] = g[k]; tb.m[c] = m[k]; tb.v[c] = v[k]; tb.n[c] = n[k];
      tb.block0[c] = blocks;
      blocks += (int)((n[k] + 1023) / 1024);
    }
    if (tb.count == 0) continue;
    tb.block0[tb.count] = blocks;
    clamp_adam_multi_kernel<<<blocks, 256, 0, st>>>(tb, (float)(lr / bc1), (float)beta1, (float)beta2,
                                                   (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                   (float)(1.0 / sqrt(bc2)), grad_clip, grad_scale);
    SNT_LAUNCH_CHECK("clamp_adam_multi_kernel");
  }
  return SNT_OK;
}

int clamp_adam(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
               double eps, float grad_clip, float grad_scale, int64_t step, cudaStream_t st) {
  if (n <= 0) return SNT_OK;
  SNT_REQUIRE(step >= 1, "clamp_adam: step must be >= 1");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const int vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  clamp_adam_kernel<<<nblocks((n + 3) / 4, 256), 256, 0, st>>>(p, g, m, v, n, (float)(lr / bc1), (float)beta1,
                                                              (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                              (float)eps, (float)(1.0 / sqrt(bc2)), grad_clip,
                                                              grad_scale, vec);
  SNT_LAUNCH_CHECK("clamp_adam_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Data-parallel exchange fused with the optimizer, over NVLink / NVSwitch multicast memory (one process per GPU):
// the flat gradient and parameter buffers of every rank are mapped behind one MULTICAST address each.  Every rank owns
// one contiguous shard of the flat index range and, for it alone,
//   g   = multimem.ld_reduce.add  [mc_g + i]       the switch sums the ranks' gradients in flight: the all-reduce
//   p,m,v <- clamp + Adam                           1/world of the optimizer's work and traffic per GPU
//   multimem.st [mc_p + i] <- p                     the new parameters land in every rank's buffer: the broadcast
// so the gradient all-reduce, the optimizer and the parameter broadcast are ONE kernel whose cost shrinks with the
// number of GPUs, nothing of the exchange runs beside (and competes with) the backward pass, and replicas stay
// bit-identical by construction (each value is reduced and updated exactly once).  The caller brackets the launch with
// two cross-rank barriers: all gradients written before, all parameters delivered after.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
constexpr int DP_UNROLL = 4;  // float4s per thread: all their multicast reads are in flight before the first is used
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dp_adam_shard_kernel(const float* __restrict__ mc_g, float* __restrict__ mc_p, const float* __restrict__ p,
                     float* __restrict__ m, float* __restrict__ v, int64_t lo, int64_t hi, float step_size, float beta1,
                     float beta2, float omb1, float omb2, float eps, float rsqrt_bc2, float grad_clip, float grad_scale) {
  // a block covers THREADS * DP_UNROLL consecutive float4s per round; thread t takes float4s t, t + THREADS, ... of them
  // (coalesced); the grid walks the shard with a stride (a narrow grid serves a bucket that is exchanged beside the
  // cooperative recurrence kernel: it may never occupy more SMs than that kernel leaves free)
  constexpr int64_t PER = (int64_t)THREADS * DP_UNROLL * 4;
  for (int64_t b0 = lo + (int64_t)blockIdx.x * PER; b0 < hi; b0 += (int64_t)gridDim.x * PER) {
    const int64_t base = b0 + (int64_t)threadIdx.x * 4;
    float4 gg[DP_UNROLL], pp[DP_UNROLL], mm[DP_UNROLL], vv[DP_UNROLL];
#pragma unroll
    for (int k = 0; k < DP_UNROLL; ++k) {
      const int64_t i = base + (int64_t)k * THREADS * 4;
      if (i < hi) gg[k] = multimem_ld_reduce_add(mc_g + i);
    }
#pragma unroll
    for (int k = 0; k < DP_UNROLL; ++k) {
      const int64_t i = base + (int64_t)k * THREADS * 4;
      if (i < hi) {
        pp[k] = *reinterpret_cast<const float4*>(p + i);
        mm[k] = *reinterpret_cast<const float4*>(m + i);
        vv[k] = *reinterpret_cast<const float4*>(v + i);
      }
    }
#pragma unroll
    for (int k = 0; k < DP_UNROLL; ++k) {
      const int64_t i = base + (int64_t)k * THREADS * 4;
      if (i >= hi) continue;  // lo, hi multiples of 4: whole float4s
      adam1(pp[k].x, gg[k].x, mm[k].x, vv[k].x, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
      adam1(pp[k].y, gg[k].y, mm[k].y, vv[k].y, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
      adam1(pp[k].z, gg[k].z, mm[k].z, vv[k].z, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
      adam1(pp[k].w, gg[k].w, mm[k].w, vv[k].w, step_size, beta1, beta2, omb1, omb2, eps, rsqrt_bc2, grad_clip, grad_scale);
      *reinterpret_cast<float4*>(m + i) = mm[k];
      *reinterpret_cast<float4*>(v + i) = vv[k];
      multimem_st(mc_p + i, pp[k]);
    }
  }
}
int dp_adam_shard(const float* mc_g, float* mc_p, const float* p, float* m, float* v, int64_t lo, int64_t hi, double lr,
                  double beta1, double beta2, double eps, float grad_clip, float grad_scale, int64_t step, int max_blocks,
                  cudaStream_t st) {
  if (hi <= lo) return SNT_OK;
  SNT_REQUIRE(mc_g && mc_p && p && m && v, "dp_adam_shard: NULL pointer");
  SNT_REQUIRE(step >= 1 && lo >= 0 && (lo & 3) == 0 && (hi & 3) == 0, "dp_adam_shard: shard bounds must be multiples of 4");
  SNT_REQUIRE(((reinterpret_cast<uintptr_t>(mc_g) | reinterpret_cast<uintptr_t>(mc_p) | reinterpret_cast<uintptr_t>(p) |
                reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
              "dp_adam_shard: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float a0 = (float)(lr / bc1), a1 = (float)beta1, a2 = (float)beta2, a3 = (float)(1.0 - beta1),
              a4 = (float)(1.0 - beta2), a5 = (float)eps, a6 = (float)(1.0 / sqrt(bc2));
  // max_blocks > 0: a bounded grid that walks the shard with a stride.  The blocks stay small (256 threads) so that they
  // fit into whatever thread slots and registers free up on an SM next to other background work - a 1024-thread block
  // needs a whole empty SM and starved behind the output-bias column sums for the length of the BPTT (measured).
  unsigned grid = nblocks((hi - lo) / 4, 256 * DP_UNROLL);
  if (max_blocks > 0 && grid > (unsigned)max_blocks) grid = (unsigned)max_blocks;
  dp_adam_shard_kernel<256><<<grid, 256, 0, st>>>(mc_g, mc_p, p, m, v, lo, hi, a0, a1, a2, a3, a4, a5, a6, grad_clip,
                                                  grad_scale);
  SNT_LAUNCH_CHECK("dp_adam_shard_kernel");
  return SNT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// f4: the caller-side tail of sample() (eval.py:101-109): a caption is the words before the first <end>.
// One warp per caption; a row of `steps` ids is scanned 32 at a time, the first hit ends the caption.
// HBM-bound (8 or 16 bytes per token), latency-trivial: it exists so that eval needs ONE device->host copy of
// (ids, lengths) instead of a Python loop over every word of every caption.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
caption_trim_kernel(const int64_t* __restrict__ ids, int64_t B, int steps, int64_t end_id, int64_t pad_id,
                    int32_t* __restrict__ lengths, int64_t* ids_out) {
  const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;  // whole warps leave together: the ballots below always see a full warp
  const int lane = threadIdx.x & 31;
  const int64_t* row = ids + b * steps;
  int len = steps;
  for (int s0 = 0; s0 < steps; s0 += 32) {
    const int s = s0 + lane;
    const bool hit = s < steps && row[s] == end_id;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m != 0u) { len = s0 + __ffs((int)m) - 1; break; }  // warp-uniform
  }
  if (lane == 0 && lengths != nullptr) lengths[b] = len;
  if (ids_out != nullptr) {
    int64_t* orow = ids_out + b * steps;
    // in-place use (ids_out == ids) is safe: every lane has finished reading the row at the last ballot
    for (int s = lane; s < steps; s += 32) {
      const int64_t v = s < len ? row[s] : pad_id;
      orow[s] = v;
    }
  }
}
// steps <= 32 (the reference decodes 20): one warp per 32 captions, read as ONE flat run of 32 * steps ids - every load
// and store instruction covers 256 contiguous bytes with all lanes active (one warp per caption used 20 of 32 lanes on
// 160-byte rows: 0.38 of the HBM peak at 4 M captions).  The ids stay in registers between the scan and the masked
// write; the first <end> of each caption is a shared-memory atomicMin (a handful of hits per warp).
template <int MAXS>
__global__ void __launch_bounds__(256)
caption_trim_flat_kernel(const int64_t* __restrict__ ids, int64_t B, int steps, int64_t end_id, int64_t pad_id,
                         int32_t* __restrict__ lengths, int64_t* ids_out) {
  __shared__ int slen[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 32;
  if (row0 >= B) return;
  const int nrows = (int)min((int64_t)32, B - row0);
  const int total = nrows * steps;
  slen[warp][lane] = steps;
  __syncwarp();
  const int64_t* base = ids + row0 * steps;
  int64_t v[MAXS];
#pragma unroll
  for (int i = 0; i < MAXS; ++i) {  // all loads first (an atomic between them would serialise the loads)
    const int e = i * 32 + lane;
    v[i] = (i < steps && e < total) ? __ldcs(base + e) : pad_id;
  }
#pragma unroll
  for (int i = 0; i < MAXS; ++i) {
    const int e = i * 32 + lane;
    if (i < steps && e < total && v[i] == end_id) atomicMin(&slen[warp][e / steps], e % steps);
  }
  __syncwarp();
  if (lengths != nullptr && lane < nrows) lengths[row0 + lane] = slen[warp][lane];
  if (ids_out != nullptr) {
    int64_t* obase = ids_out + row0 * steps;  // in-place use is safe: the warp has read its whole run by now
#pragma unroll
    for (int i = 0; i < MAXS; ++i) {
      const int e = i * 32 + lane;
      if (i < steps && e < total) __stcs(obase + e, (e % steps) < slen[warp][e / steps] ? v[i] : pad_id);
    }
  }
}
int caption_trim(const int64_t* ids, int64_t B, int steps, int64_t end_id, int64_t pad_id, int32_t* lengths,
                 int64_t* ids_out, cudaStream_t st) {
  if (B <= 0 || steps <= 0) return SNT_OK;
  if (steps <= 20) {  // the reference's decode length: fewer registers, three blocks per SM in flight
    caption_trim_flat_kernel<20><<<(unsigned)((B + 255) / 256), 256, 0, st>>>(ids, B, steps, end_id, pad_id, lengths, ids_out);
    SNT_LAUNCH_CHECK("caption_trim_flat_kernel");
    return SNT_OK;
  }
  if (steps <= 32) {
    caption_trim_flat_kernel<32><<<(unsigned)((B + 255) / 256), 256, 0, st>>>(ids, B, steps, end_id, pad_id, lengths, ids_out);
    SNT_LAUNCH_CHECK("caption_trim_flat_kernel");
    return SNT_OK;
  }
  caption_trim_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(ids, B, steps, end_id, pad_id, lengths, ids_out);
  SNT_LAUNCH_CHECK("caption_trim_kernel");
  return SNT_OK;
}

}  // namespace snt
