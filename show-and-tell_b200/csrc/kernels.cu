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
                      __nv_bfloat16* __restrict__ x_bf16, int* flags) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int N = pk.off[pk.T];
  if (row >= N) return;
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
                   cudaStream_t st) {
  const int N = pk.off[pk.T];
  embed_pack_fwd_kernel<<<nblocks(N, 8), 256, 0, st>>>(pk, features, w_emb, captions, cap_stride, (int)E, V,
                                                      x_f32, x_bf16, device_flags());
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

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
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
int cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return SNT_OK;
  cast_bf16_kernel<<<nblocks((n + 3) / 4, 256), 256, 0, st>>>(src, dst, n);
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
// a2: BatchNorm1d over the batch (models.py:17,28; momentum 0.01).  Block = 32 features x 32 row-lanes.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
bn_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
              float* __restrict__ running_mean, float* __restrict__ running_var, int training, float momentum,
              float eps, int B, int E, float* __restrict__ out, float* __restrict__ yhat,
              float* __restrict__ rstd_out) {
  __shared__ float red[32][33];
  __shared__ float s_mean[32], s_rstd[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + tx;
  const bool ok = e < E;
  if (training) {
    float s = 0.f;
    if (ok) for (int r = ty; r < B; r += 32) s += y[(int64_t)r * E + e];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
      for (int i = 0; i < 32; ++i) t += red[i][tx];
      s_mean[tx] = t / (float)B;
    }
    __syncthreads();
    const float mu = s_mean[tx];
    float q = 0.f;
    if (ok) for (int r = ty; r < B; r += 32) { float d = y[(int64_t)r * E + e] - mu; q += d * d; }
    __syncthreads();
    red[ty][tx] = q;
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
      for (int i = 0; i < 32; ++i) t += red[i][tx];
      const float var = t / (float)B;
      s_rstd[tx] = rsqrtf(var + eps);
      if (ok) {
        const float unbiased = B > 1 ? t / (float)(B - 1) : var;
        running_mean[e] = (1.f - momentum) * running_mean[e] + momentum * mu;
        running_var[e] = (1.f - momentum) * running_var[e] + momentum * unbiased;
      }
    }
    __syncthreads();
  } else {
    if (ty == 0 && ok) {
      s_mean[tx] = running_mean[e];
      s_rstd[tx] = rsqrtf(running_var[e] + eps);
    }
    __syncthreads();
  }
  if (!ok) return;
  const float mu = s_mean[tx], rs = s_rstd[tx], ga = gamma[e], be = beta[e];
  if (ty == 0) rstd_out[e] = rs;
  for (int r = ty; r < B; r += 32) {
    const float yh = (y[(int64_t)r * E + e] - mu) * rs;
    yhat[(int64_t)r * E + e] = yh;
    out[(int64_t)r * E + e] = yh * ga + be;
  }
}
int bn_fwd(const float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
           int training, float momentum, float eps, int64_t B, int64_t E, float* out, float* yhat,
           float* rstd, cudaStream_t st) {
  bn_fwd_kernel<<<nblocks(E, 32), 1024, 0, st>>>(y, gamma, beta, running_mean, running_var, training, momentum,
                                                 eps, (int)B, (int)E, out, yhat, rstd);
  SNT_LAUNCH_CHECK("bn_fwd_kernel");
  return SNT_OK;
}

__global__ void __launch_bounds__(1024)
bn_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ yhat, const float* __restrict__ rstd,
              const float* __restrict__ gamma, int training, int B, int E, float* __restrict__ dy,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red1[32][33], red2[32][33];
  __shared__ float s_sum[32], s_dot[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + tx;
  const bool ok = e < E;
  float s = 0.f, d = 0.f;
  if (ok)
    for (int r = ty; r < B; r += 32) {
      const float g = dout[(int64_t)r * E + e];
      s += g;
      d += g * yhat[(int64_t)r * E + e];
    }
  red1[ty][tx] = s;
  red2[ty][tx] = d;
  __syncthreads();
  if (ty == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 32; ++i) { a += red1[i][tx]; c += red2[i][tx]; }
    s_sum[tx] = a;
    s_dot[tx] = c;
    if (ok) { dbeta[e] = a; dgamma[e] = c; }
  }
  __syncthreads();
  if (!ok) return;
  const float ga = gamma[e], rs = rstd[e];
  const float ms = s_sum[tx] / (float)B, md = s_dot[tx] / (float)B;
  for (int r = ty; r < B; r += 32) {
    const int64_t i = (int64_t)r * E + e;
    const float g = dout[i];
    dy[i] = training ? ga * rs * (g - ms - yhat[i] * md) : ga * rs * g;
  }
}
int bn_bwd(const float* dout, const float* yhat, const float* rstd, const float* gamma, int training,
           int64_t B, int64_t E, float* dy, float* dgamma, float* dbeta, cudaStream_t st) {
  bn_bwd_kernel<<<nblocks(E, 32), 1024, 0, st>>>(dout, yhat, rstd, gamma, training, (int)B, (int)E, dy, dgamma,
                                                 dbeta);
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
//   emb_multi_kernel   longer segments: persistent blocks walk the work list, the segment's row indices are ranked into
//                      ascending order in shared memory, then 32 warps sum the rows in that fixed order
constexpr int EMB_SEG_MAX = 8192;  // longest segment (rows sharing one token, e.g. <start>: one per caption)
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
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_s = 0; multi[0] = 0; small[0] = 0; }
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
      if (lane == 0) b0 = atomicAdd(&small[0], __popc(ms));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (is_small) small[1 + b0 + __popc(ms & lt)] = i;
    }
    if (mb) {
      int b0 = 0;
      if (lane == 0) b0 = atomicAdd(&multi[0], __popc(mb));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (is_big) multi[1 + b0 + __popc(mb & lt)] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) start[V] = carry_s;
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
    for (int r = 0; r < n; ++r) {
      const unsigned who = __ballot_sync(0xffffffffu, lane < n && rank == r);
      const int idx = __shfl_sync(0xffffffffu, mine, __ffs((int)who) - 1);
      const float* src = dx1 + (int64_t)idx * E + lane * 4;
#pragma unroll
      for (int k = 0; k < SLABS; ++k)
        if (k * 128 + lane * 4 < E) {
          const float4 a = *reinterpret_cast<const float4*>(src + k * 128);
          acc[k].x += a.x; acc[k].y += a.y; acc[k].z += a.z; acc[k].w += a.w;
        }
    }
#pragma unroll
    for (int k = 0; k < SLABS; ++k)
      if (k * 128 + lane * 4 < E) *reinterpret_cast<float4*>(d_w_emb + (int64_t)v * E + k * 128 + lane * 4) = acc[k];
  }
}
template <int SLABS>
__global__ void __launch_bounds__(1024)
emb_multi_kernel(const float* __restrict__ dx1, const int* __restrict__ start, const int* __restrict__ perm0,
                 int* __restrict__ gsorted, const int* __restrict__ multi, int E, float* __restrict__ d_w_emb) {
  extern __shared__ int sm_i[];                  // raw[EMB_SEG_MAX] | sorted[EMB_SEG_MAX] | part[32][E]
  int* raw = sm_i;
  int* sorted_s = sm_i + EMB_SEG_MAX;
  float* part = reinterpret_cast<float*>(sm_i + 2 * EMB_SEG_MAX);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_multi = multi[0];
  for (int li = blockIdx.x; li < n_multi; li += gridDim.x) {
    const int v = multi[1 + li];
    const int s0 = start[v];
    const int n = start[v + 1] - s0;
    const bool big = n > EMB_SEG_MAX;  // rare: rank straight from global memory into a global scratch segment
    const int* rawp = big ? perm0 + s0 : raw;
    int* sorted_w = big ? gsorted + s0 : sorted_s;
    __syncthreads();  // previous token's shared data no longer in use
    if (!big) {
      const int n4 = (n + 3) & ~3;
      for (int a = threadIdx.x; a < n4; a += 1024) raw[a] = a < n ? perm0[s0 + a] : 0x7fffffff;
      __syncthreads();
      // O(n^2) ranking out of shared memory, 4 indices per (broadcast) 16-byte load
      for (int a = threadIdx.x; a < n; a += 1024) {
        const int mine = raw[a];
        int rank = 0;
        const int4* r4 = reinterpret_cast<const int4*>(raw);
#pragma unroll 4
        for (int b = 0; b < n4 / 4; ++b) {
          const int4 v = r4[b];
          rank += (v.x < mine) + (v.y < mine) + (v.z < mine) + (v.w < mine);
        }
        sorted_s[rank] = mine;
      }
    } else {
      for (int a = threadIdx.x; a < n; a += 1024) {
        const int mine = rawp[a];
        int rank = 0;
        for (int b = 0; b < n; ++b) rank += (rawp[b] < mine);  // row indices are distinct; broadcast reads
        sorted_w[rank] = mine;
      }
    }
    __syncthreads();
    const int* sorted = sorted_w;
    // warp w sums rows w, w+32, ... (ascending): every lane owns SLABS float4 of the row, 2 rows in flight
    float4 acc[SLABS];
#pragma unroll
    for (int k = 0; k < SLABS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = w;
    for (; i + 96 < n; i += 128) {  // 4 rows in flight per warp, added in ascending order
      const float* rp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rp[u] = dx1 + (int64_t)sorted[i + 32 * u] * E + lane * 4;
      float4 a[4][SLABS];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < SLABS; ++k)
          a[u][k] = (k * 128 + lane * 4 < E) ? *reinterpret_cast<const float4*>(rp[u] + k * 128)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < SLABS; ++k) {
          acc[k].x += a[u][k].x; acc[k].y += a[u][k].y; acc[k].z += a[u][k].z; acc[k].w += a[u][k].w;
        }
    }
    for (; i < n; i += 32) {
      const float* r0 = dx1 + (int64_t)sorted[i] * E + lane * 4;
#pragma unroll
      for (int k = 0; k < SLABS; ++k)
        if (k * 128 + lane * 4 < E) {
          const float4 a = *reinterpret_cast<const float4*>(r0 + k * 128);
          acc[k].x += a.x; acc[k].y += a.y; acc[k].z += a.z; acc[k].w += a.w;
        }
    }
#pragma unroll
    for (int k = 0; k < SLABS; ++k)
      if (k * 128 + lane * 4 < E) *reinterpret_cast<float4*>(part + w * E + k * 128 + lane * 4) = acc[k];
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += 1024) {
      float t = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) t += part[k * E + e];
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

int64_t embed_bwd_ws_bytes(int64_t N, int64_t V) {
  return ws_bytes_for(N, 4) * 3 + ws_bytes_for(V + 1, 4) * 5;
}
int embed_pack_bwd(const PackInfo& pk, const float* dx, const int64_t* captions, int64_t cap_stride,
                   int64_t B, int64_t E, int64_t V, float* dfeatures, float* d_w_emb, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
  const int N = pk.off[pk.T];
  const int n1 = N - pk.off[1];
  if (dfeatures) {
    dfeatures_kernel<<<nblocks(B * E, 256), 256, 0, st>>>(dx, pk.off[1], B, E, dfeatures);
    SNT_LAUNCH_CHECK("dfeatures_kernel");
  }
  if (!d_w_emb) return SNT_OK;
  Workspace w(ws, ws_bytes);
  int* tok = w.take<int>(N);
  int* perm0 = w.take<int>(N);
  int* gsorted = w.take<int>(N);
  int* count = w.take<int>(V + 1);
  int* start = w.take<int>(V + 1);
  int* cursor = w.take<int>(V + 1);
  int* multi = w.take<int>(V + 1);
  int* small = w.take<int>(V + 1);
  if (!w.ok()) { set_error("embed_pack_bwd: workspace too small"); return SNT_EWORKSPACE; }
  SNT_REQUIRE(E % 4 == 0 && E <= 1024, "embed_pack_bwd: E must be a multiple of 4 and <= 1024");
  SNT_REQUIRE(V < (1LL << 31), "embed_pack_bwd: V too large");
  SNT_CUDA(cudaMemsetAsync(d_w_emb, 0, sizeof(float) * (size_t)V * E, st));
  if (n1 <= 0) return SNT_OK;
  SNT_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(V + 1), st));
  const float* dx1 = dx + (int64_t)pk.off[1] * E;
  emb_tok_kernel<<<nblocks(n1, 256), 256, 0, st>>>(pk, captions, cap_stride, V, tok, count, device_flags());
  SNT_LAUNCH_CHECK("emb_tok_kernel");
  emb_scan_kernel<<<1, 1024, 0, st>>>(count, (int)V, start, cursor, multi, small, device_flags());
  SNT_LAUNCH_CHECK("emb_scan_kernel");
  emb_place_kernel<<<nblocks(n1, 8), 256, 0, st>>>(tok, n1, count, cursor, perm0, dx1, (int)E, d_w_emb);
  SNT_LAUNCH_CHECK("emb_place_kernel");
  const size_t sm = sizeof(int) * 2 * EMB_SEG_MAX + sizeof(float) * 32 * (size_t)E;
  const int max_sm = (int)(sizeof(int) * 2 * EMB_SEG_MAX + sizeof(float) * 32 * 1024);
  static bool attr_set = false;
  if (!attr_set) {
    SNT_CUDA(cudaFuncSetAttribute(emb_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
    SNT_CUDA(cudaFuncSetAttribute(emb_multi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
    SNT_CUDA(cudaFuncSetAttribute(emb_multi_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sm));
    attr_set = true;
  }
  const int slabs = (int)((E + 127) / 128);
  const int sgrid = 4 * 148;  // persistent warps walk the small-segment list
  if (slabs <= 2) emb_small_kernel<2><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  else if (slabs <= 4) emb_small_kernel<4><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  else emb_small_kernel<8><<<sgrid, 256, 0, st>>>(dx1, start, perm0, small, (int)E, d_w_emb);
  SNT_LAUNCH_CHECK("emb_small_kernel");
  if (slabs <= 2) emb_multi_kernel<2><<<148, 1024, sm, st>>>(dx1, start, perm0, gsorted, multi, (int)E, d_w_emb);
  else if (slabs <= 4) emb_multi_kernel<4><<<148, 1024, sm, st>>>(dx1, start, perm0, gsorted, multi, (int)E, d_w_emb);
  else emb_multi_kernel<8><<<148, 1024, sm, st>>>(dx1, start, perm0, gsorted, multi, (int)E, d_w_emb);
  SNT_LAUNCH_CHECK("emb_multi_kernel");
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
  for (int base = 0; base < count; base += ADAM_MAX_TENSORS) {
    AdamTable tb;
    tb.count = 0;
    int blocks = 0;
    for (int k = base; k < count && tb.count < ADAM_MAX_TENSORS; ++k) {
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

}  // namespace snt
