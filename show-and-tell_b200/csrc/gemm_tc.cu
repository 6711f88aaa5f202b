// Host side of the tcgen05 contraction core: tensor-map encoding (driver entry point resolved at run time, so
// the library has no link-time dependency on libcuda and still loads on a machine without a GPU), tile-shape
// and split-K selection, the deterministic split-K reduction and the exported snt_gemm_bf16.
#include "gemm_tc.cuh"

#include <atomic>
#include <stdlib.h>

#include <cudaTypedefs.h>
#include <mutex>

namespace snt {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
  }
  return n;
}

static std::atomic<int> g_sm_reserve{0};
static thread_local int g_grid_cap = 0;  // > 0: upper bound on the grids this thread launches (set_grid_cap)
int grid_sms() {
  int n = sm_count() - g_sm_reserve.load();
  if (g_grid_cap > 0 && n > g_grid_cap) n = g_grid_cap;
  return n < 1 ? 1 : n;
}
int set_grid_cap(int n) {
  const int old = g_grid_cap;
  g_grid_cap = n < 0 ? 0 : n;
  return old;
}
int set_sm_reserve(int n) { return g_sm_reserve.exchange(n < 0 ? 0 : n); }

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                   int box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return SNT_ECUDA; }
  SNT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  SNT_REQUIRE(ld % 8 == 0, "tensor map: leading dimension %lld not a multiple of 8 bf16", (long long)ld);
  SNT_REQUIRE(inner >= 1 && outer >= 1 && ld >= inner, "tensor map: bad extents");
  SNT_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256, "tensor map: bad box");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (inner %lld outer %lld ld %lld box %dx%d)", (int)r,
              (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
    return SNT_ECUDA;
  }
  return SNT_OK;
}

int make_operand_tmap(CUtensorMap* out, const void* base, bool mn_major, int64_t rows, int64_t K, int64_t ld,
                      int box_rows) {
  if (!mn_major) return make_tmap_bf16(out, base, K, rows, ld, BK, box_rows);
  return make_tmap_bf16(out, base, rows, K, ld, 64, BK);
}

// ---- split-K reduction: C = beta*C + bias + sum_s ws[s] -----------------------------------------------------------
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t M, int64_t N, int64_t ldc,
                     int64_t split_stride, float beta, const float* __restrict__ bias, float* __restrict__ C,
                     __nv_bfloat16* __restrict__ Cb) {
  pdl_launch_dependents();
  pdl_wait();  // the partial slices come from the preceding contraction
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= M * N) return;
  const int64_t m = i / N, n = i % N;
  const int64_t o = m * ldc + n;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += ws[(int64_t)k * split_stride + o];
  if (bias) s += bias[n];
  if (C) {
    if (beta != 0.f) s += beta * C[o];
    C[o] = s;
  }
  if (Cb) Cb[o] = __float2bfloat16_rn(s);
}
// same, four columns per thread (N, ldc, split_stride multiples of 4, 16-byte aligned bases, fp32 output only); the
// partial slices are read once and never again: streaming loads
__global__ void __launch_bounds__(256)
splitk_reduce4_kernel(const float* __restrict__ ws, int splits, int64_t M, int64_t N4, int64_t ldc,
                      int64_t split_stride, float beta, const float* __restrict__ bias, float* __restrict__ C) {
  pdl_launch_dependents();
  pdl_wait();  // the partial slices come from the preceding contraction
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= M * N4) return;
  const int64_t m = i / N4, n = (i - m * N4) * 4;
  const int64_t o = m * ldc + n;
  float4 s = __ldcs(reinterpret_cast<const float4*>(ws + o));
  for (int k = 1; k < splits; ++k) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + (int64_t)k * split_stride + o));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
  }
  if (beta != 0.f) {
    const float4 c = *reinterpret_cast<const float4*>(C + o);
    s.x += beta * c.x; s.y += beta * c.y; s.z += beta * c.z; s.w += beta * c.w;
  }
  *reinterpret_cast<float4*>(C + o) = s;
}

int64_t gemm_tc_split_ws_elems(int64_t M, int64_t ldc, int splits) { return splits > 1 ? splits * M * ldc : 0; }

static int pick_bn(int64_t M, int64_t N) {
  const int64_t num_m = (M + BM - 1) / BM;
  const int sms = sm_count();
  if (N >= 256 && num_m * ((N + 255) / 256) >= sms) return 256;
  if (N >= 128 && num_m * ((N + 127) / 128) >= sms / 2) return 128;
  return N > 64 ? 128 : 64;
}

int choose_splits(int64_t M, int64_t N, int64_t K, int bn) {
  if (bn <= 0) bn = pick_bn(M, N);
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const int64_t kblocks = (K + BK - 1) / BK;
  const int sms = sm_count();
  if (tiles >= sms / 2 || kblocks < 8) return 1;
  int64_t s = sms / tiles;
  if (s > kblocks / 4) s = kblocks / 4;
  if (s > 16) s = 16;
  return s < 1 ? 1 : (int)s;
}

template <int BN, bool A_MN, bool B_MN>
static int run_plain(const CUtensorMap& ta, const CUtensorMap& tb, const TileSched& ts, int64_t M, int64_t N,
                     float alpha, float beta, float* C, __nv_bfloat16* Cb, int64_t ldc, const float* bias,
                     int64_t split_stride, cudaStream_t st, int perm, const float* adev, const float* rs) {
  if (BN == 256 && ts.kblocks_per_split <= 8 && !getenv("SNT_NO_WIDE_EPI")) {  // short K: epilogue-bound
    PlainEpi<256, true> e;
    e.M = (int)M; e.N = (int)N; e.alpha = alpha; e.beta = beta; e.C = C; e.Cb = Cb; e.ldc = ldc; e.bias = bias;
    e.split_stride = split_stride; e.row_perm_h = perm; e.alpha_dev = adev; e.row_scale = rs;
    return launch_gemm_tc<256, A_MN, B_MN, PlainEpi<256, true>>(ta, tb, ts, e, st);
  }
  PlainEpi<BN> e;
  e.M = (int)M; e.N = (int)N; e.alpha = alpha; e.beta = beta; e.C = C; e.Cb = Cb; e.ldc = ldc; e.bias = bias;
  e.split_stride = split_stride; e.row_perm_h = perm; e.alpha_dev = adev; e.row_scale = rs;
  return launch_gemm_tc<BN, A_MN, B_MN, PlainEpi<BN>>(ta, tb, ts, e, st);
}

template <int BN>
static int run_plain_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TileSched& ts,
                           int64_t M, int64_t N, float alpha, float beta, float* C, __nv_bfloat16* Cb, int64_t ldc,
                           const float* bias, int64_t split_stride, cudaStream_t st, int perm, const float* adev, const float* rs) {
  if (!a_mn && !b_mn) return run_plain<BN, false, false>(ta, tb, ts, M, N, alpha, beta, C, Cb, ldc, bias, split_stride, st, perm, adev, rs);
  if (!a_mn && b_mn) return run_plain<BN, false, true>(ta, tb, ts, M, N, alpha, beta, C, Cb, ldc, bias, split_stride, st, perm, adev, rs);
  if (a_mn && !b_mn) return run_plain<BN, true, false>(ta, tb, ts, M, N, alpha, beta, C, Cb, ldc, bias, split_stride, st, perm, adev, rs);
  return run_plain<BN, true, true>(ta, tb, ts, M, N, alpha, beta, C, Cb, ldc, bias, split_stride, st, perm, adev, rs);
}

int gemm_tc(bool a_mn, bool b_mn, int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A,
            int64_t lda, const __nv_bfloat16* B, int64_t ldb, float beta, float* C, __nv_bfloat16* Cb,
            int64_t ldc, const float* bias, int splits, float* split_ws, cudaStream_t st, int row_perm_h,
            const float* alpha_dev, bool keep_partials, int* splits_used, int force_bn, int n_fastest, const float* row_scale) {
  if (M <= 0 || N <= 0) return SNT_OK;
  SNT_REQUIRE(row_perm_h == 0 || M == 4 * (int64_t)row_perm_h, "gemm_tc: row permutation needs M == 4H");
  SNT_REQUIRE(K >= 1 && A && B && (C || Cb), "gemm_tc: bad arguments");
  SNT_REQUIRE(M < (1 << 30) && N < (1 << 30) && K < (1 << 30), "gemm_tc: extent too large");
  const int bn = (force_bn == 64 || force_bn == 128 || force_bn == 256) ? force_bn : pick_bn(M, N);
  TileSched ts;
  ts.num_m = (int)((M + BM - 1) / BM);
  ts.num_n = (int)((N + bn - 1) / bn);
  ts.kblocks = (int)((K + BK - 1) / BK);
  if (splits < 1) splits = 1;
  if (splits > ts.kblocks) splits = ts.kblocks;
  ts.kblocks_per_split = (ts.kblocks + splits - 1) / splits;
  splits = (ts.kblocks + ts.kblocks_per_split - 1) / ts.kblocks_per_split;  // no empty split
  ts.splits = splits;
  ts.a_row0 = 0;
  ts.b_row0 = 0;
  ts.n_fastest = n_fastest;
  CUtensorMap ta, tb;
  SNT_CHECK(make_operand_tmap(&ta, A, a_mn, M, K, lda, BM));
  SNT_CHECK(make_operand_tmap(&tb, B, b_mn, N, K, ldb, bn));

  float* out = C;
  __nv_bfloat16* outb = Cb;
  float e_alpha = alpha, e_beta = beta;
  const float* e_bias = bias;
  int64_t split_stride = 0;
  if (splits_used) *splits_used = splits;
  if (splits > 1 || keep_partials) {
    SNT_REQUIRE(split_ws != nullptr, "gemm_tc: split-K needs a workspace");
    out = split_ws; outb = nullptr; e_beta = 0.f; e_bias = nullptr;
    split_stride = M * ldc;
  }
  int rc;
  if (bn == 256) rc = run_plain_major<256>(a_mn, b_mn, ta, tb, ts, M, N, e_alpha, e_beta, out, outb, ldc, e_bias, split_stride, st, row_perm_h, alpha_dev, row_scale);
  else if (bn == 128) rc = run_plain_major<128>(a_mn, b_mn, ta, tb, ts, M, N, e_alpha, e_beta, out, outb, ldc, e_bias, split_stride, st, row_perm_h, alpha_dev, row_scale);
  else rc = run_plain_major<64>(a_mn, b_mn, ta, tb, ts, M, N, e_alpha, e_beta, out, outb, ldc, e_bias, split_stride, st, row_perm_h, alpha_dev, row_scale);
  SNT_CHECK(rc);
  if (splits > 1 && !keep_partials) {
    const int64_t total = M * N;
    const bool vec = C && !Cb && N % 4 == 0 && ldc % 4 == 0 && split_stride % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(split_ws) |
                       reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
    if (vec) {
      SNT_CUDA(launch_chained(splitk_reduce4_kernel, dim3((unsigned)((total / 4 + 255) / 256)), dim3(256), 0, st,
                              (const float*)split_ws, splits, M, N / 4, ldc, split_stride, beta, bias, C));
      SNT_LAUNCH_CHECK("splitk_reduce4_kernel");
    } else {
      SNT_CUDA(launch_chained(splitk_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st,
                              (const float*)split_ws, splits, M, N, ldc, split_stride, beta, bias, C, Cb));
      SNT_LAUNCH_CHECK("splitk_reduce_kernel");
    }
  }
  return SNT_OK;
}

// Wave-balanced variant for long contractions whose tile count is not a multiple of the SM count.  The row tiles are
// cut in two groups: the first fills whole waves of unsplit tiles (written straight to C); the remaining row tiles - less
// than one wave - are split along K just enough to occupy every SM once more, and reduced by the (small) split-K
// reduction.  Instead of 2 rounds of full-length tiles (158 tiles on 148 SMs) the job takes 1 round + 1/14 round.
// `split_ws` must hold 16 * tail_rows * ldc floats with tail_rows < ceil(sms / num_n) * 128 + 128.
int gemm_tc_balanced(bool a_mn, bool b_mn, int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A,
                     int64_t lda, const __nv_bfloat16* B, int64_t ldb, float* C, int64_t ldc, float* split_ws,
                     int64_t split_ws_elems, cudaStream_t st, const float* alpha_dev, int bn, const float* row_scale) {
  SNT_REQUIRE(bn == 128 || bn == 256, "gemm_tc_balanced: bn must be 128 or 256");
  const int64_t num_m = (M + BM - 1) / BM, num_n = (N + bn - 1) / bn, kb = (K + BK - 1) / BK;
  const int64_t sms = grid_sms();
  const int64_t tiles = num_m * num_n;
  const int64_t rounds = tiles / sms;
  int64_t m_a = rounds > 0 ? (rounds * sms) / num_n : 0;  // row tiles of the unsplit part
  if (m_a > num_m) m_a = num_m;
  const int64_t tail_m = num_m - m_a;
  if (tail_m == 0 || tiles % sms == 0 || split_ws == nullptr)
    return gemm_tc(a_mn, b_mn, M, N, K, alpha, A, lda, B, ldb, 0.f, C, nullptr, ldc, nullptr, 1, nullptr, st, 0, alpha_dev,
                   false, nullptr, bn, 1, row_scale);
  const int64_t tail_rows = M - m_a * BM, tail_tiles = tail_m * num_n;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 16 && s <= kb; ++s) {
    if ((int64_t)s * tail_rows * ldc > split_ws_elems) break;
    const int64_t per = (kb + s - 1) / s;
    const int64_t units = tail_tiles * ((kb + per - 1) / per);
    const double cost = (double)((units + sms - 1) / sms) * (double)per + (s > 1 ? 2.0 + 0.5 * s : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  if (m_a > 0)
    SNT_CHECK(gemm_tc(a_mn, b_mn, m_a * BM, N, K, alpha, A, lda, B, ldb, 0.f, C, nullptr, ldc, nullptr, 1, nullptr, st, 0,
                      alpha_dev, false, nullptr, bn, 1, row_scale));
  const int64_t r0 = m_a * BM;
  const __nv_bfloat16* A2 = a_mn ? A + r0 : A + r0 * lda;
  return gemm_tc(a_mn, b_mn, tail_rows, N, K, alpha, A2, lda, B, ldb, 0.f, C + r0 * ldc, nullptr, ldc, nullptr, best,
                 split_ws, st, 0, alpha_dev, false, nullptr, bn, 1, row_scale ? row_scale + r0 : nullptr);
}

}  // namespace tc
}  // namespace snt

extern "C" int snt_set_sm_reserve(int n) { return snt::tc::set_sm_reserve(n); }

// C[M,N] = alpha*op(A).op(B) + beta*C + bias.  transA=0: A [M,K]; 1: A [K,M].  transB=0: B [K,N]; 1: B [N,K].
extern "C" int snt_gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                             int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc,
                             int c_is_bf16, const float* bias, void* stream) {
  using namespace snt;
  return tc::gemm_tc(transA != 0, transB == 0, M, N, K, alpha, (const __nv_bfloat16*)A, lda,
                     (const __nv_bfloat16*)B, ldb, c_is_bf16 ? 0.f : beta, c_is_bf16 ? nullptr : (float*)C,
                     c_is_bf16 ? (__nv_bfloat16*)C : nullptr, ldc, bias, 1, nullptr, (cudaStream_t)stream);
}
