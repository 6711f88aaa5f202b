// fp32 SIMT contraction for the fp32-faithful mode (SNT_PREC_FP32): fp32 operands, FFMA accumulation,
// no operand rounding.  This is the arithmetic the reference gets from cuBLAS sgemm / MKL sgemm behind
// nn.Linear / nn.LSTM (models.py:16,52,53); only the summation order differs.
// Tile 128x128x16, 256 threads, 8x8 outputs per thread, register-prefetch double buffering.
#include "common.cuh"
#include "x3.cuh"

namespace snt {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NTHREADS = 256;

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(NTHREADS)
gemm_f32_kernel(int64_t M, int64_t N, int64_t K, float alpha, const float* __restrict__ A, int64_t lda,
                const float* __restrict__ B, int64_t ldb, float beta, float* __restrict__ C, int64_t ldc,
                const float* __restrict__ bias) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // element (m,k) of A: A_KCONTIG ? A[m*lda + k] : A[k*lda + m];  (k,n) of B: B_KCONTIG ? B[n*ldb + k] : B[k*ldb + n]
  float ra[8], rb[8];
  auto load_tile = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, mm;
      if (A_KCONTIG) { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
      else           { mm = tid & 127; kk = (tid >> 7) + 2 * i; }
      int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) v = A_KCONTIG ? A[m * lda + k] : A[k * lda + m];
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, nn;
      if (B_KCONTIG) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
      else           { nn = tid & 127; kk = (tid >> 7) + 2 * i; }
      int64_t n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < K) v = B_KCONTIG ? B[n * ldb + k] : B[k * ldb + n];
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, mm;
      if (A_KCONTIG) { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
      else           { mm = tid & 127; kk = (tid >> 7) + 2 * i; }
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, nn;
      if (B_KCONTIG) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
      else           { nn = tid & 127; kk = (tid >> 7) + 2 * i; }
      Bs[buf][kk][nn] = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int64_t ktiles = (K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int64_t kt = 0; kt < ktiles; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < ktiles) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tile(buf ^ 1);  // the other buffer was last read in iteration kt-1, fenced by the sync below
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (beta != 0.f) v += beta * C[m * ldc + n];
      C[m * ldc + n] = v;
    }
  }
}

int gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
             const float* bias, cudaStream_t st) {
  if (M <= 0 || N <= 0) return SNT_OK;
  SNT_REQUIRE(K >= 0 && A && B && C, "gemm_f32: bad arguments");
  // large contractions run on the tensor cores as six bf16 partial products per fp32 product (gemm_x3.cu); small ones,
  // and everything under SNT_FP32_FFMA=1, on the FFMA kernel below
  if (x3::worth_it(M, N, K)) return x3::gemm_f32_tc(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, st);
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
  SNT_REQUIRE(grid.y <= 65535, "gemm_f32: M too large");
  const bool a_k = (transA == 0), b_k = (transB != 0);
  if (a_k && b_k)
    gemm_f32_kernel<true, true><<<grid, NTHREADS, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (a_k && !b_k)
    gemm_f32_kernel<true, false><<<grid, NTHREADS, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (!a_k && b_k)
    gemm_f32_kernel<false, true><<<grid, NTHREADS, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else
    gemm_f32_kernel<false, false><<<grid, NTHREADS, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  SNT_LAUNCH_CHECK("gemm_f32_kernel");
  return SNT_OK;
}

}  // namespace snt

extern "C" int snt_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
                            const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                            float* C, int64_t ldc, const float* bias, void* stream) {
  return snt::gemm_f32(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias,
                       (cudaStream_t)stream);
}
