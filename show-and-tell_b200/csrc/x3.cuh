// Split-bf16 ("x3") contractions for the fp32-faithful mode on the tensor cores (gemm_x3.cu).
#pragma once
#include "common.cuh"

namespace snt {
namespace x3 {

// an fp32 matrix expanded into its six bf16 contraction slots
struct Operand {
  void* p = nullptr;    // bf16 [extent, K6] (mn = false: contraction index contiguous) or [K6, ld] (mn = true)
  int64_t ld = 0;       // row pitch in elements (multiple of 8)
  int64_t K6 = 0;       // expanded contraction length (six slots per chunk of the contraction index, zero padded)
  int chunks = 1;       // chunks of at most 2048 contraction values; > 1: run one K slice per chunk (see gemm_x3.cu)
  int64_t extent = 0;   // M (for an A operand) or N (for a B operand)
  bool mn = false;
};

bool enabled();                                    // false under SNT_FP32_FFMA=1
bool worth_it(int64_t M, int64_t N, int64_t K);    // large enough for the tensor-core path to pay
// bf16 elements an expanded operand needs: src[rows, cols], contraction index = cols (k_contig) or rows
int64_t operand_elems(int64_t rows, int64_t cols, bool k_contig);
// is_b selects the slot pattern (A: lo hi mid mid hi hi;  B: hi lo mid hi mid hi - small partial products first)
int expand(const float* src, int64_t rows, int64_t cols, int64_t ld, bool k_contig, bool is_b, void* dst, Operand* out,
           cudaStream_t st);
// C[M,N] = alpha * A.B + beta * C + bias over expanded operands; split_ws (optional) enables a K split for small outputs
int gemm(const Operand& a, const Operand& b, int64_t M, int64_t N, float alpha, float beta, float* C, int64_t ldc,
         const float* bias, float* split_ws, int64_t split_ws_elems, cudaStream_t st);
int scratch_alloc(void** p, int64_t bytes, cudaStream_t st);
int scratch_free(void* p, cudaStream_t st);
// drop-in for gemm_f32 (same argument meaning): expands both operands into stream-ordered scratch, contracts, frees
int gemm_f32_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, cudaStream_t st);

}  // namespace x3
}  // namespace snt
