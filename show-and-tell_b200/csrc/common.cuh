// Shared helpers for the snt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/snt_b200.h"

namespace snt {

// Packed-sequence geometry, passed to kernels BY VALUE (no H2D copy, no sync): T <= SNT_MAX_T.
struct PackInfo {
  int T;
  int off[SNT_MAX_T + 1];  // off[t] = first packed row of timestep t; off[T] = N
};

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
// Validates batch_sizes (positive, non-increasing, T in range) and fills PackInfo.
int make_pack(const int32_t* batch_sizes, int T, PackInfo* out);
int* device_flags();  // sticky device-side flag word

// Side stream (one per device) for small bandwidth-bound kernels - column sums of a gradient matrix - that read the same
// data as a tensor-core contraction but need almost no SM resources: they run next to it instead of after it.  Usage:
// record `fork` on the main stream, make `s` wait for it, launch, record `join` on `s`, make the main stream wait.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, aux = nullptr, aux2 = nullptr;
};
// which = 0: the stream the stage functions use for their column sums; which = 1: the step executor's stream for whole
// stages that run next to the main stream (head backward, embedding-gradient plan); which = 2: the executor's stream for
// the dW_out contraction beside the BPTT recurrence.  NULL if it could not be created.
SideStream* side_stream(int which = 0);
void count_launch();   // bumps the process-wide kernel-launch counter (snt_launch_count)

#define SNT_CHECK(expr)                                  \
  do {                                                   \
    int _e = (expr);                                     \
    if (_e != SNT_OK) return _e;                         \
  } while (0)
#define SNT_CUDA(expr) SNT_CHECK(::snt::check_cuda((expr), #expr))
#define SNT_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    ::snt::count_launch();                                       \
    SNT_CHECK(::snt::check_cuda(cudaGetLastError(), name));      \
  } while (0)
#define SNT_REQUIRE(cond, ...)                           \
  do {                                                   \
    if (!(cond)) {                                       \
      ::snt::set_error(__VA_ARGS__);                     \
      return SNT_EINVAL;                                 \
    }                                                    \
  } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Workspace {
  char* base;
  int64_t size, used;
  Workspace(void* p, int64_t n) : base((char*)p), size(n), used(0) {}
  template <typename T>
  T* take(int64_t count) {
    int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
    if (base == nullptr || used + bytes > size) {
      used = size + 1;
      return nullptr;
    }
    T* r = (T*)(base + used);
    used += bytes;
    return r;
  }
  bool ok() const { return used <= size; }
};
static inline int64_t ws_bytes_for(int64_t count, int64_t elem) { return align_up(count * elem, 256); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- chained launches (programmatic dependent launch) ------------------------------------------------------------------
// A kernel launched with launch_chained() may become resident while its predecessor in the stream is still running (as
// soon as every block of the predecessor has executed pdl_launch_dependents() or exited), so that its launch latency and
// prologue overlap the predecessor's tail.  The rule every such kernel follows: NO global-memory access before
// pdl_wait(), which returns when the preceding grid has completed and its writes are visible.  In a kernel launched the
// ordinary way both instructions are no-ops.  SNT_NO_PDL=1 launches everything in plain stream order.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_all();
// Plain stream order for the launches of this thread while the object lives: for a chain of full-GPU contractions that
// has bandwidth-bound work running BESIDE it on another stream - chained contractions take over every SM the moment
// their predecessor leaves it, and the blocks of the side work only get in when the whole chain has drained (measured
// at configs[3]: the output-bias column sums beside the vocabulary contractions, +2.5 % step time).
struct PdlSuppress {
  explicit PdlSuppress(bool active);
  ~PdlSuppress();
  PdlSuppress(const PdlSuppress&) = delete;
  PdlSuppress& operator=(const PdlSuppress&) = delete;
 private:
  bool active_;
};
template <class... KArgs, class... Args>
inline cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_all() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- internal launchers shared across translation units -------------------------------------------------
int gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
             const float* bias, cudaStream_t st);

}  // namespace snt
