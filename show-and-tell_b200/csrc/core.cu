// Error plumbing, packed-sequence geometry and the small utility exports of the C ABI.
#include <stdlib.h>
#include "kernels.cuh"

#include <atomic>
#include <mutex>

namespace snt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SNT_OK;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SNT_ECUDA;
}

int make_pack(const int32_t* batch_sizes, int T, PackInfo* out) {
  SNT_REQUIRE(batch_sizes != nullptr, "batch_sizes is NULL");
  SNT_REQUIRE(T >= 1 && T <= SNT_MAX_T, "T=%d outside [1,%d]", T, SNT_MAX_T);
  out->T = T;
  int64_t acc = 0;
  for (int t = 0; t < T; ++t) {
    SNT_REQUIRE(batch_sizes[t] >= 1, "batch_sizes[%d]=%d must be >= 1", t, batch_sizes[t]);
    SNT_REQUIRE(t == 0 || batch_sizes[t] <= batch_sizes[t - 1],
                "batch_sizes must be non-increasing (lengths sorted descending, data_loader.py:50)");
    out->off[t] = (int)acc;
    acc += batch_sizes[t];
    SNT_REQUIRE(acc < (int64_t)1 << 30, "packed sequence too long");
  }
  out->off[T] = (int)acc;
  for (int t = T + 1; t <= SNT_MAX_T; ++t) out->off[t] = (int)acc;
  return SNT_OK;
}

// one sticky flag word per device, allocated on first use and never freed
int* device_flags() {
  static int* flags[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (flags[dev] == nullptr) {
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(int));
    flags[dev] = p;
  }
  return flags[dev];
}

SideStream* side_stream(int which) {
  static SideStream tab[3][64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || which < 0 || which > 2) return nullptr;
  SideStream& x = tab[which][dev];
  if (!x.s) {
    if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.aux, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.aux2, cudaEventDisableTiming);
  }
  return &x;
}

static thread_local int g_pdl_suppress = 0;
bool pdl_all() {
  static const bool on = [] { const char* e = getenv("SNT_NO_PDL"); return !(e && e[0] == '1'); }();
  return on && g_pdl_suppress == 0;
}
PdlSuppress::PdlSuppress(bool active) : active_(active) { if (active_) ++g_pdl_suppress; }
PdlSuppress::~PdlSuppress() { if (active_) --g_pdl_suppress; }

}  // namespace snt

using namespace snt;

extern "C" int snt_abi_version(void) { return SNT_ABI_VERSION; }
extern "C" int64_t snt_launch_count(int reset) {
  return reset ? (int64_t)g_launches.exchange(0) : (int64_t)g_launches.load();
}
extern "C" const char* snt_last_error(void) { return g_err; }

extern "C" int snt_device_query(int device, int* sm_count, int* cc, int64_t* smem_optin) {
  cudaDeviceProp p;
  SNT_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc) *cc = p.major * 10 + p.minor;
  if (smem_optin) *smem_optin = (int64_t)p.sharedMemPerBlockOptin;
  return SNT_OK;
}

extern "C" int snt_read_flags(int* flags_out, int reset, void* stream) {
  int* f = device_flags();
  SNT_REQUIRE(f != nullptr, "no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  int h = 0;
  SNT_CUDA(cudaMemcpyAsync(&h, f, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (reset) SNT_CUDA(cudaMemsetAsync(f, 0, sizeof(int), st));
  SNT_CUDA(cudaStreamSynchronize(st));
  if (flags_out) *flags_out = h;
  return SNT_OK;
}
