// Library plumbing: thread-local error text, device info, launch counter.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace tcavp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static thread_local const char* g_last_kernel = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_arg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return TCAVP_ERR_ARG;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_last_kernel = what;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return TCAVP_ERR_CUDA;
  }
  return TCAVP_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// One warp spins for `ns` nanoseconds next to whatever else is resident on its SM and records (SM cycles, nanoseconds):
// the true average SM clock while another kernel runs (NVML's sampled clock hides fast power-cap modulation).
__global__ void clock_probe_kernel(unsigned long long* out, unsigned long long ns) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t1, c0, c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  c0 = clock64();
  do {
    __nanosleep(2000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
  c1 = clock64();
  out[0] = c1 - c0;
  out[1] = t1 - t0;
}

}  // namespace tcavp

extern "C" {

int tcavp_clock_probe(unsigned long long* out2, unsigned long long ns, tcavp_stream_t stream) {
  TCAVP_REQUIRE(out2 != nullptr, "tcavp_clock_probe: null out");
  tcavp::clock_probe_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out2, ns);
  return tcavp::check_launch("clock_probe_kernel");
}

const char* tcavp_last_error(void) { return tcavp::g_err; }
int tcavp_version(void) { return 100; }
long long tcavp_launch_count(void) { return tcavp::g_launches.load(); }
const char* tcavp_last_kernel(void) { return tcavp::g_last_kernel; }

int tcavp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TCAVP_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  TCAVP_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  TCAVP_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  TCAVP_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return TCAVP_OK;
}
}
