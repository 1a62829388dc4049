// Library plumbing: thread-local error text, device info, launch counter.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace tcavp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

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

}  // namespace tcavp

extern "C" {

const char* tcavp_last_error(void) { return tcavp::g_err; }
int tcavp_version(void) { return 100; }
long long tcavp_launch_count(void) { return tcavp::g_launches.load(); }

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
