// Error plumbing and device queries of the libvfr C ABI.
#include "vfr_common.cuh"
#include <stdarg.h>
#include <atomic>

namespace vfr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};   // kernels this library has launched (every launch goes through check_launch)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return VFR_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return VFR_OK;
}

}  // namespace vfr

extern "C" int64_t vfr_launch_count(void) { return (int64_t)vfr::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* vfr_last_error(void) { return vfr::g_err; }

extern "C" int vfr_version(void) { return 100; }

extern "C" int vfr_device_sms(void) {
  int dev = 0, sms = 0;
  VFR_CUDA(cudaGetDevice(&dev));
  VFR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return sms;
}
