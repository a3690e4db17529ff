// Error channel + version for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace halo {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  return HALO_ERR_CUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}
}  // namespace halo

extern "C" int halo_abi_version(void) { return HALO_ABI_VERSION; }
extern "C" const char* halo_last_error(void) { return halo::g_err; }
