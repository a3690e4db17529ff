// Error channel + version for the C ABI.
#include <stdarg.h>
#include <stdlib.h>
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

static thread_local int g_path = 0;
void note_path(int bits, bool reset) { g_path = reset ? bits : (g_path | bits); }

// One line on stderr, once per process and reason, when a problem the caller probably expected on the tensor cores
// falls back to the fp32 CUDA-core kernels (3-5x slower).  HALO_QUIET=1 silences it; halo_last_path() always tells.
void warn_slow_path_once(int reason, const char* fmt, ...) {
  static unsigned warned = 0;  // benign race: at worst a duplicate line
  if (reason < 0 || reason > 31 || (warned & (1u << reason))) return;
  warned |= 1u << reason;
  const char* q = getenv("HALO_QUIET");
  if (q && q[0] == '1') return;
  char msg[384];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  fprintf(stderr, "[halo_b200] slow path: %s (query halo_last_path(); HALO_QUIET=1 silences this)\n", msg);
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

#ifndef HALO_SOURCE_HASH
#define HALO_SOURCE_HASH "unknown"
#endif
// content hash of the sources this library was built from; halo_b200/_build.py greps it to detect a stale library
extern "C" const char* halo_source_hash(void) { return "HALO_SRC_SHA256=" HALO_SOURCE_HASH; }
extern "C" int halo_abi_version(void) { return HALO_ABI_VERSION; }
extern "C" const char* halo_last_error(void) { return halo::g_err; }
extern "C" int halo_last_path(void) { return halo::g_path; }
