// Shared helpers for libhalo_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/halo_b200.h"

namespace halo {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);
void note_path(int bits, bool reset);                          // halo_last_path(): which kernel variant the last head call took
void warn_slow_path_once(int reason, const char* fmt, ...);    // one stderr line per process and reason

#define HALO_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      halo::set_error(__VA_ARGS__);               \
      return HALO_ERR_BAD_ARG;                    \
    }                                             \
  } while (0)

#define HALO_CUDA(call)                                        \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return halo::cuda_fail(_e, #call);  \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return HALO_OK;
}

int sm_count();

// ---- constants of the head that must be formed in double on the host ------------------------------
struct HeadConsts {
  float c, inv_c, s, inv_s, two_over_s, two_s;
  float z_clip;      // artanh(1-1e-5): tanh(z) > 1-1e-5  <=>  z > z_clip   (geoopt project eps for fp64)
  float t_clip;      // 1-1e-5
  float omega_clip;  // 1-(1-1e-5)^2
  float maxnorm;     // (1-1e-3)/s                         (hyperbolic.py:162)
  float om_max;      // 1-(1-1e-3)^2 = 1 - c*maxnorm^2
  float out_scale;   // s * maxnorm * 2 / om_max           (projected-branch constant)
  float inv_log19;   // 1/log(19)                          (floating_region.py:74-76)
};

inline HeadConsts make_head_consts(float c) {
  HeadConsts h;
  double cd = (double)c, s = sqrt(cd);
  h.c = c;
  h.inv_c = (float)(1.0 / cd);
  h.s = (float)s;
  h.inv_s = (float)(1.0 / s);
  h.two_over_s = (float)(2.0 / s);
  h.two_s = (float)(2.0 * s);
  h.z_clip = (float)atanh(1.0 - 1e-5);
  h.t_clip = (float)(1.0 - 1e-5);
  h.omega_clip = (float)(1.0 - (1.0 - 1e-5) * (1.0 - 1e-5));
  h.maxnorm = (float)((1.0 - 1e-3) / s);
  h.om_max = (float)(1.0 - (1.0 - 1e-3) * (1.0 - 1e-3));
  h.out_scale = (float)(s * ((1.0 - 1e-3) / s) * 2.0 / (1.0 - (1.0 - 1e-3) * (1.0 - 1e-3)));
  h.inv_log19 = (float)(1.0 / log(19.0));
  return h;
}

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// order-preserving float <-> uint map (so integer atomics give float min/max for any sign)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k) {
  unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace halo
