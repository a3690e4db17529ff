// Pieces of the head shared by the forward (head_fwd.cu) and backward (head_bwd.cu) translation units.
#pragma once
#include "common.cuh"

namespace halo {

// classes are padded to a multiple of 4 so that the 2*OP weight columns of a channel are whole float4s
inline int head_op_pad(int O) { return round_up(O, 4); }

// ---- parameter packing -------------------------------------------------------------------------------
// ws layout (floats): Wt[CPAD][2*OP]  (columns [0,OP): -P_k ; [OP,2OP): A_k/max(|A_k|,1e-12)), then
// cls[4][OP] = {pp=|P_k|^2, an=|A_k|, pa=<-P_k,a_hat_k>, Bk=1-c*pp}.  Padded rows/columns are zero.
static __global__ void head_pack_kernel(const float* __restrict__ P, const float* __restrict__ A, float c, int O, int OP,
                                 int C, int CPAD, float* __restrict__ ws, float* __restrict__ stats, int N) {
  const int k = blockIdx.x;
  const int KP = 2 * OP;
  float* cls = ws + (size_t)CPAD * KP;
  __shared__ double red[3][32];
  double pp = 0.0, aa = 0.0, pa = 0.0;
  if (k < O) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      double p = (double)P[(size_t)k * C + ch], a = (double)A[(size_t)k * C + ch];
      pp += p * p;
      aa += a * a;
      pa -= p * a;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    pp += __shfl_xor_sync(0xffffffffu, pp, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    pa += __shfl_xor_sync(0xffffffffu, pa, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = pp; red[1][warp] = aa; red[2][warp] = pa; }
  __syncthreads();
  pp = aa = pa = 0.0;
  for (int w = 0; w < nw; ++w) { pp += red[0][w]; aa += red[1][w]; pa += red[2][w]; }
  const double an = sqrt(aa);
  const double den = an > 1e-12 ? an : 1e-12;  // F.normalize eps (hyperbolic.py:173)
  for (int ch = threadIdx.x; ch < CPAD; ch += blockDim.x) {
    float q = 0.f, ah = 0.f;
    if (k < O && ch < C) {
      q = -P[(size_t)k * C + ch];
      ah = (float)((double)A[(size_t)k * C + ch] / den);
    }
    ws[(size_t)ch * KP + k] = q;
    ws[(size_t)ch * KP + OP + k] = ah;
  }
  if (threadIdx.x == 0) {
    cls[0 * OP + k] = (k < O) ? (float)pp : 0.f;
    cls[1 * OP + k] = (k < O) ? (float)an : 0.f;
    cls[2 * OP + k] = (k < O) ? (float)(pa / den) : 0.f;
    cls[3 * OP + k] = (k < O) ? (float)(1.0 - (double)c * pp) : 1.f;
  }
  if (stats != nullptr && k == 0) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      stats[4 * n + 0] = __int_as_float(0x7f800000);  // +inf: running min of a non-negative plane
      stats[4 * n + 1] = 0.f;                         // running max
      stats[4 * n + 2] = 0.f;
      stats[4 * n + 3] = 0.f;
    }
  }
}


}  // namespace halo
