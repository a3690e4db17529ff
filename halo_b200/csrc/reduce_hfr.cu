// K7 -- channel reduction + hyperbolic feature re-weighting (HFR) immediately upstream of the head, evaluation mode
// (the acquisition round runs the classifier in eval mode, core/active/build.py:72-73).
//
// Reference (core/models/classifier.py:526-550, DepthwiseSeparableASPP_Hyper.forward; the v2 head has the same block :187-214):
//     y  = conv_reduce(f)                                  1x1 conv, Cin (512) -> C, with bias                      :527
//     t  = wn_mlp(y as (N*h*w, C) rows)                    Linear -> BatchNorm1d (eval: running stats) -> ReLU -> Linear  :531-536
//     wt = clamp(mean over the image's pixels of t, 1e-5)  (N, C, 1, 1)                                              :537-542
//     z  = F.normalize(y as (N, C, h*w), dim=-1) * wt      every channel plane scaled to unit L2 norm over the pixels  :543-550
// z is what HyperMapper.expmap / the fused head read next.  The second Linear commutes with the pixel mean, so only the
// hidden activations relu(bn(W1 y + b1)) are averaged per pixel.  Three kernels, y written once:
//   reduce_kernel   y = Wr f + br on 64 px x 64 ch register tiles (fp32 FMA; the GEMM is 4 % of the backbone's cost at this
//                   shape, not worth a tcgen05 pipeline), per-block partial sums of y^2 per channel
//   hidden_kernel   per pixel h = relu(bn(W1 y + b1)), per-block partial sums of h per channel      (only with HFR)
//   scale_kernel    fixed-order finish of the partials -> s[n][c] = clamp(W2 hbar + b2, 1e-5) / max(|y_c|, 1e-12); z = y * s
// All sums are reduced in a fixed order: results are bitwise reproducible.
#include "common.cuh"

namespace halo {

constexpr int RH_TP = 128;       // pixels per tile
constexpr int RH_TC = 64;        // output channels per tile
constexpr int RH_TK = 32;        // input channels per shared-memory stage
constexpr int RH_THREADS = 256;  // 16 x 16 threads, 8 px x 4 ch each

struct ReduceArgs {
  const float* f;     // [N,Cin,HW]
  const float* Wr;    // [C,Cin]
  const float* br;    // [C] | NULL
  float* y;           // [N,C,HW]
  float* part_sq;     // [N][tiles][C]  partial sums of y^2 | NULL
  int N, Cin, C, HW, tiles;
};

// y[c][p] = sum_k W[c][k] f[k][p] + b[c].  A thread owns pixels {4 tx .. 4 tx + 3} and {64 + 4 tx ..} of the tile and channels
// {4 ty ..}: per input channel two conflict-free LDS.128 of features and one broadcast LDS.128 of weights feed 32 FMAs.
__global__ void __launch_bounds__(RH_THREADS) reduce_kernel(const ReduceArgs a) {
  __shared__ __align__(16) float sF[RH_TK][RH_TP];   // [k][px]
  __shared__ __align__(16) float sW[RH_TK][RH_TC];   // [k][ch]
  __shared__ float sSq[16][RH_TC];
  const int tile = blockIdx.x, n = blockIdx.z, cb = blockIdx.y * RH_TC;
  const int p0 = tile * RH_TP;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: pixel quads, ty: channel quad
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* fn = a.f + (size_t)n * a.Cin * a.HW;
  for (int k0 = 0; k0 < a.Cin; k0 += RH_TK) {
    for (int i = threadIdx.x; i < RH_TK * RH_TP; i += RH_THREADS) {
      const int k = i / RH_TP, p = i - k * RH_TP;
      sF[k][p] = (k0 + k < a.Cin && p0 + p < a.HW) ? __ldg(fn + (size_t)(k0 + k) * a.HW + p0 + p) : 0.f;
    }
    for (int i = threadIdx.x; i < RH_TK * RH_TC; i += RH_THREADS) {
      const int c = i / RH_TK, k = i - c * RH_TK;
      sW[k][c] = (k0 + k < a.Cin && cb + c < a.C) ? __ldg(a.Wr + (size_t)(cb + c) * a.Cin + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < RH_TK; ++k) {
      const float4 f0 = *reinterpret_cast<const float4*>(&sF[k][tx * 4]);
      const float4 f1 = *reinterpret_cast<const float4*>(&sF[k][64 + tx * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&sW[k][ty * 4]);
      const float fp[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(fp[i], wv.x, acc[i][0]);
        acc[i][1] = fmaf(fp[i], wv.y, acc[i][1]);
        acc[i][2] = fmaf(fp[i], wv.z, acc[i][2]);
        acc[i][3] = fmaf(fp[i], wv.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cb + ty * 4 + j;
    if (c >= a.C) continue;
    const float b = (a.br != nullptr) ? a.br[c] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pp = p0 + h * 64 + tx * 4;
      float* yr = a.y + ((size_t)n * a.C + c) * a.HW + pp;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = acc[h * 4 + i][j] + b;
      if (pp + 3 < a.HW && (a.HW & 3) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
        *reinterpret_cast<float4*>(yr) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) sq[j] = fmaf(v[i], v[i], sq[j]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (pp + i < a.HW) {
            yr[i] = v[i];
            sq[j] = fmaf(v[i], v[i], sq[j]);
          }
        }
      }
    }
  }
  // per-channel sum of squares of this tile: fixed-order reduction over the 16 pixel groups
#pragma unroll
  for (int j = 0; j < 4; ++j) sSq[tx][ty * 4 + j] = sq[j];
  __syncthreads();
  if (a.part_sq != nullptr && threadIdx.x < RH_TC && cb + threadIdx.x < a.C) {
    float s = 0.f;
    for (int q = 0; q < 16; ++q) s += sSq[q][threadIdx.x];
    a.part_sq[((size_t)n * a.tiles + tile) * a.C + cb + threadIdx.x] = s;
  }
}

struct HiddenArgs {
  const float* y;      // [N,C,HW]
  const float* W1;     // [C,C]
  const float* b1;     // [C]
  const float* bn_scale;  // [C] gamma / sqrt(running_var + eps)
  const float* bn_shift;  // [C] beta - running_mean * bn_scale
  float* part_h;       // [N][blocks][C]
  int N, C, HW, blocks;
};

// thread = pixel; the pixel's C reduced features sit in registers (C <= 128), W1 rows are read from shared memory
template <int CMAX>
__global__ void __launch_bounds__(128) hidden_kernel(const HiddenArgs a) {
  extern __shared__ float sm[];
  float* sW1 = sm;                      // [C][C]
  float* sRed = sm + (size_t)a.C * a.C; // [4 warps][C]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < a.C * a.C; i += 128) sW1[i] = a.W1[i];
  for (int i = threadIdx.x; i < 4 * a.C; i += 128) sRed[i] = 0.f;
  __syncthreads();
  const int p = blockIdx.x * 128 + threadIdx.x;
  const bool live = p < a.HW;
  float yv[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) yv[c] = (live && c < a.C) ? a.y[((size_t)n * a.C + c) * a.HW + p] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < a.C; ++j) {
    const float* wr = sW1 + (size_t)j * a.C;
    float t = a.b1[j];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < a.C) t = fmaf(wr[c], yv[c], t);
    float h = fmaxf(fmaf(t, a.bn_scale[j], a.bn_shift[j]), 0.f);
    if (!live) h = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) sRed[warp * a.C + j] = h;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.C; j += 128)
    a.part_h[((size_t)n * a.blocks + blockIdx.x) * a.C + j] = sRed[j] + sRed[a.C + j] + sRed[2 * a.C + j] + sRed[3 * a.C + j];
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* m, const float* v, float eps, float* bn, int C) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
    const float sc = g[j] / sqrtf(v[j] + eps);
    bn[j] = sc;
    bn[C + j] = b[j] - m[j] * sc;
  }
}

struct ScaleArgs {
  float* y;               // [N,C,HW] in place -> z
  const float* part_sq;   // [N][tiles][C]
  const float* part_h;    // [N][blocks][C] | NULL (no HFR: z = y)
  const float* W2;        // [C,C]
  const float* b2;        // [C]
  float* scale;           // [N][C] out (diagnostic / tests)
  float* small;           // [N][3][C] | NULL: hbar, tbar (before the clamp), |y|  -- what the training backward needs
  int N, C, HW, tiles, blocks;
};

// one block per image: finish the partial sums in a fixed order, derive the per-channel scale
__global__ void hfr_scale_kernel(const ScaleArgs a) {
  extern __shared__ float sm[];
  float* hbar = sm;   // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < a.blocks; ++b) s += a.part_h[((size_t)n * a.blocks + b) * a.C + c];
    hbar[c] = s / (float)a.HW;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float sq = 0.f;
    for (int t = 0; t < a.tiles; ++t) sq += a.part_sq[((size_t)n * a.tiles + t) * a.C + c];
    float wt = a.b2[c];
    for (int j = 0; j < a.C; ++j) wt = fmaf(a.W2[(size_t)c * a.C + j], hbar[j], wt);
    if (a.small != nullptr) {
      a.small[((size_t)n * 3 + 0) * a.C + c] = hbar[c];
      a.small[((size_t)n * 3 + 1) * a.C + c] = wt;
      a.small[((size_t)n * 3 + 2) * a.C + c] = sqrtf(sq);
    }
    wt = fmaxf(wt, 1e-5f);                         // torch.clamp(norm_weights, min=1e-5)   (:542)
    a.scale[(size_t)n * a.C + c] = wt / fmaxf(sqrtf(sq), 1e-12f);   // F.normalize eps        (:546)
  }
}

__global__ void hfr_apply_kernel(float* __restrict__ y, const float* __restrict__ scale, int C, int HW, long long total4) {
  const int hw4 = HW / 4;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total4; g += (long long)gridDim.x * blockDim.x) {
    const long long plane = g / hw4;            // n * C + c
    const float s = scale[plane];
    float4 v = reinterpret_cast<float4*>(y)[g];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(y)[g] = v;
  }
}
__global__ void hfr_apply_scalar_kernel(float* __restrict__ y, const float* __restrict__ scale, int HW, long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x)
    y[g] *= scale[g / HW];
}


// =====================================================================================================================
// Training mode (the training step's side of classifier.py:526-550): BatchNorm1d normalises with the
// statistics of the batch (all N*h*w rows), and autograd runs back through the re-weighting, the normalisation and the
// 1x1 convolution.  Forward = the kernels above with the batch statistics folded in; backward:
//   dwt[n,c] = <dz, y_hat> ; dy1 = wt/|y| (dz - y_hat dwt)                               (normalise * weight)
//   dtbar = dwt [tbar >= 1e-5] ; dW2 = sum_n dtbar (x) hbar ; db2 = sum_n dtbar ; v[n,:] = W2^T dtbar[n] / HW  (mean, Linear 2)
//   g = v [b > 0] ; dbeta = sum g ; dgamma = sum g x_hat ; da = gamma/sigma (g - dbeta/M - x_hat dgamma/M)   (ReLU, BatchNorm)
//   dW1 = sum da (x) y ; db1 = sum da ; dy = dy1 + W1^T da                                (Linear 1)
//   df = Wr^T dy ; dWr = sum dy (x) f ; dbr = sum dy                                        (1x1 convolution)
// Every sum over pixels goes through per-block partials and a fixed-order finish: bitwise reproducible.
// =====================================================================================================================

// The hidden pre-activations a = W1 y + b1 are one more 1x1 convolution (reduce_kernel with Cin = C) kept as a plane tensor
// [N][C][HW]; everything BatchNorm / ReLU needs in either direction is then a per-plane reduction or an element-wise pass.

// fixed-order block sum of one value per thread (256 threads).  The plane sums are accumulated in DOUBLE: the backward's
// BatchNorm terms cancel over all N*HW pixels (sum of da = 0), so a 1e-5 relative error in a plane sum -- what 200
// same-signed fp32 additions per thread give -- comes back multiplied by the pixel count in dW1 and dfeat.
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// one block per (n, j) plane: sum and M2 about the plane's own mean (two passes over an L2-resident plane; a one-pass
// E[a^2] - mean^2 loses the variance when |mean| >> sigma)
__global__ void __launch_bounds__(256) plane_stats_kernel(const float* __restrict__ a, double* __restrict__ part /* [planes][2] */, int HW) {
  __shared__ double red[256];
  const float* pl = a + (size_t)blockIdx.x * HW;
  double s = 0.0;
  for (int p = threadIdx.x; p < HW; p += 256) s += (double)pl[p];
  const double sum = block_sum_256(s, red);
  const float mean = (float)(sum / (double)HW);
  double m2 = 0.0;
  for (int p = threadIdx.x; p < HW; p += 256) {
    const float d = pl[p] - mean;
    m2 += (double)(d * d);
  }
  m2 = block_sum_256(m2, red);
  // M2 about the rounded mean -> about the exact one: sum (a - m)^2 = sum (a - mf)^2 - HW (m - mf)^2
  const double dm = sum / (double)HW - (double)mean;
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = sum; part[2 * blockIdx.x + 1] = m2 - (double)HW * dm * dm; }
}

// combine the N planes of every hidden unit in a fixed order (double), fold the batch statistics into scale / shift
__global__ void plane_stats_finish_kernel(const double* __restrict__ part, int N, int HW, int C, const float* g, const float* b, float eps,
                                          float* bn /* [2][C] */, float* stats /* [2][C] */) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
    double S = 0.0;
    for (int n = 0; n < N; ++n) S += part[2 * ((size_t)n * C + j)];
    const double mean = S / ((double)N * HW);
    double M2 = 0.0;
    for (int n = 0; n < N; ++n) {
      const double d = part[2 * ((size_t)n * C + j)] / (double)HW - mean;
      M2 += part[2 * ((size_t)n * C + j) + 1] + d * d * HW;
    }
    const double var = M2 / ((double)N * HW);   // biased, what BatchNorm normalises with
    const float sc = (float)((double)g[j] / sqrt(var + (double)eps));
    bn[j] = sc;
    bn[C + j] = (float)((double)b[j] - mean * (double)sc);
    stats[j] = (float)mean;
    stats[C + j] = (float)var;
  }
}

// one block per (n, j) plane: sum over the pixels of relu(bn(a)) -> part_h[n][j]
__global__ void __launch_bounds__(256) plane_hidden_sum_kernel(const float* __restrict__ a, const float* __restrict__ bn, float* __restrict__ out,
                                                               int C, int HW) {
  __shared__ double red[256];
  const int j = blockIdx.x % C;
  const float sc = bn[j], sh = bn[C + j];
  const float* pl = a + (size_t)blockIdx.x * HW;
  double s = 0.0;
  for (int p = threadIdx.x; p < HW; p += 256) s += (double)fmaxf(fmaf(pl[p], sc, sh), 0.f);
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)s;
}

// one block per (n, j) plane: sum g and sum g * x_hat with g = v[n][j] * [bn(a) > 0]  -> part[n][2][C]
__global__ void __launch_bounds__(256) plane_bn_sums_kernel(const float* __restrict__ a, const float* __restrict__ bn, const float* __restrict__ stats,
                                                            const float* __restrict__ v, float eps, double* __restrict__ part, int C, int HW) {
  __shared__ double red[256];
  const int n = blockIdx.x / C, j = blockIdx.x - n * C;
  const float sc = bn[j], sh = bn[C + j], mean = stats[j], isd = rsqrtf(stats[C + j] + eps);
  const double vj = (double)v[blockIdx.x];
  const float* pl = a + (size_t)blockIdx.x * HW;
  double cnt = 0.0, sx = 0.0;     // number of active pixels, sum of their x_hat
  for (int p = threadIdx.x; p < HW; p += 256) {
    const float t = pl[p];
    if (fmaf(t, sc, sh) > 0.f) { cnt += 1.0; sx += (double)((t - mean) * isd); }
  }
  cnt = block_sum_256(cnt, red);
  sx = block_sum_256(sx, red);
  if (threadIdx.x == 0) { part[((size_t)n * 2 + 0) * C + j] = vj * cnt; part[((size_t)n * 2 + 1) * C + j] = vj * sx; }
}

// element-wise: da = gamma/sigma (g - dbeta/M - x_hat dgamma/M)
__global__ void hfr_da_elem_kernel(const float* __restrict__ a, const float* __restrict__ bn, const float* __restrict__ stats,
                                   const float* __restrict__ v, const float* __restrict__ dbg, float eps, float inv_m,
                                   float* __restrict__ da, int C, int HW, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long plane = i / HW;
    const int j = (int)(plane % C);
    const float t = a[i], sc = bn[j];
    const float g = (fmaf(t, sc, bn[C + j]) > 0.f) ? v[plane] : 0.f;
    const float xh = (t - stats[j]) * rsqrtf(stats[C + j] + eps);
    da[i] = sc * (g - dbg[j] * inv_m - xh * dbg[C + j] * inv_m);
  }
}
// element-wise: dy = dy2 + coef0 dz - coef1 y   (dy holds W1^T da on entry)
__global__ void hfr_dy_finish_kernel(float* __restrict__ dy, const float* __restrict__ dz, const float* __restrict__ y,
                                     const float* __restrict__ coef, int C, int HW, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long plane = i / HW;
    const int n = (int)(plane / C), c = (int)(plane - (long long)n * C);
    dy[i] = fmaf(coef[((size_t)n * 2 + 0) * C + c], dz[i], fmaf(-coef[((size_t)n * 2 + 1) * C + c], y[i], dy[i]));
  }
}

__global__ void hfr_apply_oop_kernel(const float* __restrict__ y, const float* __restrict__ scale, float* __restrict__ z, int HW,
                                     long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x)
    z[g] = y[g] * scale[g / HW];
}

// q[plane] = sum_p a[plane][p] * b[plane][p], one block per (n, c) plane, fixed-order tree, double accumulation
__global__ void __launch_bounds__(256) dot_planes_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ q, int HW) {
  __shared__ double red[256];
  const size_t base = (size_t)blockIdx.x * HW;
  double s = 0.0;
  for (int p = threadIdx.x; p < HW; p += 256) s += (double)a[base + p] * (double)b[base + p];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) q[blockIdx.x] = (float)red[0];
}

struct HfrSmallArgs {
  const float* q;        // [N][C]  <dz, y>
  const float* small;    // [N][3][C]  hbar, tbar (before the clamp), |y|
  const float* W2;       // [C][C]
  float* coef;           // [N][2][C]  dy1 = coef0 * dz - coef1 * y
  float* v;              // [N][C]     dL/dh of every pixel of image n
  float* dW2;            // [C][C]
  float* db2;            // [C]
  int N, C, HW;
};
// one block: the per-image scalars of the backward
__global__ void hfr_bwd_small_kernel(const HfrSmallArgs a) {
  extern __shared__ float sm[];
  float* dtb = sm;   // [N][C]
  const int C = a.C;
  for (int i = threadIdx.x; i < a.N * C; i += blockDim.x) {
    const int n = i / C, c = i - n * C;
    const float tbar = a.small[((size_t)n * 3 + 1) * C + c], yn = a.small[((size_t)n * 3 + 2) * C + c];
    const float wt = fmaxf(tbar, 1e-5f);
    const float den = fmaxf(yn, 1e-12f);
    const float dwt = a.q[i] / den;                    // <dz, y_hat>
    dtb[i] = (tbar >= 1e-5f) ? dwt : 0.f;              // clamp(min) passes the gradient where the input is >= min
    const float sc = wt / den;
    a.coef[((size_t)n * 2 + 0) * C + c] = sc;
    a.coef[((size_t)n * 2 + 1) * C + c] = (yn > 1e-12f) ? sc * dwt / den : 0.f;   // below eps F.normalize divides by a constant
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.N * C; i += blockDim.x) {
    const int n = i / C, j = i - n * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(a.W2[(size_t)c * C + j], dtb[n * C + c], s);
    a.v[i] = s / (float)a.HW;
  }
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    const int c = i / C, j = i - c * C;
    float s = 0.f;
    for (int n = 0; n < a.N; ++n) s = fmaf(dtb[n * C + c], a.small[((size_t)n * 3 + 0) * C + j], s);
    a.dW2[i] = s;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < a.N; ++n) s += dtb[n * C + c];
    a.db2[c] = s;
  }
}

__global__ void sum_partials_f64_kernel(const double* __restrict__ part, int groups, int rows, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < groups; ++g) s += part[(size_t)g * rows + i];
    out[i] = (float)s;
  }
}
// fixed-order finish of [groups][rows] partials in double: out[i] = sum_g part[g][i]
__global__ void sum_partials_kernel(const float* __restrict__ part, int groups, int rows, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < groups; ++g) s += part[(size_t)g * rows + i];
    out[i] = (float)s;
  }
}

// Reduction over pixels: part[n][chunk][i][j] = sum_{p in chunk} A[n][i][p] * B[n][j][p], and rsum[n][chunk][i] = sum_p A[n][i][p]
// (dW1 / db1 with A = da, B = y; dWr / dbr with A = dy, B = f).  64 x 128 output tiles, 32 pixels per shared-memory stage; a
// thread owns rows {4 ty ..} of A and rows {4 tx ..}, {64 + 4 tx ..} of B: three LDS.128 feed 32 FMAs per pixel.
struct GramArgs {
  const float* A;   // [N][Ca][HW]
  const float* B;   // [N][Cb][HW]
  float* part;      // [N*chunks][Ca][Cb]
  float* rsum;      // [N*chunks][Ca] | NULL
  int Ca, Cb, HW, chunk, chunks;
};
__global__ void __launch_bounds__(256) gram_px_kernel(const GramArgs a) {
  __shared__ __align__(16) float sA[32][64];    // [px][row]
  __shared__ __align__(16) float sB[32][128];
  const int n = blockIdx.z / a.chunks, ch = blockIdx.z - n * a.chunks;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 128;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int p_lo = ch * a.chunk, p_hi = min(a.HW, p_lo + a.chunk);
  float acc[4][8];
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float* An = a.A + (size_t)n * a.Ca * a.HW;
  const float* Bn = a.B + (size_t)n * a.Cb * a.HW;
  const int ra = threadIdx.x & 63, pa = (threadIdx.x >> 6) * 8;      // loader: row, first of 8 pixels  (A: 64 rows x 32 px)
  const int rb = threadIdx.x & 127, pb = (threadIdx.x >> 7) * 16;    //                                 (B: 128 rows x 32 px)
  for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
    {
      const bool okr = i0 + ra < a.Ca;
      const float* src = An + (size_t)(i0 + ra) * a.HW + p0 + pa;
#pragma unroll
      for (int e = 0; e < 8; ++e) sA[pa + e][ra] = (okr && p0 + pa + e < p_hi) ? __ldg(src + e) : 0.f;
    }
    {
      const bool okr = j0 + rb < a.Cb;
      const float* src = Bn + (size_t)(j0 + rb) * a.HW + p0 + pb;
#pragma unroll
      for (int e = 0; e < 16; ++e) sB[pb + e][rb] = (okr && p0 + pb + e < p_hi) ? __ldg(src + e) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < 32; ++px) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[px][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[px][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sB[px][64 + tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i] += ar[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* o = a.part + (size_t)blockIdx.z * a.Ca * a.Cb;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (i0 + ty * 4 + i < a.Ca && col < a.Cb) o[(size_t)(i0 + ty * 4 + i) * a.Cb + col] = acc[i][j];
    }
  if (a.rsum != nullptr && blockIdx.x == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i0 + ty * 4 + i < a.Ca) a.rsum[(size_t)blockIdx.z * a.Ca + i0 + ty * 4 + i] = rs[i];
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {   // out[c][r] = in[r][c]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * Cc; i += gridDim.x * blockDim.x) {
    const int r = i / Cc, c = i - r * Cc;
    out[(size_t)c * R + r] = in[i];
  }
}

}  // namespace halo

using namespace halo;

static inline size_t rh_align(size_t b) { return (b + 255) / 256 * 256; }

extern "C" size_t halo_reduce_hfr_workspace_bytes(int N, int C, int H, int W) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const size_t HW = (size_t)H * W;
  const size_t tiles = (HW + RH_TP - 1) / RH_TP, blocks = (HW + 127) / 128;
  return rh_align((size_t)N * tiles * C * 4) + rh_align((size_t)N * blocks * C * 4) + rh_align((size_t)N * C * 4) + rh_align(2 * (size_t)C * 4);
}

extern "C" int halo_reduce_hfr_fwd(const float* feat, const float* Wr, const float* br, const float* W1, const float* b1,
                                   const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var,
                                   float bn_eps, const float* W2, const float* b2, float* out, float* scale_out, int N, int Cin,
                                   int C, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && Wr && out, "halo_reduce_hfr_fwd: NULL pointer");
  HALO_CHECK_ARG(N > 0 && Cin > 0 && C > 0 && H > 0 && W > 0, "halo_reduce_hfr_fwd: bad dims");
  const bool hfr = (W1 != nullptr);
  HALO_CHECK_ARG(!hfr || (b1 && bn_gamma && bn_beta && bn_mean && bn_var && W2 && b2),
                 "halo_reduce_hfr_fwd: the re-weighting MLP needs all of W1, b1, BatchNorm statistics, W2, b2");
  if (hfr && C > 128) {
    set_error("halo_reduce_hfr_fwd: HFR with %d reduced channels not compiled (<= 128)", C);
    return HALO_ERR_UNSUPPORTED;
  }
  HALO_CHECK_ARG(N <= 65535, "halo_reduce_hfr_fwd: batch too large");
  const size_t need = halo_reduce_hfr_workspace_bytes(N, C, H, W);
  if (!ws || ws_bytes < need) {
    set_error("halo_reduce_hfr_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int tiles = (HW + RH_TP - 1) / RH_TP, blocks = (HW + 127) / 128;
  unsigned char* w8 = (unsigned char*)ws;
  float* part_sq = (float*)w8;
  w8 += rh_align((size_t)N * tiles * C * 4);
  float* part_h = (float*)w8;
  w8 += rh_align((size_t)N * blocks * C * 4);
  float* scale = (float*)w8;
  w8 += rh_align((size_t)N * C * 4);
  float* bn = (float*)w8;   // [2][C]: scale, shift

  ReduceArgs ra;
  ra.f = feat; ra.Wr = Wr; ra.br = br; ra.y = out; ra.part_sq = part_sq;
  ra.N = N; ra.Cin = Cin; ra.C = C; ra.HW = HW; ra.tiles = tiles;
  reduce_kernel<<<dim3(tiles, (C + RH_TC - 1) / RH_TC, N), RH_THREADS, 0, st>>>(ra);
  int rc = launch_status("reduce_kernel");
  if (rc || !hfr) return rc;

  // BatchNorm1d in eval mode is an affine map per hidden unit: fold it on the device (no host round trip)
  bn_fold_kernel<<<1, 128, 0, st>>>(bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, bn, C);
  rc = launch_status("bn_fold_kernel");
  if (rc) return rc;

  HiddenArgs ha;
  ha.y = out; ha.W1 = W1; ha.b1 = b1; ha.bn_scale = bn; ha.bn_shift = bn + C; ha.part_h = part_h;
  ha.N = N; ha.C = C; ha.HW = HW; ha.blocks = blocks;
  const size_t smem = ((size_t)C * C + 4 * C) * 4;
  if (C <= 64) {
    HALO_CUDA(cudaFuncSetAttribute(hidden_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hidden_kernel<64><<<dim3(blocks, N), 128, smem, st>>>(ha);
  } else {
    HALO_CUDA(cudaFuncSetAttribute(hidden_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hidden_kernel<128><<<dim3(blocks, N), 128, smem, st>>>(ha);
  }
  rc = launch_status("hidden_kernel");
  if (rc) return rc;

  ScaleArgs sa;
  sa.y = out; sa.part_sq = part_sq; sa.part_h = part_h; sa.W2 = W2; sa.b2 = b2; sa.scale = scale_out ? scale_out : scale;
  sa.small = nullptr;
  sa.N = N; sa.C = C; sa.HW = HW; sa.tiles = tiles; sa.blocks = blocks;
  hfr_scale_kernel<<<N, 128, (size_t)C * 4, st>>>(sa);
  rc = launch_status("hfr_scale_kernel");
  if (rc) return rc;
  const long long total = (long long)N * C * HW;
  if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    long long b = (total / 4 + 255) / 256;
    if (b > (long long)sm_count() * 16) b = (long long)sm_count() * 16;
    hfr_apply_kernel<<<(int)b, 256, 0, st>>>(out, sa.scale, C, HW, total / 4);
  } else {
    long long b = (total + 255) / 256;
    if (b > (long long)sm_count() * 16) b = (long long)sm_count() * 16;
    hfr_apply_scalar_kernel<<<(int)b, 256, 0, st>>>(out, sa.scale, HW, total);
  }
  return launch_status("hfr_apply_kernel");
}

// ---- training mode ------------------------------------------------------------------------------------------------------
namespace {
struct TrainWs {
  size_t part_sq, part_h, part_st, scale, bn, q, coef, v, part_bn, dbg, da, dy, gram, rsum, wt, total;
  int tiles, chunk, chunks;
};
TrainWs train_ws(int N, int Cin, int C, int H, int W) {
  TrainWs L;
  const size_t HW = (size_t)H * W;
  L.tiles = (int)((HW + RH_TP - 1) / RH_TP);
  size_t chunk = (HW + 63) / 64;     // pixel chunks of the gradient GEMMs: enough blocks to fill the GPU, at most 64 partials per image
  if (chunk < 1024) chunk = 1024;
  chunk = (chunk + 31) / 32 * 32;
  L.chunk = (int)chunk;
  L.chunks = (int)((HW + chunk - 1) / chunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += rh_align(bytes); return o; };
  L.part_sq = take((size_t)N * L.tiles * C * 4);
  L.part_h = take((size_t)N * C * 4);
  L.part_st = take((size_t)N * C * 2 * 8);
  L.scale = take((size_t)N * C * 4);
  L.bn = take(2 * (size_t)C * 4);
  L.q = take((size_t)N * C * 4);
  L.coef = take((size_t)N * 2 * C * 4);
  L.v = take((size_t)N * C * 4);
  L.part_bn = take((size_t)N * 2 * C * 8);
  L.dbg = take(2 * (size_t)C * 4);
  L.da = take((size_t)N * C * HW * 4);
  L.dy = take((size_t)N * C * HW * 4);
  const size_t cmax = (size_t)(Cin > C ? Cin : C);
  L.gram = take((size_t)N * L.chunks * C * cmax * 4);
  L.rsum = take((size_t)N * L.chunks * C * 4);
  L.wt = take(cmax * C * 4);
  L.total = off;
  return L;
}
int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b > cap ? cap : b);
}
int conv1x1(const float* in, const float* Wm, const float* bias, float* out, float* part_sq, int N, int Cin, int C, int HW, int tiles,
            cudaStream_t st) {
  ReduceArgs ra;
  ra.f = in; ra.Wr = Wm; ra.br = bias; ra.y = out; ra.part_sq = part_sq;
  ra.N = N; ra.Cin = Cin; ra.C = C; ra.HW = HW; ra.tiles = tiles;
  reduce_kernel<<<dim3(tiles, (C + RH_TC - 1) / RH_TC, N), RH_THREADS, 0, st>>>(ra);
  return launch_status("reduce_kernel");
}
}  // namespace

extern "C" size_t halo_reduce_hfr_train_workspace_bytes(int N, int Cin, int C, int H, int W) {
  if (N <= 0 || Cin <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return train_ws(N, Cin, C, H, W).total;
}

extern "C" int halo_reduce_hfr_train_fwd(const float* feat, const float* Wr, const float* br, const float* W1, const float* b1,
                                         const float* bn_gamma, const float* bn_beta, float bn_eps, const float* W2,
                                         const float* b2, const float* fixed_stats, float* y_out, float* a_out, float* z_out,
                                         float* batch_stats, float* small, int N, int Cin, int C, int H, int W, void* ws,
                                         size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && Wr && y_out, "halo_reduce_hfr_train_fwd: NULL pointer");
  HALO_CHECK_ARG(N > 0 && Cin > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, "halo_reduce_hfr_train_fwd: bad dims");
  const bool hfr = (W1 != nullptr);
  HALO_CHECK_ARG(!hfr || (b1 && bn_gamma && bn_beta && W2 && b2 && a_out && z_out && batch_stats && small),
                 "halo_reduce_hfr_train_fwd: the re-weighting MLP needs W1, b1, gamma, beta, W2, b2 and the a / z / statistics outputs");
  const TrainWs L = train_ws(N, Cin, C, H, W);
  if (!ws || ws_bytes < L.total) {
    set_error("halo_reduce_hfr_train_fwd: workspace %zu < %zu bytes", ws_bytes, L.total);
    return HALO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  unsigned char* w8 = (unsigned char*)ws;
  float* part_sq = (float*)(w8 + L.part_sq);
  float* part_h = (float*)(w8 + L.part_h);
  double* part_st = (double*)(w8 + L.part_st);
  float* scale = (float*)(w8 + L.scale);
  float* bn = (float*)(w8 + L.bn);

  int rc = conv1x1(feat, Wr, br, y_out, hfr ? part_sq : nullptr, N, Cin, C, HW, L.tiles, st);
  if (rc || !hfr) return rc;
  rc = conv1x1(y_out, W1, b1, a_out, nullptr, N, C, C, HW, L.tiles, st);        // hidden pre-activations, kept for the backward
  if (rc) return rc;
  if (fixed_stats != nullptr) {   // BatchNorm1d in evaluation mode inside a differentiated step: its running statistics
    HALO_CUDA(cudaMemcpyAsync(batch_stats, fixed_stats, 2 * (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    bn_fold_kernel<<<1, 128, 0, st>>>(bn_gamma, bn_beta, fixed_stats, fixed_stats + C, bn_eps, bn, C);
    rc = launch_status("bn_fold_kernel");
  } else {
    plane_stats_kernel<<<N * C, 256, 0, st>>>(a_out, part_st, HW);
    rc = launch_status("plane_stats_kernel");
    if (rc) return rc;
    plane_stats_finish_kernel<<<1, 128, 0, st>>>(part_st, N, HW, C, bn_gamma, bn_beta, bn_eps, bn, batch_stats);
    rc = launch_status("plane_stats_finish_kernel");
  }
  if (rc) return rc;
  plane_hidden_sum_kernel<<<N * C, 256, 0, st>>>(a_out, bn, part_h, C, HW);
  rc = launch_status("plane_hidden_sum_kernel");
  if (rc) return rc;

  ScaleArgs sa;
  sa.y = y_out; sa.part_sq = part_sq; sa.part_h = part_h; sa.W2 = W2; sa.b2 = b2; sa.scale = scale; sa.small = small;
  sa.N = N; sa.C = C; sa.HW = HW; sa.tiles = L.tiles; sa.blocks = 1;
  hfr_scale_kernel<<<N, 128, (size_t)C * 4, st>>>(sa);
  rc = launch_status("hfr_scale_kernel");
  if (rc) return rc;
  const long long total = (long long)N * C * HW;
  hfr_apply_oop_kernel<<<grid_for(total), 256, 0, st>>>(y_out, scale, z_out, HW, total);
  return launch_status("hfr_apply_oop_kernel");
}

extern "C" int halo_reduce_hfr_train_bwd(const float* feat, const float* Wr, const float* W1, const float* bn_gamma,
                                         const float* bn_beta, float bn_eps, const float* W2, const float* y, const float* a,
                                         const float* batch_stats, const float* small, const float* dz, float* dfeat,
                                         float* dWr, float* dbr, float* dW1, float* db1, float* dgamma, float* dbeta, float* dW2,
                                         float* db2, int stats_are_batch, int N, int Cin, int C, int H, int W, void* ws,
                                         size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && Wr && dz && dWr, "halo_reduce_hfr_train_bwd: NULL pointer");
  HALO_CHECK_ARG(N > 0 && Cin > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, "halo_reduce_hfr_train_bwd: bad dims");
  const bool hfr = (W1 != nullptr);
  HALO_CHECK_ARG(!hfr || (bn_gamma && bn_beta && W2 && y && a && batch_stats && small && dW1 && db1 && dgamma && dbeta && dW2 && db2),
                 "halo_reduce_hfr_train_bwd: the re-weighting MLP needs its parameters, the saved y / a / statistics and all gradient outputs");
  const TrainWs L = train_ws(N, Cin, C, H, W);
  if (!ws || ws_bytes < L.total) {
    set_error("halo_reduce_hfr_train_bwd: workspace %zu < %zu bytes", ws_bytes, L.total);
    return HALO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const long long total = (long long)N * C * HW;
  unsigned char* w8 = (unsigned char*)ws;
  float* bn = (float*)(w8 + L.bn);
  float* q = (float*)(w8 + L.q);
  float* coef = (float*)(w8 + L.coef);
  float* v = (float*)(w8 + L.v);
  double* part_bn = (double*)(w8 + L.part_bn);
  float* dbg = (float*)(w8 + L.dbg);
  float* da = (float*)(w8 + L.da);
  float* dyb = (float*)(w8 + L.dy);
  float* gram = (float*)(w8 + L.gram);
  float* rsum = (float*)(w8 + L.rsum);
  float* wt = (float*)(w8 + L.wt);
  int rc;
  const float* dy = dz;   // without the re-weighting the reduced features ARE the output
  if (hfr) {
    dot_planes_kernel<<<N * C, 256, 0, st>>>(dz, y, q, HW);
    rc = launch_status("dot_planes_kernel");
    if (rc) return rc;
    HfrSmallArgs sa;
    sa.q = q; sa.small = small; sa.W2 = W2; sa.coef = coef; sa.v = v; sa.dW2 = dW2; sa.db2 = db2; sa.N = N; sa.C = C; sa.HW = HW;
    const size_t sm_small = (size_t)N * C * 4;
    HALO_CHECK_ARG(sm_small <= 200 * 1024, "halo_reduce_hfr_train_bwd: batch too large for the per-image scalar kernel");
    HALO_CUDA(cudaFuncSetAttribute(hfr_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
    hfr_bwd_small_kernel<<<1, 256, sm_small, st>>>(sa);
    rc = launch_status("hfr_bwd_small_kernel");
    if (rc) return rc;
    bn_fold_kernel<<<1, 128, 0, st>>>(bn_gamma, bn_beta, batch_stats, batch_stats + C, bn_eps, bn, C);
    rc = launch_status("bn_fold_kernel");
    if (rc) return rc;
    plane_bn_sums_kernel<<<N * C, 256, 0, st>>>(a, bn, batch_stats, v, bn_eps, part_bn, C, HW);
    rc = launch_status("plane_bn_sums_kernel");
    if (rc) return rc;
    sum_partials_f64_kernel<<<1, 128, 0, st>>>(part_bn, N, 2 * C, dbg);
    rc = launch_status("sum_partials_kernel");
    if (rc) return rc;
    HALO_CUDA(cudaMemcpyAsync(dbeta, dbg, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    HALO_CUDA(cudaMemcpyAsync(dgamma, dbg + C, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    // statistics that do not depend on the batch (evaluation-mode BatchNorm): the two correction terms vanish
    const float inv_m = stats_are_batch ? (float)(1.0 / ((double)N * HW)) : 0.f;
    hfr_da_elem_kernel<<<grid_for(total), 256, 0, st>>>(a, bn, batch_stats, v, dbg, bn_eps, inv_m, da, C, HW, total);
    rc = launch_status("hfr_da_elem_kernel");
    if (rc) return rc;
    // dy = W1^T da + (normalise * weight part)
    transpose_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(W1, wt, C, C);
    rc = conv1x1(da, wt, nullptr, dyb, nullptr, N, C, C, HW, L.tiles, st);
    if (rc) return rc;
    hfr_dy_finish_kernel<<<grid_for(total), 256, 0, st>>>(dyb, dz, y, coef, C, HW, total);
    rc = launch_status("hfr_dy_finish_kernel");
    if (rc) return rc;
    dy = dyb;
    // dW1 = sum da (x) y, db1 = sum da
    GramArgs g1;
    g1.A = da; g1.B = y; g1.part = gram; g1.rsum = rsum; g1.Ca = C; g1.Cb = C; g1.HW = HW; g1.chunk = L.chunk; g1.chunks = L.chunks;
    gram_px_kernel<<<dim3((C + 127) / 128, (C + 63) / 64, N * L.chunks), 256, 0, st>>>(g1);
    rc = launch_status("gram_px_kernel");
    if (rc) return rc;
    sum_partials_kernel<<<(C * C + 127) / 128, 128, 0, st>>>(gram, N * L.chunks, C * C, dW1);
    sum_partials_kernel<<<1, 128, 0, st>>>(rsum, N * L.chunks, C, db1);
    rc = launch_status("sum_partials_kernel");
    if (rc) return rc;
  }
  // dWr = sum dy (x) f, dbr = sum dy
  GramArgs g2;
  g2.A = dy; g2.B = feat; g2.part = gram; g2.rsum = rsum; g2.Ca = C; g2.Cb = Cin; g2.HW = HW; g2.chunk = L.chunk; g2.chunks = L.chunks;
  gram_px_kernel<<<dim3((Cin + 127) / 128, (C + 63) / 64, N * L.chunks), 256, 0, st>>>(g2);
  rc = launch_status("gram_px_kernel");
  if (rc) return rc;
  sum_partials_kernel<<<(C * Cin + 127) / 128, 128, 0, st>>>(gram, N * L.chunks, C * Cin, dWr);
  if (dbr != nullptr) sum_partials_kernel<<<1, 128, 0, st>>>(rsum, N * L.chunks, C, dbr);
  rc = launch_status("sum_partials_kernel");
  if (rc) return rc;
  if (dfeat != nullptr) {   // df = Wr^T dy: the forward's own tile kernel with the transposed weight
    transpose_kernel<<<(C * Cin + 255) / 256, 256, 0, st>>>(Wr, wt, C, Cin);
    rc = conv1x1(dy, wt, nullptr, dfeat, nullptr, N, C, Cin, HW, L.tiles, st);
  }
  return rc;
}
