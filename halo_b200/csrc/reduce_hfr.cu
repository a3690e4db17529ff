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

constexpr int RH_TP = 64;        // pixels per tile
constexpr int RH_TC = 64;        // output channels per tile
constexpr int RH_TK = 32;        // input channels per shared-memory stage
constexpr int RH_THREADS = 256;  // 16 x 16 threads, 4 px x 4 ch each

struct ReduceArgs {
  const float* f;     // [N,Cin,HW]
  const float* Wr;    // [C,Cin]
  const float* br;    // [C] | NULL
  float* y;           // [N,C,HW]
  float* part_sq;     // [N][tiles][C]  partial sums of y^2
  int N, Cin, C, HW, tiles;
};

__global__ void __launch_bounds__(RH_THREADS) reduce_kernel(const ReduceArgs a) {
  __shared__ float sF[RH_TK][RH_TP];        // [k][px]
  __shared__ float sW[RH_TK][RH_TC + 1];    // [k][ch]
  __shared__ float sSq[16][RH_TC];
  const int tile = blockIdx.x, n = blockIdx.z, cb = blockIdx.y * RH_TC;
  const int p0 = tile * RH_TP;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: pixel quad, ty: channel quad
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
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
      const float4 fv = *reinterpret_cast<const float4*>(&sF[k][tx * 4]);
      const float w0 = sW[k][ty * 4 + 0], w1 = sW[k][ty * 4 + 1], w2 = sW[k][ty * 4 + 2], w3 = sW[k][ty * 4 + 3];
      const float fp[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(fp[i], w0, acc[i][0]);
        acc[i][1] = fmaf(fp[i], w1, acc[i][1]);
        acc[i][2] = fmaf(fp[i], w2, acc[i][2]);
        acc[i][3] = fmaf(fp[i], w3, acc[i][3]);
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
    float* yr = a.y + ((size_t)n * a.C + c) * a.HW + p0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (p0 + tx * 4 + i < a.HW) {
        const float v = acc[i][j] + b;
        yr[i] = v;
        sq[j] = fmaf(v, v, sq[j]);
      }
    }
  }
  // per-channel sum of squares of this tile: fixed-order reduction over the 16 pixel quads
#pragma unroll
  for (int j = 0; j < 4; ++j) sSq[tx][ty * 4 + j] = sq[j];
  __syncthreads();
  if (threadIdx.x < RH_TC && cb + threadIdx.x < a.C) {
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
