// K4 -- fused backward of the Poincare-ball head (expmap0 + project + HyperMLR), CUDA-core fp32 path.
//
// The reference differentiates ~60 float64 autograd nodes (core/train_learners.py:238,362,457,557 through
// core/utils/hyperbolic.py:28-39,120-184).  Here the forward contractions are recomputed from the raw
// features and the whole epilogue is differentiated analytically per pixel:
//     logit_k = f(n2, S_k, T_k ; pp_k, an_k, pa_k),  n2=|u|^2, S_k=<u,-P_k>, T_k=<u,A_k/|A_k|>
//     du   = 2 u * sum_k G_k df/dn2 + sum_k (G_k df/dS_k) (-P_k) + (G_k df/dT_k) a_hat_k        (K4a, pass 2)
//     dW   = sum_pix [gS ; gT] (x) u                                                             (K4b)
//     dP, dA from dW and the per-class scalars d/dpp, d/dan, d/dpa by the chain rule             (K4c)
// Reductions over pixels are two-stage with a fixed order (per-CTA partials, then one finalize block per
// class), so gradients are bitwise reproducible for a given grid size.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"

namespace halo {

constexpr int BWD_THREADS = 256;
constexpr int BWD_PIX = 2;
constexpr int BWD_U = 4;
constexpr int DW_THREADS = 256;   // thread t owns channel t of a 256-channel block
constexpr int DW_PX = 32;         // pixels per unit (shared-memory strip)

struct BwdArgs {
  const float* feat;
  const float* dlogits;
  const float* wpack;   // Wt[CPAD][2OP] + cls[4][OP]
  float* dfeat;
  float* G;             // [N][2OP][HW]  gS rows then gT rows
  float* cls_part;      // [grid][3][OP]  per-CTA partial sums of d/dpp, d/dan, d/dpa
  int N, C, CPAD, O, HW, tiles_per_img, total_tiles;
  HeadConsts hc;
};

template <int OP, bool VEC>
__global__ void __launch_bounds__(BWD_THREADS, 2) head_bwd_pix_kernel(const BwdArgs a) {
  constexpr int KP = 2 * OP;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sCls = smem + (size_t)a.CPAD * KP;
  float* sRed = sCls + 4 * OP;  // [warps][3][OP]
  {
    const int n4 = (a.CPAD * KP + 4 * OP) / 4;
    const float4* src = reinterpret_cast<const float4*>(a.wpack);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < n4; i += BWD_THREADS) dst[i] = src[i];
    for (int i = threadIdx.x; i < (BWD_THREADS / 32) * 3 * OP; i += BWD_THREADS) sRed[i] = 0.f;
  }
  __syncthreads();
  const HeadConsts hc = a.hc;
  const int HW = a.HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img;
    const int p = (tile - n * a.tiles_per_img) * (BWD_THREADS * BWD_PIX) + threadIdx.x * BWD_PIX;
    const float* base = a.feat + (size_t)n * a.C * HW;

    float acc[BWD_PIX][KP];
    float n2[BWD_PIX];
#pragma unroll
    for (int i = 0; i < BWD_PIX; ++i) {
      n2[i] = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[i][k] = 0.f;
    }
    // ---- pass 1: recompute the forward contractions ----
    for (int cb = 0; cb < a.CPAD; cb += BWD_U) {
      float cur[BWD_U][BWD_PIX];
#pragma unroll
      for (int j = 0; j < BWD_U; ++j) {
        const int ch = cb + j;
#pragma unroll
        for (int i = 0; i < BWD_PIX; ++i) cur[j][i] = 0.f;
        if (ch < a.C) {
          if (VEC) {
            if (p < HW) {
              const float2 t = __ldg(reinterpret_cast<const float2*>(base + (size_t)ch * HW + p));
              cur[j][0] = t.x;
              cur[j][1] = t.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < BWD_PIX; ++i)
              if (p + i < HW) cur[j][i] = __ldg(base + (size_t)ch * HW + p + i);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < BWD_U; ++j) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)(cb + j) * KP);
#pragma unroll
        for (int i = 0; i < BWD_PIX; ++i) n2[i] = fmaf(cur[j][i], cur[j][i], n2[i]);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 w = w4[q];
#pragma unroll
          for (int i = 0; i < BWD_PIX; ++i) {
            acc[i][4 * q + 0] = fmaf(cur[j][i], w.x, acc[i][4 * q + 0]);
            acc[i][4 * q + 1] = fmaf(cur[j][i], w.y, acc[i][4 * q + 1]);
            acc[i][4 * q + 2] = fmaf(cur[j][i], w.z, acc[i][4 * q + 2]);
            acc[i][4 * q + 3] = fmaf(cur[j][i], w.w, acc[i][4 * q + 3]);
          }
        }
      }
    }

    // ---- epilogue: analytic derivative of the head w.r.t. (n2, S_k, T_k) and the class scalars ----
    float alpha[BWD_PIX];
    float cpp[OP], can[OP], cpa[OP];  // this thread's contribution to d/dpp, d/dan, d/dpa
#pragma unroll
    for (int k = 0; k < OP; ++k) cpp[k] = can[k] = cpa[k] = 0.f;
#pragma unroll
    for (int i = 0; i < BWD_PIX; ++i) {
      const bool live = (p + i < HW);
      const PixelScalarGrads ps = tangent_scalar_grads(n2[i], hc);
      float g_gamma = 0.f, g_t2 = 0.f, g_om = 0.f;
#pragma unroll
      for (int k = 0; k < OP; ++k) {
        float gS = 0.f, gT = 0.f;
        if (k < a.O) {
          const float G = live ? __ldg(a.dlogits + ((size_t)n * a.O + k) * HW + p + i) : 0.f;
          mlr_logit_grad(G, acc[i][k], acc[i][OP + k], ps, sCls[k], sCls[OP + k], sCls[2 * OP + k], sCls[3 * OP + k], hc,
                         gS, gT, g_gamma, g_t2, g_om, cpp[k], can[k], cpa[k]);
        }
        acc[i][k] = gS;
        acc[i][OP + k] = gT;
      }
      alpha[i] = 2.f * (g_gamma * ps.dgam + g_t2 * ps.dt2 + g_om * ps.dom);
    }
    // per-class scalars: fixed-order warp reduction, then this warp's shared-memory slot
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      float v0 = cpp[k], v1 = can[k], v2 = cpa[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      }
      if (lane == 0) {
        sRed[(warp * 3 + 0) * OP + k] += v0;
        sRed[(warp * 3 + 1) * OP + k] += v1;
        sRed[(warp * 3 + 2) * OP + k] += v2;
      }
    }
    // G planes for the weight-gradient GEMM
    if (p < HW) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        float* row = a.G + ((size_t)n * KP + k) * HW + p;
        if (VEC) *reinterpret_cast<float2*>(row) = make_float2(acc[0][k], acc[1][k]);
        else {
          row[0] = acc[0][k];
          if (p + 1 < HW) row[1] = acc[1][k];
        }
      }
    }
    // ---- pass 2: du = alpha*u + [gS gT] . Wt^T ----
    // Channels are independent here: process BWD_U of them per trip (loads issued first) and split every dot product
    // over 4 partial sums, so the FMA chains are 10 deep instead of 40 (the single-chain version was latency-bound).
    float* dbase = a.dfeat + (size_t)n * a.C * HW;
    for (int cb = 0; cb < a.C; cb += BWD_U) {
      float u[BWD_U][BWD_PIX];
#pragma unroll
      for (int j = 0; j < BWD_U; ++j) {
        const int ch = cb + j;
        u[j][0] = u[j][1] = 0.f;
        if (ch < a.C) {
          if (VEC) {
            if (p < HW) {
              const float2 t = __ldg(reinterpret_cast<const float2*>(base + (size_t)ch * HW + p));
              u[j][0] = t.x;
              u[j][1] = t.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < BWD_PIX; ++i)
              if (p + i < HW) u[j][i] = __ldg(base + (size_t)ch * HW + p + i);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < BWD_U; ++j) {
        const int ch = cb + j;
        if (ch >= a.C) break;
        float d[BWD_PIX][4];
#pragma unroll
        for (int i = 0; i < BWD_PIX; ++i) {
          d[i][0] = alpha[i] * u[j][i];
          d[i][1] = d[i][2] = d[i][3] = 0.f;
        }
        const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)ch * KP);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 w = w4[q];
#pragma unroll
          for (int i = 0; i < BWD_PIX; ++i) {
            d[i][0] = fmaf(acc[i][4 * q + 0], w.x, d[i][0]);
            d[i][1] = fmaf(acc[i][4 * q + 1], w.y, d[i][1]);
            d[i][2] = fmaf(acc[i][4 * q + 2], w.z, d[i][2]);
            d[i][3] = fmaf(acc[i][4 * q + 3], w.w, d[i][3]);
          }
        }
        const float d0 = (d[0][0] + d[0][1]) + (d[0][2] + d[0][3]);
        const float d1 = (d[1][0] + d[1][1]) + (d[1][2] + d[1][3]);
        if (VEC) {
          if (p < HW) __stcs(reinterpret_cast<float2*>(dbase + (size_t)ch * HW + p), make_float2(d0, d1));
        } else {
          if (p < HW) dbase[(size_t)ch * HW + p] = d0;
          if (p + 1 < HW) dbase[(size_t)ch * HW + p + 1] = d1;
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * OP; i += BWD_THREADS) {
    float s = 0.f;
    for (int w = 0; w < BWD_THREADS / 32; ++w) s += sRed[w * 3 * OP + i];
    a.cls_part[(size_t)blockIdx.x * 3 * OP + i] = s;
  }
}

// K4b: dW[k][c] = sum_pix G[k][pix] * u[c][pix].  Persistent CTAs, register accumulators, one partial per CTA.
// 256 threads: thread t owns channel c0+t and all 2*OP rows; the u/G strips of the NEXT unit are fetched into
// registers while the FMAs of the current unit run out of shared memory (software pipelining across units).
template <int OP>
__global__ void __launch_bounds__(DW_THREADS, 2) head_bwd_dw_kernel(const float* __restrict__ feat, const float* __restrict__ G,
                                                                  float* __restrict__ dw_part, int N, int C, int HW,
                                                                  int cblocks, long long total_units) {
  constexpr int KP = 2 * OP;
  constexpr int UPT = 256 * DW_PX / DW_THREADS;          // u elements fetched per thread per unit
  constexpr int GPT = (KP * DW_PX + DW_THREADS - 1) / DW_THREADS;
  __shared__ float sU[256][DW_PX + 1];
  __shared__ __align__(16) float sG[DW_PX][KP];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int strips = (HW + DW_PX - 1) / DW_PX;
  for (int cbk = 0; cbk < cblocks; ++cbk) {
    float acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.f;
    const int c0 = cbk * 256;
    float ru[UPT], rg[GPT];
    auto fetch = [&](long long unit) {
      const int n = (int)(unit / strips);
      const int p0 = (int)(unit - (long long)n * strips) * DW_PX;
#pragma unroll
      for (int e = 0; e < UPT; ++e) {
        const int r = wrp + e * (DW_THREADS / 32);       // channel row within the block; lane = pixel
        const int ch = c0 + r, px = p0 + lane;
        ru[e] = (ch < C && px < HW) ? __ldg(feat + ((size_t)n * C + ch) * HW + px) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < GPT; ++e) {
        const int i = tid + e * DW_THREADS;
        const int k = i / DW_PX, j = i - k * DW_PX;
        const int px = p0 + j;
        rg[e] = (i < KP * DW_PX && px < HW) ? __ldg(G + ((size_t)n * KP + k) * HW + px) : 0.f;
      }
    };
    long long unit = blockIdx.x;
    if (unit < total_units) fetch(unit);
    for (; unit < total_units; unit += gridDim.x) {
      __syncthreads();                                   // previous unit's FMAs are done with sU / sG
#pragma unroll
      for (int e = 0; e < UPT; ++e) sU[wrp + e * (DW_THREADS / 32)][lane] = ru[e];
#pragma unroll
      for (int e = 0; e < GPT; ++e) {
        const int i = tid + e * DW_THREADS;
        if (i < KP * DW_PX) sG[i % DW_PX][i / DW_PX] = rg[e];
      }
      __syncthreads();
      if (unit + gridDim.x < total_units) fetch(unit + gridDim.x);   // in flight during the FMAs below
#pragma unroll 4
      for (int j = 0; j < DW_PX; ++j) {
        const float u0 = sU[tid][j];
        const float4* g4 = reinterpret_cast<const float4*>(&sG[j][0]);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 g = g4[q];
          acc[4 * q + 0] = fmaf(g.x, u0, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(g.y, u0, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(g.z, u0, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(g.w, u0, acc[4 * q + 3]);
        }
      }
    }
    // partial [grid][KP][CP] with CP = cblocks*256
    float* out = dw_part + (size_t)blockIdx.x * KP * (cblocks * 256);
#pragma unroll
    for (int k = 0; k < KP; ++k) out[(size_t)k * (cblocks * 256) + c0 + tid] = acc[k];
    __syncthreads();
  }
}

// K4c: fixed-order reduction of the per-CTA partials + chain rule to dP, dA.  One block per class.
__global__ void head_bwd_finalize_kernel(const float* __restrict__ P, const float* __restrict__ A,
                                         const float* __restrict__ dw_part, int dw_grid, const float* __restrict__ cls_part,
                                         int cls_grid, float* __restrict__ dP, float* __restrict__ dA, int O, int OP, int C,
                                         int CP) {
  const int k = blockIdx.x;
  const int KP = 2 * OP;
  __shared__ double red[32];
  __shared__ double s_an2, s_dot;
  __shared__ float s_cls[3];
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int g = 0; g < cls_grid; ++g) s += cls_part[((size_t)g * 3 + threadIdx.x) * OP + k];
    s_cls[threadIdx.x] = s;
  }
  // |A_k|^2
  double aa = 0.0;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const double v = (double)A[(size_t)k * C + ch];
    aa += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) aa += __shfl_xor_sync(0xffffffffu, aa, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = aa;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    s_an2 = s;
  }
  __syncthreads();
  const double an = sqrt(s_an2);
  const double den = an > 1e-12 ? an : 1e-12;
  const float g_pp = s_cls[0], g_an = s_cls[1], g_pa = s_cls[2];
  // d a_hat (total) and its projection on a_hat
  double dot = 0.0;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float sT = 0.f;
    for (int g = 0; g < dw_grid; ++g) sT += dw_part[((size_t)g * KP + OP + k) * CP + ch];
    const double ahat = (double)A[(size_t)k * C + ch] / den;
    const double dah = (double)sT + (double)g_pa * (-(double)P[(size_t)k * C + ch]);
    dot += dah * ahat;
  }
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    s_dot = s;
  }
  __syncthreads();
  const double pdot = s_dot;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float sS = 0.f, sT = 0.f;
    for (int g = 0; g < dw_grid; ++g) {
      sS += dw_part[((size_t)g * KP + k) * CP + ch];
      sT += dw_part[((size_t)g * KP + OP + k) * CP + ch];
    }
    const double p = (double)P[(size_t)k * C + ch];
    const double ahat = (double)A[(size_t)k * C + ch] / den;
    const double dq = (double)sS + (double)g_pa * ahat;      // d/dq_k, q = -P
    const double dah = (double)sT + (double)g_pa * (-p);     // d/da_hat_k
    dP[(size_t)k * C + ch] = (float)(-dq + 2.0 * p * (double)g_pp);
    dA[(size_t)k * C + ch] = (float)((dah - pdot * ahat) / den + (double)g_an * ahat);
  }
}

}  // namespace halo

using namespace halo;

static int bwd_grids(int N, int HW, int* pix_grid, int* dw_grid) {
  const int tiles = ((HW + BWD_THREADS * BWD_PIX - 1) / (BWD_THREADS * BWD_PIX)) * N;
  int g = sm_count() * 2;
  if (g > tiles) g = tiles;
  *pix_grid = g;
  const long long units = (long long)N * ((HW + DW_PX - 1) / DW_PX);
  long long d = (long long)sm_count() * 3;
  if (d > units) d = units;
  *dw_grid = (int)d;
  return 0;
}

extern "C" size_t halo_head_bwd_workspace_bytes(int N, int C, int O, int H, int W) {
  if (N <= 0 || C <= 0 || O <= 0 || H <= 0 || W <= 0) return 0;
  const int OP = head_op_pad(O), KP = 2 * OP, CPAD = round_up(C, 4), CP = round_up(C, 256);
  int pg, dg;
  bwd_grids(N, H * W, &pg, &dg);
  size_t b = ((size_t)CPAD * KP + 4 * OP) * 4;          // packed parameters
  b = (b + 255) / 256 * 256;
  b += (size_t)N * (KP + 1) * H * W * 4;                 // G planes (two-kernel paths) or recomputed saved planes (streaming path)
  b = (b + 255) / 256 * 256;
  (void)pg;
  b += (size_t)sm_count() * 2 * 3 * OP * 4;              // class-scalar partials (one slot per CTA of either pixel pass)
  b = (b + 255) / 256 * 256;
  b += (size_t)dg * KP * CP * 4;                         // dW partials
  b = (b + 255) / 256 * 256;
  b += head_tc_pack_floats(O, C) * 4;                    // tensor-core parameter planes (MMA1)
  b = (b + 255) / 256 * 256;
  b += head_tc_pack_floats(O, C) * 4;                    // transposed planes (MMA2)
  return b + 256;
}

template <int OP>
static int launch_bwd(const BwdArgs& a, bool vec, size_t smem, int pix_grid, int dw_grid, const float* feat, float* dwp,
                      int cblocks, cudaStream_t st) {
  if (vec) {
    HALO_CUDA(cudaFuncSetAttribute(head_bwd_pix_kernel<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_bwd_pix_kernel<OP, true><<<pix_grid, BWD_THREADS, smem, st>>>(a);
  } else {
    HALO_CUDA(cudaFuncSetAttribute(head_bwd_pix_kernel<OP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_bwd_pix_kernel<OP, false><<<pix_grid, BWD_THREADS, smem, st>>>(a);
  }
  int rc = launch_status("head_bwd_pix_kernel");
  if (rc) return rc;
  head_bwd_dw_kernel<OP><<<dw_grid, DW_THREADS, 0, st>>>(feat, a.G, dwp, a.N, a.C, a.HW, cblocks,
                                                         (long long)a.N * ((a.HW + DW_PX - 1) / DW_PX));
  return launch_status("head_bwd_dw_kernel");
}

extern "C" int halo_head_bwd(const float* feat, const float* P, const float* A, float c, const float* dlogits,
                             const float* saved, float* dfeat, float* dP, float* dA, int N, int C, int O, int H, int W,
                             void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && P && A && dlogits && dfeat && dP && dA, "halo_head_bwd: NULL pointer");
  HALO_CHECK_ARG(N > 0 && C > 0 && O > 0 && H > 0 && W > 0 && c > 0.f, "halo_head_bwd: bad dims / curvature");
  if (O > 32) {
    set_error("halo_head_bwd: num_classes %d > 32 not compiled", O);
    return HALO_ERR_UNSUPPORTED;
  }
  const size_t need = halo_head_bwd_workspace_bytes(N, C, O, H, W);
  if (!ws || ws_bytes < need) {
    set_error("halo_head_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  const int OP = head_op_pad(O), KP = 2 * OP, CPAD = round_up(C, 4), CP = round_up(C, 256), HW = H * W;
  int pix_grid, dw_grid;
  bwd_grids(N, HW, &pix_grid, &dw_grid);
  unsigned char* w8 = (unsigned char*)ws;
  size_t off = 0;
  float* wpack = (float*)(w8 + off);
  off += ((size_t)CPAD * KP + 4 * OP) * 4; off = (off + 255) / 256 * 256;
  float* G = (float*)(w8 + off);
  off += (size_t)N * (KP + 1) * HW * 4; off = (off + 255) / 256 * 256;   // G planes, or the saved planes (KP + 1 rows)
  float* cls_part = (float*)(w8 + off);
  off += (size_t)sm_count() * 2 * 3 * OP * 4; off = (off + 255) / 256 * 256;
  float* dw_part = (float*)(w8 + off);
  off += (size_t)dw_grid * KP * CP * 4; off = (off + 255) / 256 * 256;
  float* wtc = (float*)(w8 + off);
  off += head_tc_pack_floats(O, C) * 4; off = (off + 255) / 256 * 256;
  float* w2 = (float*)(w8 + off);

  cudaStream_t st = (cudaStream_t)stream;
  head_pack_kernel<<<OP, 128, 0, st>>>(P, A, c, O, OP, C, CPAD, wpack, nullptr, 0);
  int rc = launch_status("head_pack_kernel");
  if (rc) return rc;
  const size_t smem = ((size_t)CPAD * KP + 4 * OP + (BWD_THREADS / 32) * 3 * OP) * 4;
  if (smem > 200 * 1024) {
    set_error("halo_head_bwd: C=%d x O=%d class parameters exceed the shared-memory tile", C, O);
    return HALO_ERR_UNSUPPORTED;
  }
  BwdArgs a;
  a.feat = feat; a.dlogits = dlogits; a.wpack = wpack; a.dfeat = dfeat; a.G = G; a.cls_part = cls_part;
  a.N = N; a.C = C; a.CPAD = CPAD; a.O = O; a.HW = HW;
  a.tiles_per_img = (HW + BWD_THREADS * BWD_PIX - 1) / (BWD_THREADS * BWD_PIX);
  a.total_tiles = a.tiles_per_img * N;
  a.hc = make_head_consts(c);
  const bool vec = (HW % 2 == 0) && ((uintptr_t)feat % 8 == 0) && ((uintptr_t)dfeat % 8 == 0);
  const int cblocks = CP / 256;
  int cls_grid = pix_grid;
  const char* force_cc = getenv("HALO_BWD_CUDA_CORE");  // parity tests pin the fp32 CUDA-core pixel pass
  const char* force_two = getenv("HALO_BWD_TWO_KERNEL");  // parity tests pin the round-1 two-kernel tensor-core path
  const bool pinned = (force_cc && force_cc[0] == '1') || (force_two && force_two[0] == '1');
  if (saved != nullptr && halo_head_saved_rows(C, O, H, W) == 0) {
    set_error("halo_head_bwd: saved planes given for a shape halo_head_saved_rows() does not support");
    return HALO_ERR_BAD_ARG;
  }
  if (!pinned && head_bwd_stream_supported(C, O, H, W, feat, dfeat) &&
      head_tc_supported(HALO_FEAT_TANGENT_F32, C, O, H, W, feat)) {
    // streaming backward (head_bwd_stream_tc.cu): the features are read once for du and dW.  Without saved planes the
    // contractions are recomputed first by the forward kernel in contraction-only mode (one more pass over the features).
    const float* sv = saved;
    if (sv == nullptr) {
      HeadArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.feat = feat; fa.ws = wpack; fa.saved = G;
      fa.N = N; fa.C = C; fa.CPAD = CPAD; fa.O = O; fa.HW = HW;
      fa.hc = make_head_consts(c);
      rc = head_fwd_tc_launch(fa, wpack, wtc, st);
      if (rc) return rc;
      sv = G;
    }
    dw_grid = cls_grid = head_bwd_stream_grid(N, HW);
    note_path(HALO_PATH_BWD_STREAM_TC | (saved ? 0 : HALO_PATH_BWD_RECOMPUTE), true);
    rc = head_bwd_stream_launch(feat, dlogits, sv, dfeat, dw_part, cls_part, wpack, w2, c, N, C, CPAD, O, H, W, CP, dw_grid, st);
  } else if (!(force_cc && force_cc[0] == '1') && head_bwd_tc_supported(C, O, H, W, feat, dfeat)) {
    // pixel pass on the tensor cores (head_bwd_tc.cu); the weight-gradient GEMM below is shared
    rc = head_pack_tc_launch(wpack, wtc, C, CPAD, O, st);
    if (rc) return rc;
    cls_grid = head_bwd_tc_grid(N, HW);
    note_path(HALO_PATH_BWD_PIX_TC, true);
    rc = head_bwd_tc_launch(feat, dlogits, dfeat, G, cls_part, wpack, wtc, w2, c, N, C, O, H, W, cls_grid, st);
    if (rc) return rc;
    const char* force_dw = getenv("HALO_BWD_DW_CUDA_CORE");  // parity tests pin the fp32 CUDA-core weight-gradient GEMM
    if (!(force_dw && force_dw[0] == '1') && head_bwd_dw_tc_supported(C, O, H, W, feat, G)) {
      dw_grid = head_bwd_dw_tc_grid(N, HW);
      note_path(HALO_PATH_BWD_DW_TC, false);
      rc = head_bwd_dw_tc_launch(feat, G, dw_part, N, C, O, H, W, CP, dw_grid, st);
    } else {
      note_path(HALO_PATH_BWD_DW_CUDA_CORE, false);
      if (!(force_dw && force_dw[0] == '1') && (long long)N * HW >= (1 << 16))
        warn_slow_path_once(2, "halo_head_bwd weight gradient N=%d C=%d O=%d H=%d W=%d runs on the fp32 CUDA cores", N, C, O, H, W);
      const long long units = (long long)N * ((HW + DW_PX - 1) / DW_PX);
      switch (OP) {
        case 4: head_bwd_dw_kernel<4><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 8: head_bwd_dw_kernel<8><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 12: head_bwd_dw_kernel<12><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 16: head_bwd_dw_kernel<16><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 20: head_bwd_dw_kernel<20><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 24: head_bwd_dw_kernel<24><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 28: head_bwd_dw_kernel<28><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        case 32: head_bwd_dw_kernel<32><<<dw_grid, DW_THREADS, 0, st>>>(feat, G, dw_part, N, C, HW, cblocks, units); break;
        default: set_error("halo_head_bwd: padded class count %d has no weight-gradient kernel", OP); return HALO_ERR_UNSUPPORTED;
      }
      rc = launch_status("head_bwd_dw_kernel");
    }
  } else {
    note_path(HALO_PATH_BWD_PIX_CUDA_CORE | HALO_PATH_BWD_DW_CUDA_CORE, true);
    if (!(force_cc && force_cc[0] == '1') && (long long)N * HW >= (1 << 16))
      warn_slow_path_once(1, "halo_head_bwd N=%d C=%d O=%d H=%d W=%d runs on the fp32 CUDA cores (tensor-core path needs "
                          "C %% 32 == 0, 64 <= C <= 256, O <= 24, H*W %% 4 == 0)", N, C, O, H, W);
    switch (OP) {
      case 4: rc = launch_bwd<4>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 8: rc = launch_bwd<8>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 12: rc = launch_bwd<12>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 16: rc = launch_bwd<16>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 20: rc = launch_bwd<20>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 24: rc = launch_bwd<24>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      case 28: rc = launch_bwd<28>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
      default: rc = launch_bwd<32>(a, vec, smem, pix_grid, dw_grid, feat, dw_part, cblocks, st); break;
    }
  }
  if (rc) return rc;
  head_bwd_finalize_kernel<<<O, 256, 0, st>>>(P, A, dw_part, dw_grid, cls_part, cls_grid, dP, dA, O, OP, C, CP);
  return launch_status("head_bwd_finalize_kernel");
}
