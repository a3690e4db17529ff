// K1 -- fused Poincare-ball classifier head, forward (CUDA-core fp32 contraction path).
//
// One pass over the features produces, per pixel: the 2*O+1 contractions <u,-p_k>, <u,a_k/|a_k|>, |u|^2,
// then in registers the expmap0/project scalars, the Mobius-addition algebra of HyperMLR, the asinh
// epilogue, the hyperbolic radius, the softmax entropy and the arg-max label.
// Reference semantics: core/utils/hyperbolic.py:28-39 (expmap), :74-83 (radius), :120-184 (MLR logits),
// core/active/floating_region.py:72-76,152,166 (entropy / argmax).  The closed form and its fp32-safe
// rewrites (sech^2 for 1-c|x|^2, the (1-c|p|^2)(1-c|x|^2)/D identity for 1-c|(-p)(+)x|^2) are
// derived in DESIGN.md section "K1".
#include "common.cuh"
#include "head_common.cuh"

namespace halo {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_PIX = 2;  // pixels per thread
constexpr int HEAD_U = 4;    // channels per software-pipeline stage

// ---- per-pixel epilogue ------------------------------------------------------------------------------
struct PixelScalars {
  float gamma;   // x = gamma * u
  float t2;      // c*|x|^2
  float omega;   // 1 - c*|x|^2
  float radius;  // (2/s) artanh(s|x|)
  float xnorm;   // |x|
};

// raw features: expmap0 + project fused (hyperbolic.py:37-38 with geoopt's fp64 eps 1e-5)
__device__ __forceinline__ PixelScalars tangent_scalars(float n2, const HeadConsts& hc) {
  PixelScalars ps;
  const float n = sqrtf(n2);
  const float sn = hc.s * n;
  const bool clipped = sn > hc.z_clip;  // tanh(min(sn,15)) > 1-1e-5
  const float z = fminf(sn, hc.z_clip);
  const float e = expf(-2.f * z);
  const float t = clipped ? hc.t_clip : tanhf(z);
  const float ope = 1.f + e;
  ps.omega = clipped ? hc.omega_clip : 4.f * e / (ope * ope);  // sech^2(z): never form 1 - t^2
  ps.gamma = t / (hc.s * fmaxf(n, 1e-15f));
  ps.t2 = t * t;
  ps.radius = hc.two_over_s * z;
  ps.xnorm = t * hc.inv_s;
  return ps;
}

// points already on the ball: |x|^2 arrives in double so that 1 - c|x|^2 keeps its leading digits
__device__ __forceinline__ PixelScalars ball_scalars(double n2, const HeadConsts& hc) {
  PixelScalars ps;
  const double cx = (double)hc.c * n2;
  ps.gamma = 1.f;
  ps.t2 = (float)cx;
  ps.omega = (float)(1.0 - cx);
  const double t = fmin(sqrt(cx), 1.0 - 1e-7);  // geoopt artanh clamp
  ps.radius = hc.two_over_s * (float)(0.5 * (log1p(t) - log1p(-t)));
  ps.xnorm = (float)sqrt(n2);
  return ps;
}

// HyperMLR logit for one class from the two contractions (hyperbolic.py:146-183)
__device__ __forceinline__ float mlr_logit(float S, float T, const PixelScalars& ps, float pp, float an, float pa,
                                           float Bk, const HeadConsts& hc) {
  const float px = ps.gamma * S;
  const float xa = ps.gamma * T;
  const float cpx2 = 2.f * hc.c * px;
  const float Anum = 1.f + cpx2 + ps.t2;                              // :150
  const float D = fmaxf(1.f + cpx2 + hc.c * ps.t2 * pp, 1e-12f);      // :152-153
  const float num = Bk * xa + Anum * pa;                              // D * <(-p)(+)x, a_hat>   (:175-177)
  const float bo = Bk * ps.omega;
  const float omc = bo / D;                                           // 1 - c*|(-p)(+)x|^2
  float arg;
  if (omc >= hc.om_max) {
    arg = hc.two_s * num / fmaxf(bo, 1e-12f * D);                     // inside the MLR ball: D cancels (:179-180)
  } else {
    const float m = fmaxf(1.f - omc, 0.f) * (1.f / hc.c);             // |(-p)(+)x|^2
    const float root = fmaxf(sqrtf(m), 1e-12f);
    arg = (num / D) * (hc.out_scale / root);                          // projected to maxnorm (:162-170)
  }
  return hc.two_over_s * an * asinhf(arg);                            // :181-183 (lambda_term = 2.0)
}

template <int OP, int PIX>
struct SoftmaxOut {
  float pixunc[PIX];
  int label[PIX];
};

// softmax entropy (floating_region.py:72-76) / 1-p[gt] (:77-83) and arg-max (:166) from logits in registers
template <int OP>
__device__ __forceinline__ void softmax_stats(const float (&l)[OP], int O, const HeadConsts& hc, int pixunc_mode,
                                              int label_mode, int gt, float& pixunc, int& label) {
  float mx = l[0];
  int arg = 0;
#pragma unroll
  for (int k = 1; k < OP; ++k)
    if (k < O && l[k] > mx) { mx = l[k]; arg = k; }
  float e[OP];
  float Z = 0.f;
#pragma unroll
  for (int k = 0; k < OP; ++k) {
    e[k] = (k < O) ? __expf(l[k] - mx) : 0.f;
    Z += e[k];
  }
  const float iz = 1.f / Z;
  const int gtf = (gt == 255) ? arg : gt;
  if (pixunc_mode == HALO_PIXUNC_ENTROPY) {
    float ent = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      const float p = e[k] * iz;
      if (k < O) ent -= p * __logf(p + 1e-6f);
    }
    pixunc = ent * hc.inv_log19;
  } else {
    float pg = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k)
      if (k == gtf) pg = e[k] * iz;
    pixunc = 1.f - pg;
  }
  label = (label_mode == HALO_LABEL_GT_FILLED) ? gtf : arg;
}

struct HeadArgs {
  const void* feat;
  const float* ws;
  float* logits;
  float* radius;
  float* pixunc;
  uint8_t* label;
  float* stats;
  const uint8_t* gt;
  int pixunc_mode, label_mode, norm_mode;
  int N, C, CPAD, O, HW;
  int tiles_per_img, total_tiles;
  HeadConsts hc;
};

template <int KIND>
struct FeatLoad;
template <>
struct FeatLoad<HALO_FEAT_TANGENT_F32> {
  typedef float T;
};
template <>
struct FeatLoad<HALO_FEAT_BALL_F32> {
  typedef float T;
};
template <>
struct FeatLoad<HALO_FEAT_BALL_F64> {
  typedef double T;
};

template <typename T, bool VEC>
__device__ __forceinline__ void load_pair(const T* row, int p, int HW, float (&v)[HEAD_PIX]) {
  if (VEC) {
    if (sizeof(T) == 4) {
      float2 t = (p < HW) ? __ldcs(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(row) + p))
                          : make_float2(0.f, 0.f);
      v[0] = t.x;
      v[1] = t.y;
    } else {
      double2 t = (p < HW) ? __ldcs(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(row) + p))
                           : make_double2(0.0, 0.0);
      v[0] = (float)t.x;
      v[1] = (float)t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < HEAD_PIX; ++i) v[i] = (p + i < HW) ? (float)__ldcs(row + p + i) : 0.f;
  }
}
// exact double copy of the pair for the |x|^2 accumulator of the BALL kinds
template <typename T, bool VEC>
__device__ __forceinline__ void load_pair_d(const T* row, int p, int HW, float (&v)[HEAD_PIX], double (&d)[HEAD_PIX]) {
#pragma unroll
  for (int i = 0; i < HEAD_PIX; ++i) {
    T t = (p + i < HW) ? __ldcs(row + p + i) : (T)0;
    d[i] = (double)t;
    v[i] = (float)t;
  }
}

template <int OP, int KIND, bool VEC>
__global__ void __launch_bounds__(HEAD_THREADS, 2) head_fwd_kernel(const HeadArgs a) {
  typedef typename FeatLoad<KIND>::T T;
  constexpr int KP = 2 * OP;
  constexpr bool TANGENT = (KIND == HALO_FEAT_TANGENT_F32);
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                           // [CPAD][KP]
  float* sCls = smem + (size_t)a.CPAD * KP;   // [4][OP]
  {
    const int n4 = (a.CPAD * KP + 4 * OP) / 4;
    const float4* src = reinterpret_cast<const float4*>(a.ws);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < n4; i += HEAD_THREADS) dst[i] = src[i];
  }
  __syncthreads();

  const HeadConsts hc = a.hc;
  const int HW = a.HW;
  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_img;
    const int p = (tile - n * a.tiles_per_img) * (HEAD_THREADS * HEAD_PIX) + threadIdx.x * HEAD_PIX;
    const T* base = reinterpret_cast<const T*>(a.feat) + (size_t)n * a.C * HW;

    float acc[HEAD_PIX][KP];
    float n2[HEAD_PIX];
    double n2d[HEAD_PIX];
#pragma unroll
    for (int i = 0; i < HEAD_PIX; ++i) {
      n2[i] = 0.f;
      n2d[i] = 0.0;
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[i][k] = 0.f;
    }

    if (TANGENT) {
      float cur[HEAD_U][HEAD_PIX], nxt[HEAD_U][HEAD_PIX];
#pragma unroll
      for (int j = 0; j < HEAD_U; ++j) {
        if (j < a.C) load_pair<T, VEC>(base + (size_t)j * HW, p, HW, cur[j]);
        else cur[j][0] = cur[j][1] = 0.f;
      }
      for (int cb = 0; cb < a.CPAD; cb += HEAD_U) {
#pragma unroll
        for (int j = 0; j < HEAD_U; ++j) {
          const int ch = cb + HEAD_U + j;
          if (ch < a.C) load_pair<T, VEC>(base + (size_t)ch * HW, p, HW, nxt[j]);
          else nxt[j][0] = nxt[j][1] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < HEAD_U; ++j) {
          const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)(cb + j) * KP);
#pragma unroll
          for (int i = 0; i < HEAD_PIX; ++i) n2[i] = fmaf(cur[j][i], cur[j][i], n2[i]);
#pragma unroll
          for (int q = 0; q < KP / 4; ++q) {
            const float4 w = w4[q];
#pragma unroll
            for (int i = 0; i < HEAD_PIX; ++i) {
              acc[i][4 * q + 0] = fmaf(cur[j][i], w.x, acc[i][4 * q + 0]);
              acc[i][4 * q + 1] = fmaf(cur[j][i], w.y, acc[i][4 * q + 1]);
              acc[i][4 * q + 2] = fmaf(cur[j][i], w.z, acc[i][4 * q + 2]);
              acc[i][4 * q + 3] = fmaf(cur[j][i], w.w, acc[i][4 * q + 3]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < HEAD_U; ++j)
#pragma unroll
          for (int i = 0; i < HEAD_PIX; ++i) cur[j][i] = nxt[j][i];
      }
    } else {
      // compatibility path (points already on the ball): |x|^2 in double, one channel at a time
      for (int ch = 0; ch < a.C; ++ch) {
        float v[HEAD_PIX];
        double d[HEAD_PIX];
        load_pair_d<T, VEC>(base + (size_t)ch * HW, p, HW, v, d);
        const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)ch * KP);
#pragma unroll
        for (int i = 0; i < HEAD_PIX; ++i) n2d[i] = fma(d[i], d[i], n2d[i]);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 w = w4[q];
#pragma unroll
          for (int i = 0; i < HEAD_PIX; ++i) {
            acc[i][4 * q + 0] = fmaf(v[i], w.x, acc[i][4 * q + 0]);
            acc[i][4 * q + 1] = fmaf(v[i], w.y, acc[i][4 * q + 1]);
            acc[i][4 * q + 2] = fmaf(v[i], w.z, acc[i][4 * q + 2]);
            acc[i][4 * q + 3] = fmaf(v[i], w.w, acc[i][4 * q + 3]);
          }
        }
      }
    }

    // ---- epilogue, all in registers ----
    float rmin = __int_as_float(0x7f800000), rmax = 0.f;
    float out_rad[HEAD_PIX], out_unc[HEAD_PIX];
    int out_lab[HEAD_PIX];
#pragma unroll
    for (int i = 0; i < HEAD_PIX; ++i) {
      const PixelScalars ps = TANGENT ? tangent_scalars(n2[i], hc) : ball_scalars(n2d[i], hc);
      float l[OP];
#pragma unroll
      for (int k = 0; k < OP; ++k)
        l[k] = mlr_logit(acc[i][k], acc[i][OP + k], ps, sCls[k], sCls[OP + k], sCls[2 * OP + k], sCls[3 * OP + k], hc);
#pragma unroll
      for (int k = 0; k < OP; ++k) acc[i][k] = l[k];
      const float r = (a.norm_mode == HALO_NORM_EUCLID) ? ps.xnorm : ps.radius;
      out_rad[i] = r;
      if (p + i < HW) { rmin = fminf(rmin, r); rmax = fmaxf(rmax, r); }
      out_unc[i] = 0.f;
      out_lab[i] = 0;
      if (a.pixunc != nullptr || a.label != nullptr) {
        int g = 255;
        if (a.gt != nullptr && p + i < HW) g = a.gt[(size_t)n * HW + p + i];
        softmax_stats<OP>(l, a.O, hc, a.pixunc_mode, a.label_mode, g, out_unc[i], out_lab[i]);
      }
    }

    const size_t pix0 = (size_t)n * HW + p;
    if (VEC) {
      if (p < HW) {
        if (a.logits != nullptr) {
#pragma unroll
          for (int k = 0; k < OP; ++k)
            if (k < a.O)
              __stcs(reinterpret_cast<float2*>(a.logits + ((size_t)n * a.O + k) * HW + p),
                     make_float2(acc[0][k], acc[1][k]));
        }
        if (a.radius != nullptr) __stcs(reinterpret_cast<float2*>(a.radius + pix0), make_float2(out_rad[0], out_rad[1]));
        if (a.pixunc != nullptr) __stcs(reinterpret_cast<float2*>(a.pixunc + pix0), make_float2(out_unc[0], out_unc[1]));
        if (a.label != nullptr)
          *reinterpret_cast<uchar2*>(a.label + pix0) = make_uchar2((unsigned char)out_lab[0], (unsigned char)out_lab[1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < HEAD_PIX; ++i) {
        if (p + i < HW) {
          if (a.logits != nullptr) {
#pragma unroll
            for (int k = 0; k < OP; ++k)
              if (k < a.O) a.logits[((size_t)n * a.O + k) * HW + p + i] = acc[i][k];
          }
          if (a.radius != nullptr) a.radius[pix0 + i] = out_rad[i];
          if (a.pixunc != nullptr) a.pixunc[pix0 + i] = out_unc[i];
          if (a.label != nullptr) a.label[pix0 + i] = (uint8_t)out_lab[i];
        }
      }
    }
    if (a.stats != nullptr) {
      rmin = warp_min(rmin);
      rmax = warp_max(rmax);
      if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<int*>(a.stats + 4 * n + 0), __float_as_int(rmin));
        atomicMax(reinterpret_cast<int*>(a.stats + 4 * n + 1), __float_as_int(rmax));
      }
    }
  }
}

template <int OP, int KIND>
static int launch_head(const HeadArgs& a, bool vec, size_t smem, int grid, cudaStream_t st) {
  if constexpr (KIND == HALO_FEAT_TANGENT_F32) {
    if (vec) {
      HALO_CUDA(cudaFuncSetAttribute(head_fwd_kernel<OP, KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      head_fwd_kernel<OP, KIND, true><<<grid, HEAD_THREADS, smem, st>>>(a);
      return launch_status("head_fwd_kernel");
    }
  }
  // BALL kinds (compatibility path) and unaligned / odd-sized planes use the scalar loader
  HALO_CUDA(cudaFuncSetAttribute(head_fwd_kernel<OP, KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_fwd_kernel<OP, KIND, false><<<grid, HEAD_THREADS, smem, st>>>(a);
  return launch_status("head_fwd_kernel");
}

template <int OP>
static int launch_head_kind(const HeadArgs& a, int kind, bool vec, size_t smem, int grid, cudaStream_t st) {
  switch (kind) {
    case HALO_FEAT_TANGENT_F32: return launch_head<OP, HALO_FEAT_TANGENT_F32>(a, vec, smem, grid, st);
    case HALO_FEAT_BALL_F32: return launch_head<OP, HALO_FEAT_BALL_F32>(a, vec, smem, grid, st);
    default: return launch_head<OP, HALO_FEAT_BALL_F64>(a, vec, smem, grid, st);
  }
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_head_workspace_bytes(int O, int C) {
  if (O <= 0 || C <= 0) return 0;
  const int OP = head_op_pad(O), CPAD = round_up(C, HEAD_U);
  return ((size_t)CPAD * 2 * OP + 4 * OP) * sizeof(float);
}

extern "C" int halo_head_fwd(const void* feat, int feat_kind, const float* P, const float* A, float c, float* logits,
                             float* radius, float* pixunc, uint8_t* label, float* stats, const uint8_t* gt,
                             int pixunc_mode, int label_mode, int norm_mode, int N, int C, int O, int H, int W,
                             void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && P && A, "halo_head_fwd: feat/P/A must not be NULL");
  HALO_CHECK_ARG(N > 0 && C > 0 && O > 0 && H > 0 && W > 0, "halo_head_fwd: non-positive dims N=%d C=%d O=%d H=%d W=%d", N, C, O, H, W);
  HALO_CHECK_ARG(c > 0.f, "halo_head_fwd: curvature c must be > 0 (got %g)", (double)c);
  HALO_CHECK_ARG(feat_kind >= 0 && feat_kind <= 2, "halo_head_fwd: bad feat_kind %d", feat_kind);
  HALO_CHECK_ARG(pixunc_mode >= 0 && pixunc_mode <= 1 && label_mode >= 0 && label_mode <= 1 && norm_mode >= 0 && norm_mode <= 1,
                 "halo_head_fwd: bad mode");
  HALO_CHECK_ARG(!((pixunc_mode == HALO_PIXUNC_ONE_MINUS_PGT && pixunc) || (label_mode == HALO_LABEL_GT_FILLED && label)) || gt,
                 "halo_head_fwd: gt required by the requested pixunc/label mode");
  HALO_CHECK_ARG((long long)H * W < (1LL << 30), "halo_head_fwd: image too large");
  if (O > 32) {
    set_error("halo_head_fwd: num_classes %d > 32 not compiled", O);
    return HALO_ERR_UNSUPPORTED;
  }
  const size_t need = halo_head_workspace_bytes(O, C);
  if (!ws || ws_bytes < need) {
    set_error("halo_head_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  const int OP = head_op_pad(O), CPAD = round_up(C, HEAD_U);
  const size_t smem = need;
  if (smem > 200 * 1024) {
    set_error("halo_head_fwd: C=%d x O=%d class parameters (%zu B) exceed the shared-memory tile", C, O, smem);
    return HALO_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  head_pack_kernel<<<OP, 128, 0, st>>>(P, A, c, O, OP, C, CPAD, (float*)ws, stats, N);
  int rc = launch_status("head_pack_kernel");
  if (rc) return rc;

  HeadArgs a;
  a.feat = feat; a.ws = (const float*)ws; a.logits = logits; a.radius = radius; a.pixunc = pixunc; a.label = label;
  a.stats = stats; a.gt = gt; a.pixunc_mode = pixunc_mode; a.label_mode = label_mode; a.norm_mode = norm_mode;
  a.N = N; a.C = C; a.CPAD = CPAD; a.O = O; a.HW = H * W;
  a.tiles_per_img = (a.HW + HEAD_THREADS * HEAD_PIX - 1) / (HEAD_THREADS * HEAD_PIX);
  a.total_tiles = a.tiles_per_img * N;
  a.hc = make_head_consts(c);
  const size_t esz = (feat_kind == HALO_FEAT_BALL_F64) ? 8 : 4;
  bool vec = (a.HW % 2 == 0) && (((uintptr_t)feat) % (2 * esz) == 0);
  if (logits && ((uintptr_t)logits % 8)) vec = false;
  if (radius && ((uintptr_t)radius % 8)) vec = false;
  if (pixunc && ((uintptr_t)pixunc % 8)) vec = false;
  if (label && ((uintptr_t)label % 2)) vec = false;
  const int per_sm = (smem <= 100 * 1024) ? 2 : 1;
  int grid = sm_count() * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  switch (OP) {
    case 4: return launch_head_kind<4>(a, feat_kind, vec, smem, grid, st);
    case 8: return launch_head_kind<8>(a, feat_kind, vec, smem, grid, st);
    case 12: return launch_head_kind<12>(a, feat_kind, vec, smem, grid, st);
    case 16: return launch_head_kind<16>(a, feat_kind, vec, smem, grid, st);
    case 20: return launch_head_kind<20>(a, feat_kind, vec, smem, grid, st);
    case 24: return launch_head_kind<24>(a, feat_kind, vec, smem, grid, st);
    case 28: return launch_head_kind<28>(a, feat_kind, vec, smem, grid, st);
    default: return launch_head_kind<32>(a, feat_kind, vec, smem, grid, st);
  }
}
