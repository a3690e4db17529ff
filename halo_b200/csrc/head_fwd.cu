// K1 -- fused Poincare-ball classifier head, forward (CUDA-core fp32 contraction path).
//
// One pass over the features produces, per pixel: the 2*O+1 contractions <u,-p_k>, <u,a_k/|a_k|>, |u|^2,
// then in registers the expmap0/project scalars, the Mobius-addition algebra of HyperMLR, the asinh
// epilogue, the hyperbolic radius, the softmax entropy and the arg-max label.
// Reference semantics: core/utils/hyperbolic.py:28-39 (expmap), :74-83 (radius), :120-184 (MLR logits),
// core/active/floating_region.py:72-76,152,166 (entropy / argmax).  The closed form and its fp32-safe
// rewrites (sech^2 for 1-c|x|^2, the (1-c|p|^2)(1-c|x|^2)/D identity for 1-c|(-p)(+)x|^2) are
// derived in DESIGN.md section "K1".
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"

namespace halo {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_PIX = 2;  // pixels per thread
constexpr int HEAD_U = 4;    // channels per software-pipeline stage

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
      // software pipeline, two register buffers used alternately (no buffer copies): while the FMAs of
      // HEAD_U channels issue, the loads of the next HEAD_U channels are in flight
      float bufA[HEAD_U][HEAD_PIX], bufB[HEAD_U][HEAD_PIX];
      auto load_stage = [&](float (&dst)[HEAD_U][HEAD_PIX], int cb) {
#pragma unroll
        for (int j = 0; j < HEAD_U; ++j) {
          const int ch = cb + j;
          if (ch < a.C) load_pair<T, VEC>(base + (size_t)ch * HW, p, HW, dst[j]);
          else dst[j][0] = dst[j][1] = 0.f;
        }
      };
      auto fma_stage = [&](const float (&src)[HEAD_U][HEAD_PIX], int cb) {
#pragma unroll
        for (int j = 0; j < HEAD_U; ++j) {
          const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)(cb + j) * KP);
#pragma unroll
          for (int i = 0; i < HEAD_PIX; ++i) n2[i] = fmaf(src[j][i], src[j][i], n2[i]);
#pragma unroll
          for (int q = 0; q < KP / 4; ++q) {
            const float4 w = w4[q];
#pragma unroll
            for (int i = 0; i < HEAD_PIX; ++i) {
              acc[i][4 * q + 0] = fmaf(src[j][i], w.x, acc[i][4 * q + 0]);
              acc[i][4 * q + 1] = fmaf(src[j][i], w.y, acc[i][4 * q + 1]);
              acc[i][4 * q + 2] = fmaf(src[j][i], w.z, acc[i][4 * q + 2]);
              acc[i][4 * q + 3] = fmaf(src[j][i], w.w, acc[i][4 * q + 3]);
            }
          }
        }
      };
      load_stage(bufA, 0);
      for (int cb = 0; cb < a.CPAD; cb += 2 * HEAD_U) {
        load_stage(bufB, cb + HEAD_U);
        fma_stage(bufA, cb);
        load_stage(bufA, cb + 2 * HEAD_U);
        if (cb + HEAD_U < a.CPAD) fma_stage(bufB, cb + HEAD_U);
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

    // ---- epilogue, all in registers.  The pixel loop is deliberately NOT unrolled: the per-pixel code
    // (OP classes of Mobius algebra + softmax) is ~1.5k instructions and two copies thrash the instruction cache.
    float rmin = __int_as_float(0x7f800000), rmax = 0.f;
    float out_rad[HEAD_PIX], out_unc[HEAD_PIX];
    int out_lab[HEAD_PIX];
#pragma unroll
    for (int i = 0; i < HEAD_PIX; ++i) { out_rad[i] = 0.f; out_unc[i] = 0.f; out_lab[i] = 0; }
#pragma unroll 1
    for (int i = 0; i < HEAD_PIX; ++i) {
      const bool second = (i != 0);
      const PixelScalars ps = TANGENT ? tangent_scalars(second ? n2[1] : n2[0], hc)
                                      : ball_scalars(second ? n2d[1] : n2d[0], hc);
      float l[OP];
#pragma unroll
      for (int k = 0; k < OP; ++k) {
        const float S = second ? acc[1][k] : acc[0][k];
        const float T = second ? acc[1][OP + k] : acc[0][OP + k];
        l[k] = mlr_logit(S, T, ps, sCls[k], sCls[OP + k], sCls[2 * OP + k], sCls[3 * OP + k], hc);
      }
#ifdef HALO_TC_VARIANTS
      if (a.debug_raw) {
#pragma unroll
        for (int k = 0; k < OP; ++k) l[k] = second ? acc[1][OP + k] : acc[0][OP + k];
      }
#endif
#pragma unroll
      for (int k = 0; k < OP; ++k) {
        if (second) acc[1][k] = l[k];
        else acc[0][k] = l[k];
      }
      const float r = (a.norm_mode == HALO_NORM_EUCLID) ? ps.xnorm : ps.radius;
      if (p + i < HW) { rmin = fminf(rmin, r); rmax = fmaxf(rmax, r); }
      float unc = 0.f;
      int lab = 0;
      if (a.pixunc != nullptr || a.label != nullptr) {
        int g = 255;
        if (a.gt != nullptr && p + i < HW) g = a.gt[(size_t)n * HW + p + i];
        softmax_stats<OP>(l, a.O, hc, a.pixunc_mode, a.label_mode, g, unc, lab);
      }
      if (second) { out_rad[1] = r; out_unc[1] = unc; out_lab[1] = lab; }
      else { out_rad[0] = r; out_unc[0] = unc; out_lab[0] = lab; }
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
  size_t std_bytes = ((size_t)CPAD * 2 * OP + 4 * OP) * sizeof(float);
  std_bytes = (std_bytes + 255) / 256 * 256;
  return std_bytes + head_tc_pack_floats(O, C) * sizeof(float);
}

extern "C" int halo_head_saved_rows(int C, int O, int H, int W) {
  if (C <= 0 || O <= 0 || H <= 0 || W <= 0) return 0;
  // shapes both the tensor-core forward and the streaming backward take (pointer alignment is checked per call)
  const void* aligned = reinterpret_cast<const void*>(uintptr_t(256));
  if (!head_tc_shape_ok(HALO_FEAT_TANGENT_F32, C, O, H, W, aligned)) return 0;
  if (!head_bwd_stream_shape_ok(C, O, H, W, aligned, aligned)) return 0;
  return 2 * head_op_pad(O) + 1;
}

extern "C" int halo_head_fwd(const void* feat, int feat_kind, const float* P, const float* A, float c, float* logits,
                             float* radius, float* pixunc, uint8_t* label, float* stats, float* saved, const uint8_t* gt,
                             int pixunc_mode, int label_mode, int norm_mode, int N, int C, int O, int H, int W,
                             void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && P && A, "halo_head_fwd: feat/P/A must not be NULL");
  HALO_CHECK_ARG(N > 0 && C > 0 && O > 0 && H > 0 && W > 0, "halo_head_fwd: non-positive dims N=%d C=%d O=%d H=%d W=%d", N, C, O, H, W);
  HALO_CHECK_ARG(c > 0.f, "halo_head_fwd: curvature c must be > 0 (got %g)", (double)c);
  const bool no_tc = (feat_kind & HALO_FEAT_FLAG_NO_TENSOR_CORE) != 0;
  feat_kind &= 0xff;
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
  const size_t smem = ((size_t)CPAD * 2 * OP + 4 * OP) * sizeof(float);
  const bool use_tc = !no_tc && head_tc_supported(feat_kind, C, O, H, W, feat);
  if (saved != nullptr && (!use_tc || halo_head_saved_rows(C, O, H, W) == 0)) {
    set_error("halo_head_fwd: saved planes are written by the tensor-core path only (halo_head_saved_rows(C=%d,O=%d,H=%d,W=%d) "
              "is 0, or the features are unaligned / not raw fp32)", C, O, H, W);
    return HALO_ERR_UNSUPPORTED;
  }
  if (!use_tc && smem > 200 * 1024) {
    set_error("halo_head_fwd: C=%d x O=%d class parameters (%zu B) exceed the shared-memory tile", C, O, smem);
    return HALO_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  head_pack_kernel<<<OP, 128, 0, st>>>(P, A, c, O, OP, C, CPAD, (float*)ws, stats, N);
  int rc = launch_status("head_pack_kernel");
  if (rc) return rc;

  HeadArgs a;
  a.feat = feat; a.ws = (const float*)ws; a.logits = logits; a.radius = radius; a.pixunc = pixunc; a.label = label;
  a.stats = stats; a.saved = saved; a.gt = gt; a.pixunc_mode = pixunc_mode; a.label_mode = label_mode; a.norm_mode = norm_mode;
  a.N = N; a.C = C; a.CPAD = CPAD; a.O = O; a.HW = H * W;
  a.tiles_per_img = (a.HW + HEAD_THREADS * HEAD_PIX - 1) / (HEAD_THREADS * HEAD_PIX);
  a.total_tiles = a.tiles_per_img * N;
  a.hc = make_head_consts(c);
  a.debug_raw = 0;
  a.tc_variant = 0;
#ifdef HALO_TC_VARIANTS
  if (const char* e = getenv("HALO_TC_DEBUG")) {  // numerics probe builds only: "<variant>[r]"
    a.tc_variant = atoi(e);
    a.debug_raw = strchr(e, 'r') != nullptr;
  }
#endif
  if (use_tc) {
    float* wtc = (float*)((unsigned char*)ws + (smem + 255) / 256 * 256);
    note_path(HALO_PATH_FWD_TC, true);
    return head_fwd_tc_launch(a, (const float*)ws, wtc, st);
  }
  note_path(HALO_PATH_FWD_CUDA_CORE, true);
  if (!no_tc && feat_kind == HALO_FEAT_TANGENT_F32 && (long long)N * H * W >= (1 << 16))
    warn_slow_path_once(0, "halo_head_fwd N=%d C=%d O=%d H=%d W=%d runs on the fp32 CUDA cores (tensor-core path needs "
                        "C %% 32 == 0, C <= 256, H*W %% 4 == 0, 16-byte aligned features)", N, C, O, H, W);
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
