// Fused bilinear up-sampling (align_corners=True) in front of the region score  (SURVEY.md section 8f, row 1).
//
// The reference up-samples the O-channel logits (core/active/build.py:123-125) AND the C-channel float64 embedding
// (:132-135, 1-4 GiB per image) to the label size and only then takes softmax entropy / argmax and the radius of the
// interpolated embedding (floating_region.py:152,166,188,195; "radius is taken AFTER interpolating the embedding").
// Here every output pixel interpolates its 4 low-resolution neighbours on the fly and writes the three planes K2 needs
// (per-pixel uncertainty, label, radius): neither up-sampled tensor is ever materialised.
//   * logits: fp32 arithmetic with torch's own source-index rule (scale = (in-1)/(out-1) in float, src = scale*dst);
//   * embedding: float64 throughout, like the reference (1 - c|x|^2 near the ball boundary needs it).  Only the NORM of
//     the interpolated embedding is needed, and | sum_t w_t x_t |^2 = sum_tt' w_t w_t' <x_t, x_t'>: a pre-pass
//     (gram_lr_kernel) computes, per low-resolution pixel, its squared norm and the four dot products with / between its
//     right, lower and lower-right neighbours (plus the exp-map factor gamma for raw features, x_nb = gamma_nb * u_nb);
//     an output pixel then evaluates a 4x4 quadratic form from 14 cached doubles instead of interpolating C channels
//     (1 024 fp64 FMAs per output pixel at C = 256).  Algebraically identical to interpolate-then-norm.
#include "common.cuh"
#include "head_common.cuh"

namespace halo {

// Per low-resolution pixel p = (y, x), with the clamped neighbours r = (y, x1), d = (y1, x), q = (y1, x1) the output
// pixels of its cell use (x1 = x + (x < w-1), y1 = y + (y < h-1)):
//   plane 0  <e_p, e_p>     plane 1  <e_p, e_r>     plane 2  <e_p, e_d>     plane 3  <e_p, e_q>     plane 4  <e_r, e_d>
//   plane 5  gamma_p = tanh(min(s|u|,15), clip 1-1e-5) / (s|u|) for raw features, 1 for points already on the ball
// (e = the stored embedding values; the factor gamma is folded into the interpolation weights later).  ws: [N][6][h*w].
constexpr int UP_WS_PLANES = 6;
template <typename TE>
__global__ void gram_lr_kernel(const TE* __restrict__ emb, double* __restrict__ ws, int tangent, float c, int C, int eh, int ew,
                               long long total) {
  const int hw = eh * ew;
  const double s = sqrt((double)c);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long n = g / hw;
    const int p = (int)(g - n * hw);
    const int y = p / ew, x = p - y * ew;
    const int x1 = x + ((x < ew - 1) ? 1 : 0), y1 = y + ((y < eh - 1) ? 1 : 0);
    const int jr = y * ew + x1, jd = y1 * ew + x, jq = y1 * ew + x1;
    const TE* src = emb + (size_t)n * C * hw;
    double s0 = 0.0, sr = 0.0, sd = 0.0, sq = 0.0, sa = 0.0;
    for (int ch = 0; ch < C; ++ch) {
      const TE* row = src + (size_t)ch * hw;
      const double a = (double)row[p], b = (double)row[jr], d = (double)row[jd], q = (double)row[jq];
      s0 = fma(a, a, s0);
      sr = fma(a, b, sr);
      sd = fma(a, d, sd);
      sq = fma(a, q, sq);
      sa = fma(b, d, sa);
    }
    double* out = ws + (size_t)n * UP_WS_PLANES * hw + p;
    out[0] = s0;
    out[(size_t)hw] = sr;
    out[(size_t)2 * hw] = sd;
    out[(size_t)3 * hw] = sq;
    out[(size_t)4 * hw] = sa;
    double gam = 1.0;
    if (tangent) {
      const double nn = fmax(sqrt(s0), 1e-15);
      const double t = fmin(tanh(fmin(s * nn, 15.0)), 1.0 - 1e-5);
      gam = t / (s * nn);
    }
    out[(size_t)5 * hw] = gam;
  }
}

struct UpArgs {
  const float* logits;   // [N,O,h,w] or NULL
  const void* emb;       // [N,C,h,w] or NULL
  const double* gram;    // [N,6,eh*ew] from gram_lr_kernel, or NULL
  const uint8_t* gt;     // [N,H,W] or NULL
  float* pixunc;
  uint8_t* label;
  float* radius;
  float* stats;
  double* radius64;              // optional fp64 copy of the radius plane + per-image {min,max} ("hyper" bins)
  unsigned long long* stats64;
  int emb_kind, pixunc_mode, label_mode, norm_mode;
  int N, O, C, lh, lw, eh, ew, H, W;  // logits at lh x lw, embedding at eh x ew, outputs at H x W
  float c, inv_log19;
};

__global__ void __launch_bounds__(256) upsample_inputs_kernel(const UpArgs a) {
  const int n = blockIdx.z;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  float rmin = __int_as_float(0x7f800000), rmax = 0.f;
  unsigned long long lo64 = 0x7ff0000000000000ull, hi64 = 0ull;
  if (X < a.W) {
    // torch area_pixel_compute_scale / source_index with align_corners=True, float opmath
    const size_t pix = ((size_t)n * a.H + Y) * a.W + X;

    if (a.logits != nullptr && (a.pixunc != nullptr || a.label != nullptr)) {
      const int hw = a.lh * a.lw;
      const float sy = (a.H > 1) ? (float)(a.lh - 1) / (float)(a.H - 1) : 0.f;
      const float sx = (a.W > 1) ? (float)(a.lw - 1) / (float)(a.W - 1) : 0.f;
      const float fy = sy * (float)Y, fx = sx * (float)X;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + ((y0 < a.lh - 1) ? 1 : 0), x1 = x0 + ((x0 < a.lw - 1) ? 1 : 0);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float hy = 1.f - ly, hx = 1.f - lx;
      const int i00 = y0 * a.lw + x0, i01 = y0 * a.lw + x1, i10 = y1 * a.lw + x0, i11 = y1 * a.lw + x1;
      const float* L = a.logits + (size_t)n * a.O * hw;
      auto logit = [&](int k) {
        const float* q = L + (size_t)k * hw;
        return hy * (hx * __ldg(q + i00) + lx * __ldg(q + i01)) + ly * (hx * __ldg(q + i10) + lx * __ldg(q + i11));
      };
      float mx, Z = 0.f, ent = 0.f, eg = 0.f;
      int arg = 0;
      const int g8 = (a.gt != nullptr) ? a.gt[pix] : 255;
      if (a.O <= 32) {
        // the interpolated logits stay in registers (static indices under full unrolling): one interpolation per class
        float lg[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) lg[k] = (k < a.O) ? logit(k) : -3.0e38f;
        mx = lg[0];
#pragma unroll
        for (int k = 1; k < 32; ++k)
          if (k < a.O && lg[k] > mx) { mx = lg[k]; arg = k; }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          lg[k] = (k < a.O) ? __expf(lg[k] - mx) : 0.f;
          Z += lg[k];
        }
        const float iz0 = 1.f / Z;
        const int gtf0 = (g8 == 255) ? arg : g8;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float pk = lg[k] * iz0;
          if (k < a.O) ent -= pk * __logf(pk + 1e-6f);
          if (k == gtf0) eg = lg[k];
        }
      } else {
        mx = logit(0);
        for (int k = 1; k < a.O; ++k) {
          const float v = logit(k);
          if (v > mx) { mx = v; arg = k; }
        }
        for (int k = 0; k < a.O; ++k) Z += __expf(logit(k) - mx);
        const float iz0 = 1.f / Z;
        const int gtf0 = (g8 == 255) ? arg : g8;
        for (int k = 0; k < a.O; ++k) {
          const float pk = __expf(logit(k) - mx) * iz0;
          ent -= pk * __logf(pk + 1e-6f);
        }
        eg = (gtf0 < a.O) ? __expf(logit(gtf0) - mx) : 0.f;
      }
      const float iz = 1.f / Z;
      const int gtf = (g8 == 255) ? arg : g8;
      if (a.pixunc != nullptr) {
        const float v = (a.pixunc_mode == HALO_PIXUNC_ENTROPY) ? ent * a.inv_log19 : ((gtf < a.O) ? 1.f - eg * iz : 1.f);
        a.pixunc[pix] = v;
      }
      if (a.label != nullptr) a.label[pix] = (uint8_t)((a.label_mode == HALO_LABEL_GT_FILLED) ? gtf : arg);
    }

    if (a.emb != nullptr && a.radius != nullptr) {
      // float64 interpolation like the reference (its embedding is fp64): double source index and weights
      const int hw = a.eh * a.ew;
      const double dsy = (a.H > 1) ? (double)(a.eh - 1) / (double)(a.H - 1) : 0.0;
      const double dsx = (a.W > 1) ? (double)(a.ew - 1) / (double)(a.W - 1) : 0.0;
      const double dfy = dsy * (double)Y, dfx = dsx * (double)X;
      const int ey0 = (int)dfy, ex0 = (int)dfx;
      const int ey1 = ey0 + ((ey0 < a.eh - 1) ? 1 : 0), ex1 = ex0 + ((ex0 < a.ew - 1) ? 1 : 0);
      const double dly = dfy - (double)ey0, dlx = dfx - (double)ex0;
      double w00 = (1.0 - dly) * (1.0 - dlx), w01 = (1.0 - dly) * dlx, w10 = dly * (1.0 - dlx), w11 = dly * dlx;
      const int j00 = ey0 * a.ew + ex0, j01 = ey0 * a.ew + ex1, j10 = ey1 * a.ew + ex0, j11 = ey1 * a.ew + ex1;
      const double* G0 = a.gram + (size_t)n * UP_WS_PLANES * hw;
      const double *GR = G0 + hw, *GD = G0 + (size_t)2 * hw, *GQ = G0 + (size_t)3 * hw, *GA = G0 + (size_t)4 * hw,
                   *GAM = G0 + (size_t)5 * hw;
      // raw features: fold the exp-map factor of each neighbour into its weight (1 for points already on the ball)
      w00 *= GAM[j00]; w01 *= GAM[j01]; w10 *= GAM[j10]; w11 *= GAM[j11];
      // | w00 e00 + w01 e01 + w10 e10 + w11 e11 |^2 from the cell's Gram entries (j01 = right, j10 = lower, j11 = lower-right)
      double n2 = w00 * w00 * G0[j00] + w01 * w01 * G0[j01] + w10 * w10 * G0[j10] + w11 * w11 * G0[j11];
      n2 += 2.0 * (w00 * w01 * GR[j00] + w10 * w11 * GR[j10] + w00 * w10 * GD[j00] + w01 * w11 * GD[j01] +
                   w00 * w11 * GQ[j00] + w01 * w10 * GA[j00]);
      n2 = fmax(n2, 0.0);
      double r64;
      if (a.norm_mode == HALO_NORM_EUCLID) {
        r64 = sqrt(n2);
      } else {
        const double s = sqrt((double)a.c);
        const double t = fmin(s * sqrt(n2), 1.0 - 1e-7);  // geoopt artanh clamp (hyperbolic.py:83)
        r64 = (log1p(t) - log1p(-t)) / s;
      }
      const float r = (float)r64;
      a.radius[pix] = r;
      rmin = rmax = r;
      if (a.radius64 != nullptr) {
        a.radius64[pix] = r64;
        lo64 = hi64 = (unsigned long long)__double_as_longlong(r64);
      }
    }
  }
  if (a.stats != nullptr && a.radius != nullptr) {
    rmin = warp_min(rmin);
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(reinterpret_cast<int*>(a.stats + 4 * n + 0), __float_as_int(rmin));
      atomicMax(reinterpret_cast<int*>(a.stats + 4 * n + 1), __float_as_int(rmax));
    }
  }
  if (a.stats64 != nullptr && a.radius64 != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo64, o), h2 = __shfl_xor_sync(0xffffffffu, hi64, o);
      lo64 = l2 < lo64 ? l2 : lo64;
      hi64 = h2 > hi64 ? h2 : hi64;
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(a.stats64 + 2 * n + 0, lo64);
      atomicMax(a.stats64 + 2 * n + 1, hi64);
    }
  }
}

__global__ void up_stats_init_kernel(float* stats, unsigned long long* stats64, int N) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    if (stats != nullptr) {
      stats[4 * n + 0] = __int_as_float(0x7f800000);
      stats[4 * n + 1] = 0.f;
      stats[4 * n + 2] = 0.f;
      stats[4 * n + 3] = 0.f;
    }
    if (stats64 != nullptr) {
      stats64[2 * n + 0] = 0x7ff0000000000000ull;
      stats64[2 * n + 1] = 0ull;
    }
  }
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_upsample_workspace_bytes(int N, int h, int w) {
  return (N > 0 && h > 0 && w > 0) ? (size_t)N * UP_WS_PLANES * h * w * sizeof(double) : 0;
}

extern "C" int halo_upsample_score_inputs(const float* logits_lr, const void* emb_lr, int emb_kind, float c,
                                          const uint8_t* gt, int pixunc_mode, int label_mode, int norm_mode,
                                          float* pixunc, uint8_t* label, float* radius, float* stats, double* radius64,
                                          double* stats64, int N, int O, int C, int lh, int lw, int eh, int ew, int H, int W,
                                          void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG((radius64 == nullptr) == (stats64 == nullptr), "halo_upsample_score_inputs: radius64 and stats64 go together");
  HALO_CHECK_ARG(!radius64 || (emb_lr && radius), "halo_upsample_score_inputs: radius64 needs the embedding and the fp32 radius plane");
  HALO_CHECK_ARG(N > 0 && H > 0 && W > 0, "halo_upsample_score_inputs: bad dims");
  HALO_CHECK_ARG(!logits_lr || (lh > 0 && lw > 0), "halo_upsample_score_inputs: bad logits size");
  HALO_CHECK_ARG(!emb_lr || (eh > 0 && ew > 0), "halo_upsample_score_inputs: bad embedding size");
  HALO_CHECK_ARG(logits_lr || emb_lr, "halo_upsample_score_inputs: nothing to up-sample");
  HALO_CHECK_ARG(!logits_lr || O > 0, "halo_upsample_score_inputs: bad class count");
  HALO_CHECK_ARG(!emb_lr || (C > 0 && c > 0.f && radius), "halo_upsample_score_inputs: embedding needs C, c and a radius plane");
  HALO_CHECK_ARG(emb_kind >= 0 && emb_kind <= 2, "halo_upsample_score_inputs: bad emb_kind");
  HALO_CHECK_ARG(!((pixunc_mode == HALO_PIXUNC_ONE_MINUS_PGT && pixunc) || (label_mode == HALO_LABEL_GT_FILLED && label)) || gt,
                 "halo_upsample_score_inputs: gt required by the requested mode");
  HALO_CHECK_ARG(H <= 65535 && N <= 65535, "halo_upsample_score_inputs: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  UpArgs a;
  a.logits = logits_lr; a.emb = emb_lr; a.gram = nullptr; a.gt = gt; a.pixunc = pixunc; a.label = label; a.radius = radius;
  a.stats = stats; a.radius64 = radius64; a.stats64 = reinterpret_cast<unsigned long long*>(stats64); a.emb_kind = emb_kind; a.pixunc_mode = pixunc_mode; a.label_mode = label_mode; a.norm_mode = norm_mode;
  a.N = N; a.O = O; a.C = C; a.lh = lh; a.lw = lw; a.eh = eh; a.ew = ew; a.H = H; a.W = W; a.c = c; a.inv_log19 = (float)(1.0 / log(19.0));
  if (emb_lr) {
    const size_t need = halo_upsample_workspace_bytes(N, eh, ew);
    if (!ws || ws_bytes < need) {
      set_error("halo_upsample_score_inputs: workspace %zu < %zu bytes", ws_bytes, need);
      return HALO_ERR_WORKSPACE;
    }
    const long long total = (long long)N * eh * ew;
    long long blocks = (total + 127) / 128;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (emb_kind == HALO_FEAT_BALL_F64)
      gram_lr_kernel<double><<<(int)blocks, 128, 0, st>>>((const double*)emb_lr, (double*)ws, 0, c, C, eh, ew, total);
    else
      gram_lr_kernel<float><<<(int)blocks, 128, 0, st>>>((const float*)emb_lr, (double*)ws, emb_kind == HALO_FEAT_TANGENT_F32, c, C,
                                                        eh, ew, total);
    int rc = launch_status("gram_lr_kernel");
    if (rc) return rc;
    a.gram = (const double*)ws;
  }
  if (stats || stats64) {
    up_stats_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(stats, a.stats64, N);
    int rc = launch_status("up_stats_init_kernel");
    if (rc) return rc;
  }
  dim3 grid((W + 255) / 256, H, N);
  upsample_inputs_kernel<<<grid, 256, 0, st>>>(a);
  return launch_status("upsample_inputs_kernel");
}
