// Fused bilinear up-sampling (align_corners=True) in front of the region score  (SURVEY.md section 8f, row 1).
//
// The reference up-samples the O-channel logits (core/active/build.py:123-125) AND the C-channel float64 embedding
// (:132-135, 1-4 GiB per image) to the label size and only then takes softmax entropy / argmax and the radius of the
// interpolated embedding (floating_region.py:152,166,188,195; "radius is taken AFTER interpolating the embedding").
// Here every output pixel interpolates its 4 low-resolution neighbours on the fly and writes the three planes K2 needs
// (per-pixel uncertainty, label, radius): neither up-sampled tensor is ever materialised.
//   * logits: fp32 arithmetic with torch's own source-index rule (scale = (in-1)/(out-1) in float, src = scale*dst);
//   * embedding: float64 throughout, like the reference (1 - c|x|^2 near the ball boundary needs it); for raw features the
//     exp-map factor gamma of each low-resolution pixel is precomputed once (gamma_lr_kernel), x_nb = gamma_nb * u_nb.
#include "common.cuh"
#include "head_common.cuh"

namespace halo {

// gamma(u) = tanh(min(s|u|,15), clip 1-1e-5) / (s|u|) per low-resolution pixel, in double
__global__ void gamma_lr_kernel(const float* __restrict__ u, double* __restrict__ gamma, float c, int C, int hw, long long total) {
  const double s = sqrt((double)c);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long n = g / hw;
    const int p = (int)(g - n * hw);
    const float* src = u + (size_t)n * C * hw + p;
    double n2 = 0.0;
    for (int ch = 0; ch < C; ++ch) {
      const double v = (double)src[(size_t)ch * hw];
      n2 = fma(v, v, n2);
    }
    const double nn = fmax(sqrt(n2), 1e-15);
    const double t = fmin(tanh(fmin(s * nn, 15.0)), 1.0 - 1e-5);
    gamma[g] = t / (s * nn);
  }
}

struct UpArgs {
  const float* logits;   // [N,O,h,w] or NULL
  const void* emb;       // [N,C,h,w] or NULL
  const double* gamma;   // [N,h,w] (tangent kind) or NULL
  const uint8_t* gt;     // [N,H,W] or NULL
  float* pixunc;
  uint8_t* label;
  float* radius;
  float* stats;
  int emb_kind, pixunc_mode, label_mode, norm_mode;
  int N, O, C, lh, lw, eh, ew, H, W;  // logits at lh x lw, embedding at eh x ew, outputs at H x W
  float c, inv_log19;
};

template <typename TE>
__global__ void __launch_bounds__(256) upsample_inputs_kernel(const UpArgs a) {
  const int n = blockIdx.z;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  float rmin = __int_as_float(0x7f800000), rmax = 0.f;
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
      float mx = logit(0);
      int arg = 0;
      for (int k = 1; k < a.O; ++k) {
        const float v = logit(k);
        if (v > mx) { mx = v; arg = k; }
      }
      float Z = 0.f;
      for (int k = 0; k < a.O; ++k) Z += __expf(logit(k) - mx);
      const float iz = 1.f / Z;
      const int g8 = (a.gt != nullptr) ? a.gt[pix] : 255;
      const int gtf = (g8 == 255) ? arg : g8;
      if (a.pixunc != nullptr) {
        float v;
        if (a.pixunc_mode == HALO_PIXUNC_ENTROPY) {
          float ent = 0.f;
          for (int k = 0; k < a.O; ++k) {
            const float pk = __expf(logit(k) - mx) * iz;
            ent -= pk * __logf(pk + 1e-6f);
          }
          v = ent * a.inv_log19;
        } else {
          v = (gtf < a.O) ? 1.f - __expf(logit(gtf) - mx) * iz : 1.f;
        }
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
      if (a.gamma != nullptr) {  // raw features: fold the exp-map factor of each neighbour into its weight
        const double* G = a.gamma + (size_t)n * hw;
        w00 *= G[j00]; w01 *= G[j01]; w10 *= G[j10]; w11 *= G[j11];
      }
      const TE* E = reinterpret_cast<const TE*>(a.emb) + (size_t)n * a.C * hw;
      double n2 = 0.0;
      for (int ch = 0; ch < a.C; ++ch) {
        const TE* q = E + (size_t)ch * hw;
        const double v = w00 * (double)q[j00] + w01 * (double)q[j01] + w10 * (double)q[j10] + w11 * (double)q[j11];
        n2 = fma(v, v, n2);
      }
      float r;
      if (a.norm_mode == HALO_NORM_EUCLID) {
        r = (float)sqrt(n2);
      } else {
        const double s = sqrt((double)a.c);
        const double t = fmin(s * sqrt(n2), 1.0 - 1e-7);  // geoopt artanh clamp (hyperbolic.py:83)
        r = (float)((log1p(t) - log1p(-t)) / s);
      }
      a.radius[pix] = r;
      rmin = rmax = r;
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
}

__global__ void up_stats_init_kernel(float* stats, int N) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    stats[4 * n + 0] = __int_as_float(0x7f800000);
    stats[4 * n + 1] = 0.f;
    stats[4 * n + 2] = 0.f;
    stats[4 * n + 3] = 0.f;
  }
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_upsample_workspace_bytes(int N, int h, int w) {
  return (N > 0 && h > 0 && w > 0) ? (size_t)N * h * w * sizeof(double) : 0;
}

extern "C" int halo_upsample_score_inputs(const float* logits_lr, const void* emb_lr, int emb_kind, float c,
                                          const uint8_t* gt, int pixunc_mode, int label_mode, int norm_mode,
                                          float* pixunc, uint8_t* label, float* radius, float* stats, int N, int O, int C,
                                          int lh, int lw, int eh, int ew, int H, int W, void* ws, size_t ws_bytes,
                                          halo_stream_t stream) {
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
  a.logits = logits_lr; a.emb = emb_lr; a.gamma = nullptr; a.gt = gt; a.pixunc = pixunc; a.label = label; a.radius = radius;
  a.stats = stats; a.emb_kind = emb_kind; a.pixunc_mode = pixunc_mode; a.label_mode = label_mode; a.norm_mode = norm_mode;
  a.N = N; a.O = O; a.C = C; a.lh = lh; a.lw = lw; a.eh = eh; a.ew = ew; a.H = H; a.W = W; a.c = c; a.inv_log19 = (float)(1.0 / log(19.0));
  if (emb_lr && emb_kind == HALO_FEAT_TANGENT_F32) {
    const size_t need = halo_upsample_workspace_bytes(N, eh, ew);
    if (!ws || ws_bytes < need) {
      set_error("halo_upsample_score_inputs: workspace %zu < %zu bytes", ws_bytes, need);
      return HALO_ERR_WORKSPACE;
    }
    const long long total = (long long)N * eh * ew;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    gamma_lr_kernel<<<(int)blocks, 256, 0, st>>>((const float*)emb_lr, (double*)ws, c, C, eh * ew, total);
    int rc = launch_status("gamma_lr_kernel");
    if (rc) return rc;
    a.gamma = (const double*)ws;
  }
  if (stats) {
    up_stats_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(stats, N);
    int rc = launch_status("up_stats_init_kernel");
    if (rc) return rc;
  }
  dim3 grid((W + 255) / 256, H, N);
  if (emb_lr && emb_kind == HALO_FEAT_BALL_F64) upsample_inputs_kernel<double><<<grid, 256, 0, st>>>(a);
  else upsample_inputs_kernel<float><<<grid, 256, 0, st>>>(a);
  return launch_status("upsample_inputs_kernel");
}
