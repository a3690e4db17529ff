// Eager pieces of the head for callers that need materialised tensors:
//   halo_expmap0_project  -- HyperMapper.expmap                (core/utils/hyperbolic.py:28-39)
//   halo_ball_norm        -- poincare_distance_origin / |x|    (core/utils/hyperbolic.py:74-83, floating_region.py:195)
//   halo_logits_stats     -- softmax entropy / 1-p[gt] / argmax from explicit logits
//                            (core/active/floating_region.py:151-152, 70-83, 123-127, 166, 172-173)
#include "common.cuh"

namespace halo {

constexpr int MISC_THREADS = 256;

// one thread per pixel; channel rows are pixel-contiguous so every load/store is coalesced.
// The per-pixel scalar is formed in double when the caller wants an fp64 embedding (what the reference
// returns), so that |x| near the ball boundary keeps its digits.
template <typename TOUT>
__global__ void expmap_kernel(const float* __restrict__ u, TOUT* __restrict__ x, HeadConsts hc, int C, int HW, long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long n = g / HW;
    const int p = (int)(g - n * HW);
    const float* src = u + (size_t)n * C * HW + p;
    TOUT* dst = x + (size_t)n * C * HW + p;
    if (sizeof(TOUT) == 8) {
      double n2 = 0.0;
      for (int ch = 0; ch < C; ++ch) {
        const double v = (double)src[(size_t)ch * HW];
        n2 = fma(v, v, n2);
      }
      const double s = sqrt((double)hc.c);
      const double nn = fmax(sqrt(n2), 1e-15);
      const double t = fmin(tanh(fmin(s * nn, 15.0)), 1.0 - 1e-5);
      const double gamma = t / (s * nn);
      for (int ch = 0; ch < C; ++ch) dst[(size_t)ch * HW] = (TOUT)(gamma * (double)src[(size_t)ch * HW]);
    } else {
      float n2 = 0.f;
      for (int ch = 0; ch < C; ++ch) {
        const float v = src[(size_t)ch * HW];
        n2 = fmaf(v, v, n2);
      }
      const float nn = sqrtf(n2);
      const float sn = hc.s * nn;
      const float t = (sn > hc.z_clip) ? hc.t_clip : tanhf(sn);
      const float gamma = t / (hc.s * fmaxf(nn, 1e-15f));
      for (int ch = 0; ch < C; ++ch) dst[(size_t)ch * HW] = (TOUT)(gamma * src[(size_t)ch * HW]);
    }
  }
}

template <typename TIN>
__global__ void ball_norm_kernel(const TIN* __restrict__ x, float* __restrict__ out, float* __restrict__ stats,
                                 HeadConsts hc, int norm_mode, int C, int HW, int blocks_per_img) {
  const int n = blockIdx.x / blocks_per_img;
  const int p = (blockIdx.x - n * blocks_per_img) * blockDim.x + threadIdx.x;
  float r = 0.f;
  float rmin = __int_as_float(0x7f800000), rmax = 0.f;
  if (p < HW) {
    const TIN* src = x + (size_t)n * C * HW + p;
    double n2 = 0.0;
    for (int ch = 0; ch < C; ++ch) {
      const double v = (double)src[(size_t)ch * HW];
      n2 = fma(v, v, n2);
    }
    if (norm_mode == HALO_NORM_EUCLID) {
      r = (float)sqrt(n2);
    } else {
      const double t = fmin(sqrt((double)hc.c * n2), 1.0 - 1e-7);
      r = hc.two_over_s * (float)(0.5 * (log1p(t) - log1p(-t)));
    }
    out[(size_t)n * HW + p] = r;
    rmin = rmax = r;
  }
  if (stats != nullptr) {
    rmin = warp_min(rmin);
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(reinterpret_cast<int*>(stats + 4 * n + 0), __float_as_int(rmin));
      atomicMax(reinterpret_cast<int*>(stats + 4 * n + 1), __float_as_int(rmax));
    }
  }
}

// fp64 radius plane + per-image fp64 extrema for the "hyper" purity (floating_region.py:94-110): the reference
// quantises an fp64 radius into K bins with round-half-even, so an fp32 radius flips bins at ~1e-4 of the pixels; the bins
// only match when the radius is formed the way the reference forms it, from an fp64 |u|^2 (tangent features: closed form
// of expmap0 + project + dist0) or an fp64 |x|^2 (points already on the ball).  Radii are >= 0, so their bit patterns
// order like unsigned integers and 64-bit integer atomics give the extrema.
template <typename TIN>
__global__ void radius64_kernel(const TIN* __restrict__ x, double* __restrict__ out, unsigned long long* __restrict__ stats,
                                int tangent, double c, int C, int HW, int blocks_per_img) {
  const int n = blockIdx.x / blocks_per_img;
  const int p = (blockIdx.x - n * blocks_per_img) * blockDim.x + threadIdx.x;
  unsigned long long lo = 0x7ff0000000000000ull, hi = 0ull;
  if (p < HW) {
    const TIN* src = x + (size_t)n * C * HW + p;
    double n2 = 0.0;
    for (int ch = 0; ch < C; ++ch) {
      const double v = (double)src[(size_t)ch * HW];
      n2 = fma(v, v, n2);
    }
    const double s = sqrt(c);
    double t;
    if (tangent) {
      const double nn = fmax(sqrt(n2), 1e-15);
      t = fmin(tanh(fmin(s * nn, 15.0)), 1.0 - 1e-5);   // s*|x| after expmap0 + project (fp64 eps 1e-5)
    } else {
      t = fmin(s * sqrt(n2), 1.0 - 1e-7);                // geoopt artan_k clamp (hyperbolic.py:83)
    }
    const double r = (log1p(t) - log1p(-t)) / s;         // 2 artanh(t) / s
    out[(size_t)n * HW + p] = r;
    lo = hi = (unsigned long long)__double_as_longlong(r);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stats + 2 * n + 0, lo);
    atomicMax(stats + 2 * n + 1, hi);
  }
}

__global__ void stats64_init_kernel(unsigned long long* stats, int N) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    stats[2 * n + 0] = 0x7ff0000000000000ull;  // +inf
    stats[2 * n + 1] = 0ull;
  }
}

__global__ void stats_init_kernel(float* stats, int N) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    stats[4 * n + 0] = __int_as_float(0x7f800000);
    stats[4 * n + 1] = 0.f;
    stats[4 * n + 2] = 0.f;
    stats[4 * n + 3] = 0.f;
  }
}

__global__ void logits_stats_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ gt, int pixunc_mode,
                                    int label_mode, float* __restrict__ pixunc, uint8_t* __restrict__ label, int O, int HW,
                                    long long total, float inv_log19) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long n = g / HW;
    const int p = (int)(g - n * HW);
    const float* src = logits + (size_t)n * O * HW + p;
    float mx = src[0];
    int arg = 0;
    for (int k = 1; k < O; ++k) {
      const float v = src[(size_t)k * HW];
      if (v > mx) { mx = v; arg = k; }
    }
    float Z = 0.f;
    for (int k = 0; k < O; ++k) Z += __expf(src[(size_t)k * HW] - mx);
    const float iz = 1.f / Z;
    int g8 = (gt != nullptr) ? gt[g] : 255;
    const int gtf = (g8 == 255) ? arg : g8;
    if (pixunc != nullptr) {
      float v;
      if (pixunc_mode == HALO_PIXUNC_ENTROPY) {
        float ent = 0.f;
        for (int k = 0; k < O; ++k) {
          const float pk = __expf(src[(size_t)k * HW] - mx) * iz;
          ent -= pk * __logf(pk + 1e-6f);
        }
        v = ent * inv_log19;
      } else {
        v = (gtf < O) ? 1.f - __expf(src[(size_t)gtf * HW] - mx) * iz : 1.f;
      }
      pixunc[g] = v;
    }
    if (label != nullptr) label[g] = (uint8_t)((label_mode == HALO_LABEL_GT_FILLED) ? gtf : arg);
  }
}

}  // namespace halo

using namespace halo;

extern "C" int halo_expmap0_project(const float* u, void* x_out, int out_f64, float c, int N, int C, int H, int W,
                                    halo_stream_t stream) {
  HALO_CHECK_ARG(u && x_out, "halo_expmap0_project: NULL pointer");
  HALO_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && c > 0.f, "halo_expmap0_project: bad dims / curvature");
  const long long total = (long long)N * H * W;
  const int grid = (int)((total + MISC_THREADS - 1) / MISC_THREADS < (long long)sm_count() * 16
                             ? (total + MISC_THREADS - 1) / MISC_THREADS
                             : (long long)sm_count() * 16);
  const HeadConsts hc = make_head_consts(c);
  if (out_f64)
    expmap_kernel<double><<<grid, MISC_THREADS, 0, (cudaStream_t)stream>>>(u, (double*)x_out, hc, C, H * W, total);
  else
    expmap_kernel<float><<<grid, MISC_THREADS, 0, (cudaStream_t)stream>>>(u, (float*)x_out, hc, C, H * W, total);
  return launch_status("expmap_kernel");
}

extern "C" int halo_ball_norm(const void* x, int x_f64, float c, int norm_mode, float* out, float* stats, int N, int C,
                              int H, int W, halo_stream_t stream) {
  HALO_CHECK_ARG(x && out, "halo_ball_norm: NULL pointer");
  HALO_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && c > 0.f, "halo_ball_norm: bad dims / curvature");
  HALO_CHECK_ARG(norm_mode == HALO_NORM_RADIUS || norm_mode == HALO_NORM_EUCLID, "halo_ball_norm: bad norm_mode");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int bpi = (HW + MISC_THREADS - 1) / MISC_THREADS;
  if (stats) {
    stats_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(stats, N);
    int rc = launch_status("stats_init_kernel");
    if (rc) return rc;
  }
  const HeadConsts hc = make_head_consts(c);
  if (x_f64)
    ball_norm_kernel<double><<<bpi * N, MISC_THREADS, 0, st>>>((const double*)x, out, stats, hc, norm_mode, C, HW, bpi);
  else
    ball_norm_kernel<float><<<bpi * N, MISC_THREADS, 0, st>>>((const float*)x, out, stats, hc, norm_mode, C, HW, bpi);
  return launch_status("ball_norm_kernel");
}

extern "C" int halo_radius_f64(const void* feat, int feat_kind, float c, double* radius64, double* stats64, int N, int C,
                               int H, int W, halo_stream_t stream) {
  HALO_CHECK_ARG(feat && radius64 && stats64, "halo_radius_f64: NULL pointer");
  HALO_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && c > 0.f, "halo_radius_f64: bad dims / curvature");
  HALO_CHECK_ARG(feat_kind >= 0 && feat_kind <= 2, "halo_radius_f64: bad feat_kind");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int bpi = (HW + MISC_THREADS - 1) / MISC_THREADS;
  unsigned long long* s64 = reinterpret_cast<unsigned long long*>(stats64);
  stats64_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(s64, N);
  int rc = launch_status("stats64_init_kernel");
  if (rc) return rc;
  if (feat_kind == HALO_FEAT_BALL_F64)
    radius64_kernel<double><<<bpi * N, MISC_THREADS, 0, st>>>((const double*)feat, radius64, s64, 0, (double)c, C, HW, bpi);
  else
    radius64_kernel<float><<<bpi * N, MISC_THREADS, 0, st>>>((const float*)feat, radius64, s64,
                                                              feat_kind == HALO_FEAT_TANGENT_F32, (double)c, C, HW, bpi);
  return launch_status("radius64_kernel");
}

extern "C" int halo_logits_stats(const float* logits, const uint8_t* gt, int pixunc_mode, int label_mode, float* pixunc,
                                 uint8_t* label, int N, int O, int H, int W, halo_stream_t stream) {
  HALO_CHECK_ARG(logits && (pixunc || label), "halo_logits_stats: NULL pointer");
  HALO_CHECK_ARG(N > 0 && O > 0 && H > 0 && W > 0, "halo_logits_stats: bad dims");
  HALO_CHECK_ARG(!((pixunc_mode == HALO_PIXUNC_ONE_MINUS_PGT && pixunc) || (label_mode == HALO_LABEL_GT_FILLED && label)) || gt,
                 "halo_logits_stats: gt required by the requested mode");
  const long long total = (long long)N * H * W;
  long long blocks = (total + MISC_THREADS - 1) / MISC_THREADS;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  logits_stats_kernel<<<(int)blocks, MISC_THREADS, 0, (cudaStream_t)stream>>>(
      logits, gt, pixunc_mode, label_mode, pixunc, label, O, H * W, total, (float)(1.0 / log(19.0)));
  return launch_status("logits_stats_kernel");
}
