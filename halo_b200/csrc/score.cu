// K2 -- floating-region acquisition score (core/active/floating_region.py:129-217 after the softmax).
//
// Pass A (halo-tiled): region uncertainty = zero-padded k x k box SUM of the per-pixel map (:42-51,:90),
//   region impurity = entropy of the pk x pk label histogram (:112-121) or the radius plane (:187-193),
//   uncertainty /= count (:204); per-image min/max of both maps via ordered-integer atomics.
// Pass B (elementwise): optional min-max normalisation of both maps (:22-23,:206-208), product (:210),
//   and the already-labelled mask score[active] = -inf (core/active/build.py:146).
#include "common.cuh"

namespace halo {

constexpr int SC_TW = 64, SC_TH = 16, SC_THREADS = 256, SC_ROWS_PER_THREAD = SC_TH / (SC_THREADS / SC_TW);

struct ScoreArgs {
  const float* pixunc;
  const float* radius;
  const float* radius_stats;
  const double* radius64;        // fp64 radius plane + {min,max} per image: the "hyper" bins follow the reference's
  const double* radius_stats64;  // fp64 arithmetic when given (floating_region.py:94-110)
  const uint8_t* label;
  const uint8_t* active;
  float* score;
  float* impurity;
  float* uncertainty;
  unsigned* mm;  // [N][4] ordered-uint {unc_min, unc_max, imp_min, imp_max}
  int unc_mode, pur_mode, normalize, k, pk, n_bins;
  int N, H, W;
  float inv_log_bins;
};

__global__ void score_init_kernel(unsigned* mm, int N) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    mm[4 * i + 0] = 0xffffffffu;
    mm[4 * i + 1] = 0u;
    mm[4 * i + 2] = 0xffffffffu;
    mm[4 * i + 3] = 0u;
  }
}

// quantize_uncert_map (:94-110): the second min-max is the identity (min(1-x)=0, max(1-x)=1 exactly)
__device__ __forceinline__ int radius_bin(float r, float rmin, float rmax, int K) {
  const float x = (r - rmin) / (rmax - rmin);
  float b = (1.f - x) * (float)K - 0.5f;
  b = fminf(fmaxf(b, -0.5f + 1e-5f), (float)K - 0.5f - 1e-5f);
  return (int)rintf(b);  // round half to even, like torch.round
}

// the same in the reference's own precision: fp64 radius, fp64 extrema (python floats), torch.round on doubles
__device__ __forceinline__ int radius_bin64(double r, double rmin, double rmax, int K) {
  const double x = (r - rmin) / (rmax - rmin);
  double b = (1.0 - x) * (double)K - 0.5;
  b = fmin(fmax(b, -0.5 + 1e-5), (double)K - 0.5 - 1e-5);
  return (int)rint(b);
}

__global__ void __launch_bounds__(SC_THREADS) score_pass_a_kernel(const ScoreArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ru = (a.unc_mode == HALO_UNC_BOXSUM) ? a.k / 2 : 0;
  const bool hist = (a.pur_mode == HALO_PUR_LABEL_HIST || a.pur_mode == HALO_PUR_RADIUS_BINS);
  const int rp = hist ? a.pk / 2 : 0;
  const int r = ru > rp ? ru : rp;
  const int SW = SC_TW + 2 * r, SH = SC_TH + 2 * r;
  float* sU = reinterpret_cast<float*>(smem_raw);                  // [SH][SW] per-pixel uncertainty, 0 outside
  uint8_t* sL = reinterpret_cast<uint8_t*>(sU + (size_t)SH * SW);  // [SH][SW] labels / bins, 255 outside
  __shared__ unsigned s_present[8];                                // class-presence bitmap of the tile (<=256 bins)
  __shared__ unsigned s_mm[4];

  const int n = blockIdx.z;
  const int x0 = blockIdx.x * SC_TW, y0 = blockIdx.y * SC_TH;
  const size_t plane = (size_t)n * a.H * a.W;
  if (threadIdx.x < 8) s_present[threadIdx.x] = 0u;
  if (threadIdx.x < 4) s_mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
  float rmin = 0.f, rmax = 1.f;
  double rmin64 = 0.0, rmax64 = 1.0;
  const bool bins64 = (a.pur_mode == HALO_PUR_RADIUS_BINS && a.radius64 != nullptr);
  if (bins64) {
    rmin64 = a.radius_stats64[2 * n + 0];
    rmax64 = a.radius_stats64[2 * n + 1];
  } else if (a.pur_mode == HALO_PUR_RADIUS_BINS) {
    rmin = a.radius_stats[4 * n + 0];
    rmax = a.radius_stats[4 * n + 1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SH * SW; i += SC_THREADS) {
    const int sy = i / SW, sx = i - sy * SW;
    const int y = y0 - r + sy, x = x0 - r + sx;
    const bool in = (y >= 0 && y < a.H && x >= 0 && x < a.W);
    float u = 0.f;
    int lab = 255;
    if (in) {
      const size_t g = plane + (size_t)y * a.W + x;
      if (a.unc_mode != HALO_UNC_ZERO) u = a.pixunc[g];
      if (a.pur_mode == HALO_PUR_LABEL_HIST) lab = a.label[g];
      else if (bins64) lab = radius_bin64(a.radius64[g], rmin64, rmax64, a.n_bins);
      else if (a.pur_mode == HALO_PUR_RADIUS_BINS) lab = radius_bin(a.radius[g], rmin, rmax, a.n_bins);
      if (hist && lab < 255) atomicOr(&s_present[lab >> 5], 1u << (lab & 31));
    }
    sU[i] = u;
    sL[i] = (uint8_t)lab;
  }
  __syncthreads();

  const int tx = threadIdx.x % SC_TW, ty = threadIdx.x / SC_TW;
  unsigned umin = 0xffffffffu, umax = 0u, imin = 0xffffffffu, imax = 0u;
#pragma unroll 1
  for (int j = 0; j < SC_ROWS_PER_THREAD; ++j) {
    const int ly = ty + j * (SC_THREADS / SC_TW);
    const int y = y0 + ly, x = x0 + tx;
    if (y >= a.H || x >= a.W) continue;
    const int cy = r + ly, cx = r + tx;
    // --- region uncertainty ---
    float unc = 0.f;
    if (a.unc_mode == HALO_UNC_BOXSUM) {
      for (int dy = -ru; dy <= ru; ++dy) {
        const float* row = sU + (size_t)(cy + dy) * SW + cx;
        float rs = 0.f;
        for (int dx = -ru; dx <= ru; ++dx) rs += row[dx];
        unc += rs;
      }
    } else if (a.unc_mode == HALO_UNC_PIXEL) {
      unc = sU[(size_t)cy * SW + cx];
    }
    // --- region impurity ---
    float imp = 0.f, count = 1.f;
    if (hist) {
      const int ylo = max(y - rp, 0), yhi = min(y + rp, a.H - 1), xlo = max(x - rp, 0), xhi = min(x + rp, a.W - 1);
      const int cnt_all = (yhi - ylo + 1) * (xhi - xlo + 1);
      count = (float)cnt_all;
      const float inv_cnt = 1.f / count;
      float ent = 0.f;
      for (int wd = 0; wd < 8; ++wd) {
        unsigned bits = s_present[wd];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int cls = wd * 32 + b;
          int cnt = 0;
          for (int dy = -rp; dy <= rp; ++dy) {
            const uint8_t* row = sL + (size_t)(cy + dy) * SW + cx;
            for (int dx = -rp; dx <= rp; ++dx) cnt += (row[dx] == cls);
          }
          if (cnt) {
            const float d = (float)cnt * inv_cnt;
            ent -= d * logf(d + 1e-6f);
          }
        }
      }
      imp = ent * a.inv_log_bins;
    } else if (a.pur_mode == HALO_PUR_NORM) {
      imp = a.radius[plane + (size_t)y * a.W + x];
    }
    unc = unc / count;  // :204
    const size_t g = plane + (size_t)y * a.W + x;
    a.uncertainty[g] = unc;
    if (a.impurity != nullptr) a.impurity[g] = imp;
    const unsigned uo = f2ord(unc), io = f2ord(imp);
    umin = min(umin, uo); umax = max(umax, uo); imin = min(imin, io); imax = max(imax, io);
  }
  if (a.normalize) {
    for (int o = 16; o > 0; o >>= 1) {
      umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
      umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
      imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
      imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_mm[0], umin); atomicMax(&s_mm[1], umax); atomicMin(&s_mm[2], imin); atomicMax(&s_mm[3], imax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicMin(&a.mm[4 * n + 0], s_mm[0]); atomicMax(&a.mm[4 * n + 1], s_mm[1]);
      atomicMin(&a.mm[4 * n + 2], s_mm[2]); atomicMax(&a.mm[4 * n + 3], s_mm[3]);
    }
  }
}

// Pass A, fast path for the radius-type purities (impurity := radius plane, count := 1) and small windows: no shared
// memory, one thread per column walking a strip of rows with the last 2r+1 horizontal sums in registers.
// ~30 instructions per pixel (the generic halo-tile kernel below is instruction-bound at ~300, profiles/r1_k2_k3.md).
constexpr int SCF_ROWS = 32;     // rows per thread strip
constexpr int SCF_THREADS = 256; // columns per block
template <int R>
__global__ void __launch_bounds__(SCF_THREADS) score_pass_a_fast_kernel(const ScoreArgs a) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * SCF_THREADS + threadIdx.x;
  const int y0 = blockIdx.y * SCF_ROWS;
  const size_t plane = (size_t)n * a.H * a.W;
  const float* pu = a.pixunc + plane;
  const bool box = (a.unc_mode == HALO_UNC_BOXSUM);
  unsigned umin = 0xffffffffu, umax = 0u, imin = 0xffffffffu, imax = 0u;
  __shared__ unsigned s_mm[4];
  if (threadIdx.x < 4) s_mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
  __syncthreads();
  if (x < a.W) {
    float ring[2 * R + 1];
#pragma unroll
    for (int i = 0; i < 2 * R + 1; ++i) ring[i] = 0.f;
    auto hsum = [&](int y) -> float {
      if (y < 0 || y >= a.H || a.unc_mode == HALO_UNC_ZERO) return 0.f;
      const float* row = pu + (size_t)y * a.W;
      if (!box) return row[x];
      float s = 0.f;
#pragma unroll
      for (int dx = -R; dx <= R; ++dx) {
        const int xx = x + dx;
        if (xx >= 0 && xx < a.W) s += __ldg(row + xx);
      }
      return s;
    };
    // prime the window with rows y0-R .. y0+R-1
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) ring[i + 1] = box ? hsum(y0 - R + i) : 0.f;
    const int y_end = min(y0 + SCF_ROWS, a.H);
    for (int y = y0; y < y_end; ++y) {
#pragma unroll
      for (int i = 0; i < 2 * R; ++i) ring[i] = ring[i + 1];
      ring[2 * R] = box ? hsum(y + R) : 0.f;
      float unc;
      if (box) {
        unc = 0.f;
#pragma unroll
        for (int i = 0; i < 2 * R + 1; ++i) unc += ring[i];
      } else {
        unc = hsum(y);
      }
      const size_t g = plane + (size_t)y * a.W + x;
      const float imp = (a.pur_mode == HALO_PUR_NORM) ? a.radius[g] : 0.f;
      a.uncertainty[g] = unc;
      if (a.impurity != nullptr) a.impurity[g] = imp;
      const unsigned uo = f2ord(unc), io = f2ord(imp);
      umin = min(umin, uo); umax = max(umax, uo); imin = min(imin, io); imax = max(imax, io);
    }
  }
  if (a.normalize) {
    for (int o = 16; o > 0; o >>= 1) {
      umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
      umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
      imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
      imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_mm[0], umin); atomicMax(&s_mm[1], umax); atomicMin(&s_mm[2], imin); atomicMax(&s_mm[3], imax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicMin(&a.mm[4 * n + 0], s_mm[0]); atomicMax(&a.mm[4 * n + 1], s_mm[1]);
      atomicMin(&a.mm[4 * n + 2], s_mm[2]); atomicMax(&a.mm[4 * n + 3], s_mm[3]);
    }
  }
}

// Pass B, 4 pixels per thread (planes are 16-byte aligned and H*W % 4 == 0 in the vector variant).
// normalize: 0 = off; 1 = on, the two maps are normalised in place (what the reference returns);
//            2 = on, maps left as pass A wrote them (the acquisition path only needs the score: saves 4-8 B/px)
template <bool VEC4>
__global__ void score_pass_b_kernel(const ScoreArgs a, long long total) {
  const int HW = a.H * a.W;
  constexpr int V = VEC4 ? 4 : 1;
  const float ninf = __int_as_float(0xff800000);
  for (long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V; g < total; g += (long long)gridDim.x * blockDim.x * V) {
    const int n = (int)(g / HW);  // the V pixels of a thread never straddle images (HW % 4 == 0 when VEC4)
    float unc[V], imp[V], sc[V];
    if (VEC4) {
      const float4 u4 = *reinterpret_cast<const float4*>(a.uncertainty + g);
      unc[0] = u4.x; unc[1] = u4.y; unc[2] = u4.z; unc[3] = u4.w;
      if (a.impurity != nullptr || a.pur_mode == HALO_PUR_NORM) {
        const float4 i4 = *reinterpret_cast<const float4*>((a.impurity != nullptr ? a.impurity : a.radius) + g);
        imp[0] = i4.x; imp[1] = i4.y; imp[2] = i4.z; imp[3] = i4.w;
      } else {
        imp[0] = imp[1] = imp[2] = imp[3] = 0.f;
      }
    } else {
      unc[0] = a.uncertainty[g];
      imp[0] = (a.impurity != nullptr) ? a.impurity[g] : (a.pur_mode == HALO_PUR_NORM ? a.radius[g] : 0.f);
    }
    if (a.normalize) {
      // normalize_map (:22-23): extrema are python floats, the subtraction/division run in the map's dtype
      const float ulo = ord2f(a.mm[4 * n + 0]), uhi = ord2f(a.mm[4 * n + 1]);
      const float ilo = ord2f(a.mm[4 * n + 2]), ihi = ord2f(a.mm[4 * n + 3]);
      const float ud = (float)((double)uhi - (double)ulo), id = (float)((double)ihi - (double)ilo);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        unc[e] = (unc[e] - ulo) / ud;
        imp[e] = (imp[e] - ilo) / id;
      }
      if (a.normalize == 1) {
        if (VEC4) {
          *reinterpret_cast<float4*>(a.uncertainty + g) = make_float4(unc[0], unc[1], unc[2], unc[3]);
          if (a.impurity != nullptr) *reinterpret_cast<float4*>(a.impurity + g) = make_float4(imp[0], imp[1], imp[2], imp[3]);
        } else {
          a.uncertainty[g] = unc[0];
          if (a.impurity != nullptr) a.impurity[g] = imp[0];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) sc[e] = imp[e] * unc[e];
    if (a.active != nullptr) {
      if (VEC4) {
        const uchar4 m4 = *reinterpret_cast<const uchar4*>(a.active + g);
        if (m4.x) sc[0] = ninf;
        if (m4.y) sc[1] = ninf;
        if (m4.z) sc[2] = ninf;
        if (m4.w) sc[3] = ninf;
      } else if (a.active[g]) {
        sc[0] = ninf;
      }
    }
    if (VEC4) *reinterpret_cast<float4*>(a.score + g) = make_float4(sc[0], sc[1], sc[2], sc[3]);
    else a.score[g] = sc[0];
  }
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_score_workspace_bytes(int N) { return N > 0 ? (size_t)N * 4 * sizeof(unsigned) : 0; }

extern "C" int halo_score(const float* pixunc, const float* radius, const float* radius_stats, const double* radius64,
                          const double* radius_stats64, const uint8_t* label, const uint8_t* active, int unc_mode, int pur_mode, int normalize, int k, int pk, int n_bins,
                          float* score, float* impurity, float* uncertainty, int N, int H, int W, void* ws,
                          size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(score && uncertainty, "halo_score: score and uncertainty planes are required");
  HALO_CHECK_ARG(N > 0 && H > 0 && W > 0, "halo_score: bad dims");
  HALO_CHECK_ARG(unc_mode >= 0 && unc_mode <= 2 && pur_mode >= 0 && pur_mode <= 3, "halo_score: bad mode");
  HALO_CHECK_ARG(k > 0 && (k & 1) && pk > 0 && (pk & 1), "halo_score: window sizes must be odd (got k=%d pk=%d)", k, pk);
  HALO_CHECK_ARG(unc_mode == HALO_UNC_ZERO || pixunc, "halo_score: pixunc plane required");
  HALO_CHECK_ARG(pur_mode != HALO_PUR_NORM || radius, "halo_score: radius plane required");
  HALO_CHECK_ARG(pur_mode != HALO_PUR_RADIUS_BINS || (radius && radius_stats) || (radius64 && radius_stats64),
                 "halo_score: radius bins need a radius plane with its stats (fp32 or fp64)");
  HALO_CHECK_ARG((radius64 == nullptr) == (radius_stats64 == nullptr), "halo_score: radius64 and radius_stats64 go together");
  HALO_CHECK_ARG(pur_mode != HALO_PUR_LABEL_HIST || label, "halo_score: label plane required");
  HALO_CHECK_ARG(!(pur_mode == HALO_PUR_LABEL_HIST || pur_mode == HALO_PUR_RADIUS_BINS) || (n_bins >= 2 && n_bins <= 255),
                 "halo_score: n_bins must be in [2,255] (got %d)", n_bins);
  HALO_CHECK_ARG(pur_mode == HALO_PUR_NORM || impurity || pur_mode == HALO_PUR_ZERO, "halo_score: impurity plane required for histogram purity");
  if (k > 65 || pk > 65) {
    set_error("halo_score: windows larger than 65 not compiled");
    return HALO_ERR_UNSUPPORTED;
  }
  const size_t need = halo_score_workspace_bytes(N);
  if (normalize && (!ws || ws_bytes < need)) {
    set_error("halo_score: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ScoreArgs a;
  a.pixunc = pixunc; a.radius = radius; a.radius_stats = radius_stats; a.label = label; a.active = active;
  a.radius64 = radius64; a.radius_stats64 = radius_stats64;
  a.score = score; a.impurity = impurity; a.uncertainty = uncertainty; a.mm = (unsigned*)ws;
  a.unc_mode = unc_mode; a.pur_mode = pur_mode; a.normalize = normalize; a.k = k; a.pk = pk; a.n_bins = n_bins;
  a.N = N; a.H = H; a.W = W;
  a.inv_log_bins = (n_bins >= 2) ? (float)(1.0 / log((double)n_bins)) : 0.f;
  HALO_CHECK_ARG(normalize >= 0 && normalize <= 2, "halo_score: normalize must be 0, 1 or 2");
  if (normalize) {
    score_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(a.mm, N);
    int rc = launch_status("score_init_kernel");
    if (rc) return rc;
  }
  const bool hist = (pur_mode == HALO_PUR_LABEL_HIST || pur_mode == HALO_PUR_RADIUS_BINS);
  const int ru = (unc_mode == HALO_UNC_BOXSUM) ? k / 2 : 0, rp = hist ? pk / 2 : 0, r = ru > rp ? ru : rp;
  int rc;
  if (!hist && ru <= 2) {
    dim3 grid((W + SCF_THREADS - 1) / SCF_THREADS, (H + SCF_ROWS - 1) / SCF_ROWS, N);
    HALO_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "halo_score: grid too large");
    if (ru == 0) score_pass_a_fast_kernel<0><<<grid, SCF_THREADS, 0, st>>>(a);
    else if (ru == 1) score_pass_a_fast_kernel<1><<<grid, SCF_THREADS, 0, st>>>(a);
    else score_pass_a_fast_kernel<2><<<grid, SCF_THREADS, 0, st>>>(a);
    rc = launch_status("score_pass_a_fast_kernel");
  } else {
    const size_t smem = (size_t)(SC_TW + 2 * r) * (SC_TH + 2 * r) * 5 + 16;
    HALO_CUDA(cudaFuncSetAttribute(score_pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + SC_TW - 1) / SC_TW, (H + SC_TH - 1) / SC_TH, N);
    HALO_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "halo_score: grid too large");
    score_pass_a_kernel<<<grid, SC_THREADS, smem, st>>>(a);
    rc = launch_status("score_pass_a_kernel");
  }
  if (rc) return rc;
  const long long total = (long long)N * H * W;
  auto aligned = [](const void* q, size_t al) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % al) == 0; };
  const bool vec4 = ((long long)H * W) % 4 == 0 && aligned(score, 16) && aligned(uncertainty, 16) && aligned(impurity, 16) &&
                    aligned(radius, 16) && aligned(active, 4);
  const int V = vec4 ? 4 : 1;
  long long blocks = (total / V + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  if (blocks < 1) blocks = 1;
  if (vec4) score_pass_b_kernel<true><<<(int)blocks, 256, 0, st>>>(a, total);
  else score_pass_b_kernel<false><<<(int)blocks, 256, 0, st>>>(a, total);
  return launch_status("score_pass_b_kernel");
}
