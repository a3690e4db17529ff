// K6 -- the losses that consume the head's logits in the training step, fused with the bilinear up-sampling in front of
// them and with its adjoint.
//
// Reference sequence (core/train_learners.py:343-356 target branch, :232-236 source branch; core/models/classifier.py:556-557;
// core/loss/negative_learning_loss.py:6-16):
//     out      = F.interpolate(logits_lr, size, mode="bilinear", align_corners=True)        (N,O,H,W) materialised
//     predict  = softmax(out, dim=1)
//     loss_sup = CrossEntropyLoss(ignore_index=255)(out, mask)            mean over labelled pixels, skipped when none
//     neg      = sum(-[p < thr] * log(1 - p + 1e-6)) / sum([p < thr]) * NEGATIVE_LOSS          the mask is detached
//     loss     = loss_sup + neg ;  loss.backward()  ->  d loss / d logits_lr through the up-sampling
// Here the (N,O,H,W) tensors never exist.  Pass A (thread = label-resolution pixel) interpolates the O logits in registers
// and reduces the four sums {CE, #labelled, negative term, #masked}; partials per block, fixed-order finish: bitwise
// reproducible.  Pass B (thread = low-resolution pixel) GATHERS the gradient: it revisits the label-resolution pixels whose
// interpolation stencil contains it, recomputes their softmax, and accumulates weight * dL/dz -- no atomics, unlike the
// scatter of torch's upsample_bilinear2d_backward, so the gradient is reproducible too.  Both passes read only the
// low-resolution logits (L2-resident) and the uint8 labels.
#include "common.cuh"
#include "head_common.cuh"

namespace halo {

constexpr int SL_THREADS = 256;
constexpr int SL_MAXO = 32;

struct SegLossArgs {
  const float* logits;    // [N,O,h,w]
  const uint8_t* labels;  // [N,H,W], 255 = ignore; NULL = no supervised term
  double* partial;        // [blocks][4]
  double* sums;           // [4] = {ce_sum, n_labelled, neg_sum, n_masked}
  float* losses;          // [4] = {total, loss_sup, weighted negative loss, n_labelled}
  float* dlogits;         // [N,O,h,w]
  int N, O, h, w, H, W;
  float sy, sx;           // (h-1)/(H-1), (w-1)/(W-1) in float, torch's area_pixel_compute_scale with align_corners=True
  float neg_weight, threshold;
};

// softmax of the interpolated logits of label-resolution pixel (Y, X) of image n; returns max-shifted exps in e[], 1/Z in iz
template <int OP>
__device__ __forceinline__ void interp_softmax(const SegLossArgs& a, const float* __restrict__ L, int Y, int X, int lab,
                                               float (&e)[OP], float& iz, float& zl, float& logZ) {
  const int hw = a.h * a.w;
  const float fy = a.sy * (float)Y, fx = a.sx * (float)X;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + ((y0 < a.h - 1) ? 1 : 0), x1 = x0 + ((x0 < a.w - 1) ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const int i00 = y0 * a.w + x0, i01 = y0 * a.w + x1, i10 = y1 * a.w + x0, i11 = y1 * a.w + x1;
  float mx = -3.0e38f;
  zl = 0.f;
#pragma unroll
  for (int k = 0; k < OP; ++k) {
    if (k < a.O) {
      const float* q = L + (size_t)k * hw;
      e[k] = hy * (hx * __ldg(q + i00) + lx * __ldg(q + i01)) + ly * (hx * __ldg(q + i10) + lx * __ldg(q + i11));
      mx = fmaxf(mx, e[k]);
    } else {
      e[k] = -3.0e38f;
    }
  }
  float Z = 0.f;
#pragma unroll
  for (int k = 0; k < OP; ++k) {
    if (k == lab) zl = e[k] - mx;     // shifted logit of the label class (log-softmax = zl - log Z, no exp/log round trip)
    e[k] = (k < a.O) ? __expf(e[k] - mx) : 0.f;
    Z += e[k];
  }
  iz = 1.f / Z;
  logZ = __logf(Z);
}

template <int OP>
__global__ void __launch_bounds__(SL_THREADS) seg_loss_sums_kernel(const SegLossArgs a) {
  const long long total = (long long)a.N * a.H * a.W;
  const int HWl = a.H * a.W;
  double ce = 0.0, nl = 0.0, ng = 0.0, nm = 0.0;
  for (long long g = (long long)blockIdx.x * SL_THREADS + threadIdx.x; g < total; g += (long long)gridDim.x * SL_THREADS) {
    const int n = (int)(g / HWl);
    const int r = (int)(g - (long long)n * HWl);
    const int Y = r / a.W, X = r - Y * a.W;
    const float* L = a.logits + (size_t)n * a.O * a.h * a.w;
    const int lab = (a.labels != nullptr) ? a.labels[g] : 255;
    float e[OP], iz, zl, logZ;
    interp_softmax<OP>(a, L, Y, X, lab, e, iz, zl, logZ);
    float negs = 0.f, cnt = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      if (k < a.O) {
        const float p = e[k] * iz;
        if (p < a.threshold) {
          negs -= __logf(1.f - p + 1e-6f);
          cnt += 1.f;
        }
      }
    }
    ng += (double)negs;
    nm += (double)cnt;
    if (lab != 255 && lab < a.O) {
      ce += (double)(logZ - zl);           // -log softmax(z)[label]
      nl += 1.0;
    }
  }
  // fixed-order block reduction
  __shared__ double red[4][SL_THREADS / 32];
  double v[4] = {ce, nl, ng, nm};
#pragma unroll
  for (int t = 0; t < 4; ++t)
    for (int o = 16; o > 0; o >>= 1) v[t] += __shfl_xor_sync(0xffffffffu, v[t], o);
  if ((threadIdx.x & 31) == 0)
    for (int t = 0; t < 4; ++t) red[t][threadIdx.x >> 5] = v[t];
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int wv = 0; wv < SL_THREADS / 32; ++wv) s += red[threadIdx.x][wv];
    a.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

__global__ void seg_loss_finish_kernel(const SegLossArgs a, int blocks) {
  __shared__ double tot[4];
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += a.partial[(size_t)b * 4 + threadIdx.x];
    a.sums[threadIdx.x] = s;
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double sup = tot[1] > 0.0 ? tot[0] / tot[1] : 0.0;                  // skipped when nothing is labelled (:345)
    const double neg = (a.neg_weight > 0.f) ? (tot[2] / tot[3]) * (double)a.neg_weight : 0.0;   // 0/0 = nan like the reference
    a.losses[0] = (float)(sup + neg);
    a.losses[1] = (float)sup;
    a.losses[2] = (float)neg;
    a.losses[3] = (float)tot[1];
  }
}

// weight with which label-resolution coordinate D (one axis, scale s, `in` low-resolution samples) reads low-resolution
// sample t:  [floor == t] * (1 - l) + [floor + 1 == t (clamped)] * l
__device__ __forceinline__ float tap_weight(int D, float s, int in, int t) {
  const float f = s * (float)D;
  const int i0 = (int)f;
  const int i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  const float l = f - (float)i0;
  return ((i0 == t) ? 1.f - l : 0.f) + ((i1 == t) ? l : 0.f);
}

template <int OP>
__global__ void __launch_bounds__(SL_THREADS) seg_loss_grad_kernel(const SegLossArgs a) {
  const long long total = (long long)a.N * a.h * a.w;
  const int hw = a.h * a.w;
  const double n_lab = a.sums[1], n_msk = a.sums[3];
  const float s_ce = (n_lab > 0.0) ? (float)(1.0 / n_lab) : 0.f;
  const float s_neg = (a.neg_weight > 0.f) ? (float)((double)a.neg_weight / n_msk) : 0.f;
  for (long long g = (long long)blockIdx.x * SL_THREADS + threadIdx.x; g < total; g += (long long)gridDim.x * SL_THREADS) {
    const int n = (int)(g / hw);
    const int r = (int)(g - (long long)n * hw);
    const int y = r / a.w, x = r - y * a.w;
    const float* L = a.logits + (size_t)n * a.O * hw;
    // label-resolution rows / columns whose stencil can contain (y, x): source coordinate in (y-1, y+1)
    int Y0 = (a.sy > 0.f) ? (int)floorf((float)(y - 1) / a.sy) : 0, Y1 = (a.sy > 0.f) ? (int)ceilf((float)(y + 1) / a.sy) : a.H - 1;
    int X0 = (a.sx > 0.f) ? (int)floorf((float)(x - 1) / a.sx) : 0, X1 = (a.sx > 0.f) ? (int)ceilf((float)(x + 1) / a.sx) : a.W - 1;
    Y0 = max(Y0 - 1, 0); Y1 = min(Y1 + 1, a.H - 1); X0 = max(X0 - 1, 0); X1 = min(X1 + 1, a.W - 1);
    float acc[OP];
#pragma unroll
    for (int k = 0; k < OP; ++k) acc[k] = 0.f;
    for (int Y = Y0; Y <= Y1; ++Y) {
      const float wy = tap_weight(Y, a.sy, a.h, y);
      if (wy == 0.f) continue;
      for (int X = X0; X <= X1; ++X) {
        const float wgt = wy * tap_weight(X, a.sx, a.w, x);
        if (wgt == 0.f) continue;
        const int lab = (a.labels != nullptr) ? a.labels[((size_t)n * a.H + Y) * a.W + X] : 255;
        float e[OP], iz, zl, logZ;
        interp_softmax<OP>(a, L, Y, X, lab, e, iz, zl, logZ);
        const float ce_on = (lab != 255 && lab < a.O) ? s_ce : 0.f;
        // negative term: sum_k m_k r_k p_k with r_k = 1 / (1 - p_k + 1e-6)
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < OP; ++k) {
          const float p = e[k] * iz;
          e[k] = p;
          if (k < a.O && p < a.threshold) q += p / (1.f - p + 1e-6f);
        }
#pragma unroll
        for (int k = 0; k < OP; ++k) {
          if (k < a.O) {
            const float p = e[k];
            float gk = ce_on * (p - ((k == lab) ? 1.f : 0.f));
            const float own = (p < a.threshold) ? p / (1.f - p + 1e-6f) : 0.f;
            gk += s_neg * (own - p * q);
            acc[k] = fmaf(wgt, gk, acc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < OP; ++k)
      if (k < a.O) a.dlogits[((size_t)n * a.O + k) * hw + r] = acc[k];
  }
}

// ---- tiled form of pass B ----------------------------------------------------------------------------------------------
// The gather above recomputes the softmax of a label-resolution pixel once per low-resolution pixel whose stencil holds it
// (4 x) and fetches the four corner logits of 19 classes from global memory every time: 2.1 ms at 8 x 19 x 160x320 ->
// 640x1280.  Here a CTA owns an 8 x 8 low-resolution tile: (1) the tile's logits (+ 1-pixel halo) go to shared memory,
// (2) dL/dz of every label-resolution pixel of the tile's footprint is computed ONCE into shared memory, (3) every
// (low-resolution pixel, class) sums its taps from there in a fixed order.  Same arithmetic per pixel, still no atomics.
constexpr int SLT_T = 8;          // low-resolution tile edge
#ifndef HALO_SLT_THREADS
#define HALO_SLT_THREADS 768
#endif
constexpr int SLT_THREADS = HALO_SLT_THREADS;   // one CTA per SM (the footprint fills shared memory): many warps hide the latency
constexpr int SLT_MAXTAP = 20;    // label-resolution rows (columns) that can read one low-resolution row (column)
struct SltGeom { int fh_max, fw_max; size_t smem; };

template <int OP>
__global__ void __launch_bounds__(SLT_THREADS) seg_loss_grad_tiled_kernel(const SegLossArgs a, int fh_max, int fw_max) {
  extern __shared__ float sm[];
  const int O = a.O, hw = a.h * a.w;
  const int n = blockIdx.z, y0 = blockIdx.y * SLT_T, x0 = blockIdx.x * SLT_T;
  constexpr int LT = SLT_T + 2;                       // logits tile with halo
  float* sL = sm;                                     // [O][LT][LT]
  float* sG = sL + (size_t)O * LT * LT;               // [O][fh_max * fw_max]
  float* sWy = sG + (size_t)O * fh_max * fw_max;      // [SLT_T][SLT_MAXTAP] row weights of low-resolution row y0 + i
  float* sWx = sWy + SLT_T * SLT_MAXTAP;
  __shared__ int geo[4];                              // Ylo, Yhi, Xlo, Xhi of the footprint
  __shared__ int tapY[SLT_T][2], tapX[SLT_T][2];      // first label-resolution row / column and tap count per low-res row / column
  const double n_lab = a.sums[1], n_msk = a.sums[3];
  const float s_ce = (n_lab > 0.0) ? (float)(1.0 / n_lab) : 0.f;
  const float s_neg = (a.neg_weight > 0.f) ? (float)((double)a.neg_weight / n_msk) : 0.f;
  const float* L = a.logits + (size_t)n * O * hw;
  // (0) footprint: label-resolution rows whose floor(Y * sy) lies in [y0 - 1, y0 + T - 1] (their stencil can touch the tile)
  if (threadIdx.x < 2) {
    const bool isy = (threadIdx.x == 0);
    const float sc = isy ? a.sy : a.sx;
    const int t0 = isy ? y0 : x0, D = isy ? a.H : a.W;
    int lo = (int)floorf((float)(t0 - 1) / sc) - 1, hi = (int)ceilf((float)(t0 + SLT_T) / sc) + 1;
    lo = max(lo, 0); hi = min(hi, D - 1);
    while (lo < D - 1 && (int)(sc * (float)lo) < t0 - 1) ++lo;
    while (hi > 0 && (int)(sc * (float)hi) > t0 + SLT_T - 1) --hi;
    geo[isy ? 0 : 2] = lo; geo[isy ? 1 : 3] = hi;
  }
  for (int i = threadIdx.x; i < O * LT * LT; i += SLT_THREADS) {
    const int k = i / (LT * LT), r = i - k * LT * LT;
    const int yy = min(max(y0 - 1 + r / LT, 0), a.h - 1), xx = min(max(x0 - 1 + r % LT, 0), a.w - 1);
    sL[i] = __ldg(L + (size_t)k * hw + yy * a.w + xx);
  }
  __syncthreads();
  const int Ylo = geo[0], Yhi = geo[1], Xlo = geo[2], Xhi = geo[3];
  const int FH = Yhi - Ylo + 1, FW = Xhi - Xlo + 1;        // <= fh_max, fw_max (host-side bound)
  // tap tables: for low-resolution row y0 + i the label-resolution rows with a non-zero weight are contiguous
  if (threadIdx.x < 2 * SLT_T) {
    const bool isy = threadIdx.x < SLT_T;
    const int i = isy ? threadIdx.x : threadIdx.x - SLT_T;
    const int t = (isy ? y0 : x0) + i, lo = isy ? Ylo : Xlo, hi = isy ? Yhi : Xhi;
    const float sc = isy ? a.sy : a.sx;
    const int in = isy ? a.h : a.w;
    float* wt = (isy ? sWy : sWx) + i * SLT_MAXTAP;
    int first = -1, cnt = 0;
    if (t < in) {
      for (int D = lo; D <= hi; ++D) {      // the weight is a tent in D: its support is one contiguous run
        const float wv = tap_weight(D, sc, in, t);
        if (wv != 0.f) {
          if (first < 0) first = D;
          if (cnt < SLT_MAXTAP) wt[cnt++] = wv;
        } else if (first >= 0) {
          break;
        }
      }
    }
    (isy ? tapY : tapX)[i][0] = first; (isy ? tapY : tapX)[i][1] = cnt;
  }
  // (1) dL/dz of every footprint pixel, once
  for (int f = threadIdx.x; f < FH * FW; f += SLT_THREADS) {
    const int Y = Ylo + f / FW, X = Xlo + f % FW;
    const float fy = a.sy * (float)Y, fx = a.sx * (float)X;
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = iy0 + ((iy0 < a.h - 1) ? 1 : 0), ix1 = ix0 + ((ix0 < a.w - 1) ? 1 : 0);
    const float ly = fy - (float)iy0, lx = fx - (float)ix0, hy = 1.f - ly, hx = 1.f - lx;
    const int r0 = (iy0 - (y0 - 1)) * LT, r1 = (iy1 - (y0 - 1)) * LT, c0 = ix0 - (x0 - 1), c1 = ix1 - (x0 - 1);
    const int lab = (a.labels != nullptr) ? a.labels[((size_t)n * a.H + Y) * a.W + X] : 255;
    float e[OP];
    float mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      if (k < O) {
        const float* q = sL + k * LT * LT;
        e[k] = hy * (hx * q[r0 + c0] + lx * q[r0 + c1]) + ly * (hx * q[r1 + c0] + lx * q[r1 + c1]);
        mx = fmaxf(mx, e[k]);
      } else {
        e[k] = -3.0e38f;
      }
    }
    float Z = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      e[k] = (k < O) ? __expf(e[k] - mx) : 0.f;
      Z += e[k];
    }
    const float iz = 1.f / Z;
    const float ce_on = (lab != 255 && lab < O) ? s_ce : 0.f;
    float qs = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      const float pk = e[k] * iz;
      e[k] = pk;
      if (k < O && pk < a.threshold) qs += pk / (1.f - pk + 1e-6f);
    }
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      if (k < O) {
        const float pk = e[k];
        float gk = ce_on * (pk - ((k == lab) ? 1.f : 0.f));
        const float own = (pk < a.threshold) ? pk / (1.f - pk + 1e-6f) : 0.f;
        gk += s_neg * (own - pk * qs);
        sG[(size_t)k * fh_max * fw_max + f] = gk;
      }
    }
  }
  __syncthreads();
  // (2) every (low-resolution pixel, class) of the tile sums its taps (rows outer, columns inner: fixed order)
  for (int it = threadIdx.x; it < SLT_T * SLT_T * O; it += SLT_THREADS) {
    const int k = it / (SLT_T * SLT_T), r = it - k * SLT_T * SLT_T;
    const int i = r / SLT_T, j = r - i * SLT_T;
    const int y = y0 + i, x = x0 + j;
    if (y >= a.h || x >= a.w) continue;
    const int Yf = tapY[i][0], ny = tapY[i][1], Xf = tapX[j][0], nx = tapX[j][1];
    const float* g = sG + (size_t)k * fh_max * fw_max;
    float acc = 0.f;
    for (int ty = 0; ty < ny; ++ty) {
      const float wy = sWy[i * SLT_MAXTAP + ty];
      if (wy == 0.f) continue;
      const float* gr = g + (Yf + ty - Ylo) * FW + (Xf - Xlo);
      float row = 0.f;
      for (int tx = 0; tx < nx; ++tx) row = fmaf(sWx[j * SLT_MAXTAP + tx], gr[tx], row);
      acc = fmaf(wy, row, acc);
    }
    a.dlogits[((size_t)n * O + k) * hw + y * a.w + x] = acc;
  }
}

// host-side bound of the footprint of one tile; smem = 0 when the tiled kernel does not apply
static SltGeom slt_geometry(const SegLossArgs& a) {
  SltGeom g = {0, 0, 0};
  if (a.sy <= 0.f || a.sx <= 0.f || a.N > 65535) return g;
  const int taps_y = (int)ceilf(2.f / a.sy) + 2, taps_x = (int)ceilf(2.f / a.sx) + 2;
  if (taps_y > SLT_MAXTAP || taps_x > SLT_MAXTAP) return g;
  g.fh_max = (int)ceilf((float)(SLT_T + 1) / a.sy) + 4;
  g.fw_max = (int)ceilf((float)(SLT_T + 1) / a.sx) + 4;
  const size_t LT = SLT_T + 2;
  g.smem = ((size_t)a.O * LT * LT + (size_t)a.O * g.fh_max * g.fw_max + 2 * SLT_T * SLT_MAXTAP) * sizeof(float);
  if (g.smem > 200 * 1024) g.smem = 0;
  return g;
}

template <int OP>
static int launch_seg_loss(const SegLossArgs& a, int blocks_a, int blocks_b, cudaStream_t st) {
  seg_loss_sums_kernel<OP><<<blocks_a, SL_THREADS, 0, st>>>(a);
  int rc = launch_status("seg_loss_sums_kernel");
  if (rc) return rc;
  seg_loss_finish_kernel<<<1, 32, 0, st>>>(a, blocks_a);
  rc = launch_status("seg_loss_finish_kernel");
  if (rc) return rc;
  if (a.dlogits != nullptr) {
    static const bool force_gather = (getenv("HALO_LOSS_GATHER") != nullptr);      // A/B knob: the untiled gather
    const SltGeom g = slt_geometry(a);
    if (g.smem != 0 && !force_gather) {
      HALO_CUDA(cudaFuncSetAttribute(seg_loss_grad_tiled_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
      seg_loss_grad_tiled_kernel<OP><<<dim3((a.w + SLT_T - 1) / SLT_T, (a.h + SLT_T - 1) / SLT_T, a.N), SLT_THREADS, g.smem, st>>>(
          a, g.fh_max, g.fw_max);
      rc = launch_status("seg_loss_grad_tiled_kernel");
    } else {
      seg_loss_grad_kernel<OP><<<blocks_b, SL_THREADS, 0, st>>>(a);
      rc = launch_status("seg_loss_grad_kernel");
    }
  }
  return rc;
}

static int seg_loss_blocks(long long items) {
  long long b = (items + SL_THREADS - 1) / SL_THREADS;
  const long long cap = (long long)sm_count() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_seg_loss_workspace_bytes(void) { return ((size_t)sm_count() * 8 * 4 + 4) * sizeof(double); }

extern "C" int halo_seg_loss(const float* logits_lr, const uint8_t* labels, float neg_weight, float threshold, float* losses,
                             float* dlogits_lr, int N, int O, int h, int w, int H, int W, void* ws, size_t ws_bytes,
                             halo_stream_t stream) {
  HALO_CHECK_ARG(logits_lr && losses, "halo_seg_loss: NULL pointer");
  HALO_CHECK_ARG(N > 0 && O > 0 && h > 0 && w > 0 && H > 0 && W > 0, "halo_seg_loss: bad dims");
  HALO_CHECK_ARG(neg_weight >= 0.f && threshold >= 0.f, "halo_seg_loss: negative weight / threshold");
  if (O > SL_MAXO) {
    set_error("halo_seg_loss: num_classes %d > %d not compiled", O, SL_MAXO);
    return HALO_ERR_UNSUPPORTED;
  }
  const int blocks_a = seg_loss_blocks((long long)N * H * W), blocks_b = seg_loss_blocks((long long)N * h * w);
  const size_t need = ((size_t)blocks_a * 4 + 4) * sizeof(double);
  if (!ws || ws_bytes < need) {
    set_error("halo_seg_loss: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  SegLossArgs a;
  a.logits = logits_lr; a.labels = labels; a.partial = (double*)ws; a.sums = (double*)ws + (size_t)blocks_a * 4;
  a.losses = losses; a.dlogits = dlogits_lr;
  a.N = N; a.O = O; a.h = h; a.w = w; a.H = H; a.W = W;
  a.sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  a.sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  a.neg_weight = neg_weight; a.threshold = threshold;
  cudaStream_t st = (cudaStream_t)stream;
  const int OP = head_op_pad(O);
  switch (OP) {
    case 4: return launch_seg_loss<4>(a, blocks_a, blocks_b, st);
    case 8: return launch_seg_loss<8>(a, blocks_a, blocks_b, st);
    case 12: return launch_seg_loss<12>(a, blocks_a, blocks_b, st);
    case 16: return launch_seg_loss<16>(a, blocks_a, blocks_b, st);
    case 20: return launch_seg_loss<20>(a, blocks_a, blocks_b, st);
    case 24: return launch_seg_loss<24>(a, blocks_a, blocks_b, st);
    case 28: return launch_seg_loss<28>(a, blocks_a, blocks_b, st);
    default: return launch_seg_loss<32>(a, blocks_a, blocks_b, st);
  }
}
