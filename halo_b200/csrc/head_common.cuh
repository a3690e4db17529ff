// Pieces of the head shared by the forward (head_fwd.cu) and backward (head_bwd.cu) translation units.
#pragma once
#include "common.cuh"

namespace halo {

// classes are padded to a multiple of 4 so that the 2*OP weight columns of a channel are whole float4s
inline int head_op_pad(int O) { return round_up(O, 4); }

// ---- parameter packing -------------------------------------------------------------------------------
// ws layout (floats): Wt[CPAD][2*OP]  (columns [0,OP): -P_k ; [OP,2OP): A_k/max(|A_k|,1e-12)), then
// cls[4][OP] = {pp=|P_k|^2, an=|A_k|, pa=<-P_k,a_hat_k>, Bk=1-c*pp}.  Padded rows/columns are zero.
static __global__ void head_pack_kernel(const float* __restrict__ P, const float* __restrict__ A, float c, int O, int OP,
                                 int C, int CPAD, float* __restrict__ ws, float* __restrict__ stats, int N) {
  const int k = blockIdx.x;
  const int KP = 2 * OP;
  float* cls = ws + (size_t)CPAD * KP;
  __shared__ double red[3][32];
  double pp = 0.0, aa = 0.0, pa = 0.0;
  if (k < O) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      double p = (double)P[(size_t)k * C + ch], a = (double)A[(size_t)k * C + ch];
      pp += p * p;
      aa += a * a;
      pa -= p * a;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    pp += __shfl_xor_sync(0xffffffffu, pp, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    pa += __shfl_xor_sync(0xffffffffu, pa, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = pp; red[1][warp] = aa; red[2][warp] = pa; }
  __syncthreads();
  pp = aa = pa = 0.0;
  for (int w = 0; w < nw; ++w) { pp += red[0][w]; aa += red[1][w]; pa += red[2][w]; }
  const double an = sqrt(aa);
  const double den = an > 1e-12 ? an : 1e-12;  // F.normalize eps (hyperbolic.py:173)
  for (int ch = threadIdx.x; ch < CPAD; ch += blockDim.x) {
    float q = 0.f, ah = 0.f;
    if (k < O && ch < C) {
      q = -P[(size_t)k * C + ch];
      ah = (float)((double)A[(size_t)k * C + ch] / den);
    }
    ws[(size_t)ch * KP + k] = q;
    ws[(size_t)ch * KP + OP + k] = ah;
  }
  if (threadIdx.x == 0) {
    cls[0 * OP + k] = (k < O) ? (float)pp : 0.f;
    cls[1 * OP + k] = (k < O) ? (float)an : 0.f;
    cls[2 * OP + k] = (k < O) ? (float)(pa / den) : 0.f;
    cls[3 * OP + k] = (k < O) ? (float)(1.0 - (double)c * pp) : 1.f;
  }
  if (stats != nullptr && k == 0) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      stats[4 * n + 0] = __int_as_float(0x7f800000);  // +inf: running min of a non-negative plane
      stats[4 * n + 1] = 0.f;                         // running max
      stats[4 * n + 2] = 0.f;
      stats[4 * n + 3] = 0.f;
    }
  }
}


// ---- per-pixel epilogue (shared by every head kernel) ---------------------------------------------------
// single-MUFU approximations (flush-to-zero, no denormal/range fix-up code around them); every argument they
// see in the epilogue is a normal number by construction (D >= 1e-12, p + 1e-6, x^2 + 1 ...)
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ln(float x) { return fast_lg2(x) * 0.69314718056f; }
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.44269504089f); }
struct PixelScalars {
  float gamma;   // x = gamma * u
  float t2;      // c*|x|^2
  float omega;   // 1 - c*|x|^2
  float radius;  // (2/s) artanh(s|x|)
  float xnorm;   // |x|
};

// raw features: expmap0 + project fused (hyperbolic.py:37-38 with geoopt's fp64 eps 1e-5).  Branch-free.
__device__ __forceinline__ PixelScalars tangent_scalars(float n2, const HeadConsts& hc) {
  PixelScalars ps;
  const float n = n2 * fast_rsqrt(fmaxf(n2, 1e-30f));  // sqrt(n2), 0 at the origin
  const float sn = hc.s * n;
  const bool clipped = sn > hc.z_clip;  // tanh(min(sn,15)) > 1-1e-5
  const float z = fminf(sn, hc.z_clip);
  const float e = fast_exp(-2.f * z);
  const float ope = 1.f + e;
  const float inv_ope = fast_rcp(ope);
  // tanh: odd series below 0.1 (1 - e would cancel), (1-e)/(1+e) above
  const float z2 = z * z;
  const float t_small = z * fmaf(z2, fmaf(z2, fmaf(z2, -17.f / 315.f, 2.f / 15.f), -1.f / 3.f), 1.f);
  const float t_big = (1.f - e) * inv_ope;
  const float t = clipped ? hc.t_clip : (z < 0.1f ? t_small : t_big);
  ps.omega = clipped ? hc.omega_clip : 4.f * e * inv_ope * inv_ope;  // sech^2(z): never form 1 - t^2
  ps.gamma = t * fast_rcp(hc.s * fmaxf(n, 1e-15f));
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

// asinh through the SFU, branch-free: sign(x) * ln(|x| + sqrt(x^2+1)); absolute error ~2e-7, which the logit
// tolerance (1e-5 of max|logit|) absorbs with a wide margin (DESIGN.md "K1 numerics").  One lg2: beyond 1e9 (x^2
// overflows past ~1.8e19) the ARGUMENT of the logarithm is switched to 2|x| instead of computing two logarithms.
__device__ __forceinline__ float fast_asinh(float x) {
  const float ax = fabsf(x);
  const float v = fmaf(ax, ax, 1.f);
  const float t = (ax > 1e9f) ? 2.f * ax : ax + v * fast_rsqrt(v);
  return copysignf(fast_ln(t), x);
}

// HyperMLR logit for one class from the two contractions (hyperbolic.py:146-183).  Both sides of the MLR-ball
// projection share ONE reciprocal square root: with D > 0, omc = bo/D and m = (1 - omc)/c,
//   inside  (omc >= om_max  <=>  bo >= om_max*D):  arg = 2s * num / den,                den = max(bo, 1e-12 D)   (:179-180)
//   outside (projected to maxnorm, :162-170):      arg = out_scale * num / (D sqrt(m)),  (D sqrt(m))^2 = D (D - bo) / c
// so arg = k * num * rsqrt(X2) with (k, X2) selected per side -- 3 MUFU per class (this rsqrt, asinh's rsqrt and lg2)
// instead of 6; the epilogue warpgroup was the busiest role of K1 (profiles/r1_k1.md r1o).
__device__ __forceinline__ float mlr_logit(float S, float T, const PixelScalars& ps, float pp, float an, float pa,
                                           float Bk, const HeadConsts& hc) {
  const float px = ps.gamma * S;
  const float xa = ps.gamma * T;
  const float cpx2 = 2.f * hc.c * px;
  const float Anum = 1.f + cpx2 + ps.t2;                                  // :150
  const float D = fmaxf(fmaf(hc.c * ps.t2, pp, 1.f + cpx2), 1e-12f);      // :152-153
  const float num = fmaf(Bk, xa, Anum * pa);                              // D * <(-p)(+)x, a_hat>   (:175-177)
  const float bo = Bk * ps.omega;                                         // D * (1 - c*|(-p)(+)x|^2)
  const bool inside = bo >= hc.om_max * D;
  const float den = fmaxf(bo, 1e-12f * D);
  const float x2_out = fmaxf(D * fmaxf(D - bo, 0.f) * hc.inv_c, D * D * 1e-24f);   // m clamped at 1e-24 as before
  const float x2 = inside ? den * den : x2_out;
  const float k = inside ? hc.two_s : hc.out_scale;
  const float arg = num * k * fast_rsqrt(fmaxf(x2, 1e-36f));
  return hc.two_over_s * an * fast_asinh(arg);                            // :181-183 (lambda_term = 2.0)
}

// ---- the same epilogue for TWO classes at a time on the packed fp32x2 pipe (sm_100: FFMA2 / FADD2 / FMUL2) ------------
// Every multiply / add of mlr_logit above becomes one packed instruction for the class pair; comparisons, selects, min /
// max and the MUFU calls stay scalar per half.  Same operations in the same order per component, so each half is
// bit-identical to mlr_logit.  The epilogue warpgroup is K1's critical role (profiles/r2_k1.md): ~45 -> ~25 issue slots
// per class.
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 mlr_logit2(float2 S, float2 T, const PixelScalars& ps, float2 pp, float2 an, float2 pa,
                                             float2 Bk, const HeadConsts& hc) {
  const float2 g2 = f2(ps.gamma);
  const float2 px = __fmul2_rn(g2, S);
  const float2 xa = __fmul2_rn(g2, T);
  const float2 cpx2 = __fmul2_rn(f2(2.f * hc.c), px);
  const float2 Anum = __fadd2_rn(__fadd2_rn(f2(1.f), cpx2), f2(ps.t2));                      // :150
  float2 D = __ffma2_rn(f2(hc.c * ps.t2), pp, __fadd2_rn(f2(1.f), cpx2));                   // :152-153
  D.x = fmaxf(D.x, 1e-12f); D.y = fmaxf(D.y, 1e-12f);
  const float2 num = __ffma2_rn(Bk, xa, __fmul2_rn(Anum, pa));                              // :175-177
  const float2 bo = __fmul2_rn(Bk, f2(ps.omega));
  const float2 thr = __fmul2_rn(f2(hc.om_max), D);
  const bool in0 = bo.x >= thr.x, in1 = bo.y >= thr.y;
  const float2 dlo = __fmul2_rn(f2(1e-12f), D);
  const float2 den = make_float2(fmaxf(bo.x, dlo.x), fmaxf(bo.y, dlo.y));
  float2 dmb = __fadd2_rn(D, make_float2(-bo.x, -bo.y));
  dmb.x = fmaxf(dmb.x, 0.f); dmb.y = fmaxf(dmb.y, 0.f);
  const float2 xo_a = __fmul2_rn(__fmul2_rn(D, dmb), f2(hc.inv_c));
  const float2 xo_b = __fmul2_rn(__fmul2_rn(D, D), f2(1e-24f));
  const float2 dd = __fmul2_rn(den, den);
  const float x2a = in0 ? dd.x : fmaxf(xo_a.x, xo_b.x);
  const float x2b = in1 ? dd.y : fmaxf(xo_a.y, xo_b.y);
  const float2 kk = make_float2(in0 ? hc.two_s : hc.out_scale, in1 ? hc.two_s : hc.out_scale);
  const float2 rs = make_float2(fast_rsqrt(fmaxf(x2a, 1e-36f)), fast_rsqrt(fmaxf(x2b, 1e-36f)));
  const float2 arg = __fmul2_rn(__fmul2_rn(num, kk), rs);
  // asinh, two at a time
  const float2 ax = make_float2(fabsf(arg.x), fabsf(arg.y));
  const float2 v = __ffma2_rn(ax, ax, f2(1.f));
  const float2 vr = make_float2(fast_rsqrt(v.x), fast_rsqrt(v.y));
  const float2 tm = __ffma2_rn(v, vr, ax);                                                   // ax + v * rsqrt(v)
  const float ta = (ax.x > 1e9f) ? 2.f * ax.x : tm.x, tb = (ax.y > 1e9f) ? 2.f * ax.y : tm.y;
  const float2 ln = __fmul2_rn(make_float2(fast_lg2(ta), fast_lg2(tb)), f2(0.69314718056f));
  const float2 ash = make_float2(copysignf(ln.x, arg.x), copysignf(ln.y, arg.y));
  return __fmul2_rn(__fmul2_rn(f2(hc.two_over_s), an), ash);                                 // :181-183
}

// entropy-only softmax over OP/2 class pairs (the packed twin of softmax_entropy_only below)
template <int OP>
__device__ __forceinline__ float softmax_entropy_only2(const float2 (&l)[OP / 2], const HeadConsts& hc) {
  float mx = fmaxf(l[0].x, l[0].y);
#pragma unroll
  for (int j = 1; j < OP / 2; ++j) mx = fmaxf(mx, fmaxf(l[j].x, l[j].y));
  float2 e[OP / 2];
  float2 Z = f2(0.f);
  const float2 nmx = f2(-mx);
#pragma unroll
  for (int j = 0; j < OP / 2; ++j) {
    const float2 d = __fmul2_rn(__fadd2_rn(l[j], nmx), f2(1.44269504089f));
    e[j] = make_float2(fast_ex2(d.x), fast_ex2(d.y));
    Z = __fadd2_rn(Z, e[j]);
  }
  const float2 iz = f2(fast_rcp(Z.x + Z.y));
  float2 ent = f2(0.f);
#pragma unroll
  for (int j = 0; j < OP / 2; ++j) {
    const float2 p = __fmul2_rn(e[j], iz);
    const float2 q = __fadd2_rn(p, f2(1e-6f));
    const float2 lg = make_float2(fast_lg2(q.x), fast_lg2(q.y));
    ent = __ffma2_rn(make_float2(-p.x, -p.y), lg, ent);
  }
  return (ent.x + ent.y) * (0.69314718056f * hc.inv_log19);
}

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
    e[k] = (k < O) ? fast_exp(l[k] - mx) : 0.f;
    Z += e[k];
  }
  const float iz = fast_rcp(Z);
  const int gtf = (gt == 255) ? arg : gt;
  if (pixunc_mode == HALO_PIXUNC_ENTROPY) {
    float ent = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k) {
      const float p = e[k] * iz;
      if (k < O) ent -= p * fast_lg2(p + 1e-6f);
    }
    pixunc = ent * (0.69314718056f * hc.inv_log19);
  } else {
    float pg = 0.f;
#pragma unroll
    for (int k = 0; k < OP; ++k)
      if (k == gtf) pg = e[k] * iz;
    pixunc = 1.f - pg;
  }
  label = (label_mode == HALO_LABEL_GT_FILLED) ? gtf : arg;
}

// the acquisition path's common case: entropy only, no label / ground truth.  No arg-max bookkeeping; padded classes carry
// logit -3e38 (their exp and p*log p terms vanish by themselves), so the loops run without per-class predicates
// (~60 instructions fewer per pixel at 19 classes: the epilogue warps are K1's critical role, profiles/r2_k1.md)
template <int OP>
__device__ __forceinline__ float softmax_entropy_only(const float (&l)[OP], const HeadConsts& hc) {
  float mx = l[0];
#pragma unroll
  for (int k = 1; k < OP; ++k) mx = fmaxf(mx, l[k]);
  float e[OP];
  float Z = 0.f;
#pragma unroll
  for (int k = 0; k < OP; ++k) {
    e[k] = fast_exp(l[k] - mx);
    Z += e[k];
  }
  const float iz = fast_rcp(Z);
  float ent = 0.f;
#pragma unroll
  for (int k = 0; k < OP; ++k) {
    const float p = e[k] * iz;
    ent -= p * fast_lg2(p + 1e-6f);
  }
  return ent * (0.69314718056f * hc.inv_log19);
}

// ---- analytic derivative of the epilogue (shared by the CUDA-core and tensor-core backward kernels) -----------
struct PixelScalarGrads {
  float gamma, t2, omega;   // as in PixelScalars
  float dgam, dt2, dom;     // d/d(n2) of gamma, t2 = c|x|^2, omega = 1 - c|x|^2
};

__device__ __forceinline__ PixelScalarGrads tangent_scalar_grads(float n2, const HeadConsts& hc) {
  PixelScalarGrads g;
  const float nn = sqrtf(n2);
  const float nsafe = fmaxf(nn, 1e-15f);
  const float sn = hc.s * nn;
  const bool clipped = sn > hc.z_clip;
  const float z = fminf(sn, hc.z_clip);
  const float e = expf(-2.f * z);
  const float t = clipped ? hc.t_clip : tanhf(z);
  const float ope = 1.f + e;
  g.omega = clipped ? hc.omega_clip : 4.f * e / (ope * ope);
  g.gamma = t / (hc.s * nsafe);
  g.t2 = t * t;
  if (clipped) {  // project(): x = u/|u| * maxnorm, only the direction carries gradient
    g.dgam = -g.gamma / (2.f * fmaxf(n2, 1e-30f));
    g.dt2 = 0.f;
    g.dom = 0.f;
  } else {
    g.dgam = (z < 1e-3f) ? (-hc.c * (1.f / 3.f)) : (g.omega - g.gamma) / (2.f * fmaxf(n2, 1e-30f));
    g.dt2 = t * hc.s * g.omega / nsafe;
    g.dom = -t * g.omega * hc.s / nsafe;
    if (nn < 1e-15f) { g.dt2 = hc.c; g.dom = -hc.c; }  // limits at the origin
  }
  return g;
}

// One class: upstream gradient G of the logit -> gradients w.r.t. the two contractions (gS, gT), accumulated
// gradients w.r.t. the per-pixel scalars (g_gamma, g_t2, g_om) and the class scalars (d_pp, d_an, d_pa).
// Same single-rsqrt form as mlr_logit: arg = k * num * rsqrt(X2) with
//   inside  (bo >= om_max*D):  k = 2s,        X2 = den^2, den = max(bo, 1e-12 D)   d arg/d bo = -arg/den,  d arg/d D = 0
//   outside (projected):       k = out_scale, X2 = D (D - bo) / c                  d arg/d X2 = -arg / (2 X2),
//                                                                                  d X2/d D = (2D - bo)/c,  d X2/d bo = -D/c
// Branch-free, 3 MUFU per class (this rsqrt, rsqrt(1 + arg^2) shared with asinh, lg2).
__device__ __forceinline__ void mlr_logit_grad(float G, float S, float T, const PixelScalarGrads& ps, float pp, float an,
                                               float pa, float Bk, const HeadConsts& hc, float& gS, float& gT,
                                               float& g_gamma, float& g_t2, float& g_om, float& d_pp, float& d_an,
                                               float& d_pa) {
  const float px = ps.gamma * S, xa = ps.gamma * T;
  const float cpx2 = 2.f * hc.c * px;
  const float Anum = 1.f + cpx2 + ps.t2;
  const float Draw = fmaf(hc.c * ps.t2, pp, 1.f + cpx2);
  const bool dclamp = Draw < 1e-12f;
  const float D = fmaxf(Draw, 1e-12f);
  const float num = fmaf(Bk, xa, Anum * pa);
  const float bo = Bk * ps.omega;
  const bool inside = bo >= hc.om_max * D;
  const float den = fmaxf(bo, 1e-12f * D);
  const float dmb = D - bo;                                   // D * c * m
  const bool mclamp = dmb <= D * 1e-24f * hc.c;               // m clamped at 1e-24: no gradient through it
  const float x2_out = fmaxf(D * fmaxf(dmb, 0.f) * hc.inv_c, D * D * 1e-24f);
  const float x2 = inside ? den * den : x2_out;
  const float rinv = fast_rsqrt(fmaxf(x2, 1e-36f));
  const float a_num = (inside ? hc.two_s : hc.out_scale) * rinv;
  const float arg = num * a_num;
  const float q = arg * rinv * rinv * (0.5f * hc.inv_c);      // arg / (2 c X2)
  const float a_bo = inside ? -arg * rinv : (mclamp ? 0.f : q * D);
  const float a_D = (inside || dclamp || mclamp) ? 0.f : -q * (2.f * D - bo);
  // asinh(arg) and its derivative 1/sqrt(1 + arg^2) share the rsqrt
  const float ax = fabsf(arg);
  const float v = fmaf(ax, ax, 1.f);
  const float rs = fast_rsqrt(v);
  const float ash = copysignf(fast_ln((ax > 1e9f) ? 2.f * ax : ax + v * rs), arg);
  const float gl = G * hc.two_over_s;
  const float g = gl * an * rs;
  const float g_num = g * a_num, g_bo = g * a_bo, g_D = g * a_D;
  const float g_Bk = fmaf(g_num, xa, g_bo * ps.omega);
  const float g_xa = g_num * Bk;
  const float g_Anum = g_num * pa;
  const float g_cpx2 = g_Anum + g_D;
  g_t2 += fmaf(g_D * hc.c, pp, g_Anum);
  g_om = fmaf(g_bo, Bk, g_om);
  const float g_px = 2.f * hc.c * g_cpx2;
  g_gamma += fmaf(g_px, S, g_xa * T);
  gS = g_px * ps.gamma;
  gT = g_xa * ps.gamma;
  d_pp += hc.c * fmaf(g_D, ps.t2, -g_Bk);
  d_an = fmaf(gl, ash, d_an);
  d_pa = fmaf(g_num, Anum, d_pa);
}

// The derivative for TWO classes at a time on the packed fp32x2 pipe (sm_100 FFMA2 / FADD2 / FMUL2): every multiply / add
// of mlr_logit_grad is one packed instruction for the pair; comparisons, selects, min / max and MUFU stay scalar per half.
// g_gamma / g_t2 / g_om accumulate per half (the caller adds the halves), the class-scalar partials come back per half.
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ void mlr_logit_grad2(float2 G, float2 S, float2 T, const PixelScalarGrads& ps, float2 pp, float2 an,
                                                float2 pa, float2 Bk, const HeadConsts& hc, float2& gS, float2& gT,
                                                float2& g_gamma, float2& g_t2, float2& g_om, float2& d_pp, float2& d_an,
                                                float2& d_pa) {
  const float2 gam = f2(ps.gamma);
  const float2 px = __fmul2_rn(gam, S), xa = __fmul2_rn(gam, T);
  const float2 cpx2 = __fmul2_rn(f2(2.f * hc.c), px);
  const float2 one_cpx = __fadd2_rn(f2(1.f), cpx2);
  const float2 Anum = __fadd2_rn(one_cpx, f2(ps.t2));
  const float2 Draw = __ffma2_rn(f2(hc.c * ps.t2), pp, one_cpx);
  const bool dc0 = Draw.x < 1e-12f, dc1 = Draw.y < 1e-12f;
  const float2 D = make_float2(fmaxf(Draw.x, 1e-12f), fmaxf(Draw.y, 1e-12f));
  const float2 num = __ffma2_rn(Bk, xa, __fmul2_rn(Anum, pa));
  const float2 bo = __fmul2_rn(Bk, f2(ps.omega));
  const float2 thr = __fmul2_rn(f2(hc.om_max), D);
  const bool in0 = bo.x >= thr.x, in1 = bo.y >= thr.y;
  const float2 dlo = __fmul2_rn(f2(1e-12f), D);
  const float2 den = make_float2(fmaxf(bo.x, dlo.x), fmaxf(bo.y, dlo.y));
  const float2 dmb = __fadd2_rn(D, neg2(bo));                                   // D * c * m
  const float2 mthr = __fmul2_rn(D, f2(1e-24f * hc.c));
  const bool mc0 = dmb.x <= mthr.x, mc1 = dmb.y <= mthr.y;                       // m clamped at 1e-24: no gradient through it
  const float2 dmp = make_float2(fmaxf(dmb.x, 0.f), fmaxf(dmb.y, 0.f));
  const float2 xo_a = __fmul2_rn(__fmul2_rn(D, dmp), f2(hc.inv_c));
  const float2 xo_b = __fmul2_rn(__fmul2_rn(D, D), f2(1e-24f));
  const float2 dd = __fmul2_rn(den, den);
  const float x2a = in0 ? dd.x : fmaxf(xo_a.x, xo_b.x), x2b = in1 ? dd.y : fmaxf(xo_a.y, xo_b.y);
  const float2 rinv = make_float2(fast_rsqrt(fmaxf(x2a, 1e-36f)), fast_rsqrt(fmaxf(x2b, 1e-36f)));
  const float2 a_num = __fmul2_rn(make_float2(in0 ? hc.two_s : hc.out_scale, in1 ? hc.two_s : hc.out_scale), rinv);
  const float2 arg = __fmul2_rn(num, a_num);
  const float2 ar = __fmul2_rn(arg, rinv);
  const float2 q = __fmul2_rn(__fmul2_rn(ar, rinv), f2(0.5f * hc.inv_c));       // arg / (2 c X2)
  const float2 qD = __fmul2_rn(q, D);
  const float2 a_bo = make_float2(in0 ? -ar.x : (mc0 ? 0.f : qD.x), in1 ? -ar.y : (mc1 ? 0.f : qD.y));
  const float2 aD_raw = __fmul2_rn(neg2(q), __ffma2_rn(f2(2.f), D, neg2(bo)));
  const float2 a_D = make_float2((in0 || dc0 || mc0) ? 0.f : aD_raw.x, (in1 || dc1 || mc1) ? 0.f : aD_raw.y);
  // asinh(arg) and its derivative 1/sqrt(1 + arg^2) share the rsqrt
  const float2 ax = make_float2(fabsf(arg.x), fabsf(arg.y));
  const float2 v = __ffma2_rn(ax, ax, f2(1.f));
  const float2 rs = make_float2(fast_rsqrt(v.x), fast_rsqrt(v.y));
  const float2 tm = __ffma2_rn(v, rs, ax);
  const float ta = (ax.x > 1e9f) ? 2.f * ax.x : tm.x, tb = (ax.y > 1e9f) ? 2.f * ax.y : tm.y;
  const float2 ln = __fmul2_rn(make_float2(fast_lg2(ta), fast_lg2(tb)), f2(0.69314718056f));
  const float2 ash = make_float2(copysignf(ln.x, arg.x), copysignf(ln.y, arg.y));
  const float2 gl = __fmul2_rn(G, f2(hc.two_over_s));
  const float2 g = __fmul2_rn(__fmul2_rn(gl, an), rs);
  const float2 g_num = __fmul2_rn(g, a_num), g_bo = __fmul2_rn(g, a_bo), g_D = __fmul2_rn(g, a_D);
  const float2 g_Bk = __ffma2_rn(g_num, xa, __fmul2_rn(g_bo, f2(ps.omega)));
  const float2 g_xa = __fmul2_rn(g_num, Bk);
  const float2 g_Anum = __fmul2_rn(g_num, pa);
  const float2 g_cpx2 = __fadd2_rn(g_Anum, g_D);
  g_t2 = __fadd2_rn(g_t2, __ffma2_rn(__fmul2_rn(g_D, f2(hc.c)), pp, g_Anum));
  g_om = __ffma2_rn(g_bo, Bk, g_om);
  const float2 g_px = __fmul2_rn(f2(2.f * hc.c), g_cpx2);
  g_gamma = __fadd2_rn(g_gamma, __ffma2_rn(g_px, S, __fmul2_rn(g_xa, T)));
  gS = __fmul2_rn(g_px, gam);
  gT = __fmul2_rn(g_xa, gam);
  d_pp = __fadd2_rn(d_pp, __fmul2_rn(f2(hc.c), __ffma2_rn(g_D, f2(ps.t2), neg2(g_Bk))));
  d_an = __ffma2_rn(gl, ash, d_an);
  d_pa = __ffma2_rn(g_num, Anum, d_pa);
}

// arguments shared by the CUDA-core and tensor-core forward kernels
struct HeadArgs {
  const void* feat;
  const float* ws;
  float* logits;
  float* radius;
  float* pixunc;
  uint8_t* label;
  float* stats;
  float* saved;   // [N][2*OP+1][HW] | NULL: the contractions S_k, T_k (rows k, OP+k; k < O) and |u|^2 (row 2*OP) for the backward
  const uint8_t* gt;
  int pixunc_mode, label_mode, norm_mode;
  int N, C, CPAD, O, HW;
  int tiles_per_img, total_tiles;
  int debug_raw, tc_variant;  // numerics probes (0 in production)
  HeadConsts hc;
};


}  // namespace halo
