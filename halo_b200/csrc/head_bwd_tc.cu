// K4a-TC -- pixel pass of the fused head backward on the Blackwell tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   du[px][c] = alpha[px] * u[px][c] + sum_n G[px][n] * W[n][c]      G = [gS | gT] (analytic epilogue derivative)
//
// Two chained GEMMs per 128-pixel tile, both 3xTF32 with the pixel operand in TMEM (thread = pixel = TMEM lane):
//   MMA1  [128 px x C] . [C x NP]   recompute S, T            A = features (converter warps), B = parameter planes, K-major
//   MMA2  [128 px x NR] . [NR x C]  the du contraction        A = G, written by the derivative warps IN PLACE over the S, T
//                                   accumulators (hi over "main", lo over "correction"); B = a second, transposed
//                                   K-major copy of the parameter planes ([n/4][c][n%4]); N = C (<= 256) TMEM columns.
//                                   (Reading the forward planes through an MN-major descriptor instead was tried first:
//                                   with the MN-major bit set kind::tf32 returned all-zero accumulators -- profiles/r1_k4.md.)
// One persistent 512-thread CTA per SM; the S,T/G region of TMEM is double-buffered by tile parity so that MMA1 of tile
// i+1 overlaps the derivative math of tile i:
//   warp 0   TMA producer            warp 1   MMA1 issuer          warp 3   MMA2 issuer
//   warps 4-7    converters: u -> (hi, lo) -> TMEM, |u|^2
//   warps 8-11   derivative warps: S,T from TMEM four classes at a time (a ROLLED loop: the fully unrolled 20-class
//                body was 128 KB of SASS and ran at IPC 0.15 on instruction-cache misses), dlogits -> gS, gT, alpha,
//                class scalars; G -> TMEM (+ fp32 planes for K4b); then the output pass of the lower quarter of the
//                channels of the same tile
//   warps 12-15  output warps: D2 from TMEM, + alpha*u (u re-read, two channel groups of loads in flight), store du for
//                the upper three quarters of the channels
// Control warps run warp-wide and predicate the TMA / tcgen05 instructions on an elected lane (tc_common.cuh).
// The weight gradient dW = G^T.U is K4b (head_bwd_dw_tc.cu); the finalisation (K4c) stays in head_bwd.cu.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"
#include "tc_common.cuh"

namespace halo {

constexpr int BT_BM = 128, BT_BK = 32, BT_HK = 16;
constexpr int BT_STAGES = 6;
constexpr int BT_STAGE_FLOATS = BT_BK * BT_BM;
constexpr int BT_THREADS = 512;
constexpr int BT_TMEM_COLS = 512;
// TMEM columns: A halves [0,64) | S,T -> G region of even tiles [64, 64+2NP) | of odd tiles [.., +2NP) | D2 [256, 256+C)
constexpr int BT_A_COL = 0, BT_FG_COL = 64, BT_D2_COL = 256;

// i-th tile of CTA b (G CTAs): every CTA takes RUNS of BT_RUN consecutive 128-pixel tiles (b*R .. b*R+R-1, then +R*G ...),
// so it touches R*512 contiguous bytes of every channel row within a short window (DRAM page locality of the feature
// reads, the re-reads and the du stores), like K1's adjacent-tile mapping.
#ifndef HALO_BT_RUN
#define HALO_BT_RUN 2
#endif
constexpr int BT_RUN = HALO_BT_RUN;
__device__ __forceinline__ int bt_tile_of(int i, int b, int G) { return (i / BT_RUN) * (BT_RUN * G) + b * BT_RUN + (i % BT_RUN); }
__device__ __forceinline__ int bt_my_tiles(int total, int b, int G) {
  const int per_round = BT_RUN * G;
  const int rounds = total / per_round, rem = total - rounds * per_round;
  const int tail = rem - b * BT_RUN;
  return rounds * BT_RUN + (tail <= 0 ? 0 : (tail >= BT_RUN ? BT_RUN : tail));
}

struct BwdTcArgs {
  const float* feat;
  const float* dlogits;
  float* dfeat;
  float* G;         // [N][2OP][HW] fp32 planes for the weight-gradient kernel
  float* cls_part;  // [grid][3][OP]
  int N, C, O, HW, tiles_per_img, total_tiles;
  HeadConsts hc;
};

struct BtSmem {
  size_t w2_off, ring_off, bar_off, tmem_off, cls_off, n2_off, alpha_off, red_off, total;
  int stages;
};
__host__ __device__ inline BtSmem bt_smem_layout(int NP, int OP, int C) {
  BtSmem L;
  (void)NP;
  const size_t w_bytes = (size_t)2 * (2 * OP) * C * 4;  // forward planes (MMA1) and transposed planes (MMA2): same size, 2*OP stored rows
  L.w2_off = (w_bytes + 1023) / 1024 * 1024;
  L.ring_off = (L.w2_off + w_bytes + 1023) / 1024 * 1024;
  const size_t tail = 4096;                            // barriers, class constants, |u|^2, alpha, reduction slots
  const size_t budget = (size_t)227 * 1024;
  int st = (int)((budget - L.ring_off - tail) / ((size_t)BT_STAGE_FLOATS * 4));
  L.stages = st > BT_STAGES ? BT_STAGES : st;          // 3 stages at C=256/O=19, 6 for small heads
  L.bar_off = L.ring_off + (size_t)L.stages * BT_STAGE_FLOATS * 4;
  const int nbars = 2 * BT_STAGES + 4 + 6 + 2 + 2;
  L.tmem_off = L.bar_off + (size_t)nbars * 8;
  L.cls_off = (L.tmem_off + 16 + 15) / 16 * 16;
  L.n2_off = L.cls_off + (size_t)4 * OP * 4;
  L.alpha_off = L.n2_off + 2 * BT_BM * 4;              // |u|^2 and alpha are double-buffered by tile parity
  L.red_off = L.alpha_off + 2 * BT_BM * 4;
  L.total = L.red_off + (size_t)4 * 3 * OP * 4;
  return L;
}

// du[c][p] = alpha * u[c][p] + D2[p][c] for the channels [c_lo, c_hi) (multiples of 32) of this thread's pixel.
// The features do not depend on D2: the first two channel groups (2 x 32 loads per thread) are requested before waiting
// for the MMA, and two groups stay in flight under the TMEM read / store of the current one.  The re-read misses L2 two
// times out of three (profiles/r1_k4.md), so this stage lives on memory-level parallelism -- which is why BOTH the
// derivative warpgroup (after it has handed G to MMA2) and the output warpgroup run it, on disjoint channel ranges.
__device__ __forceinline__ void bt_output_range(const float* ubase, float* dbase, size_t HW, bool live, uint32_t d2_taddr,
                                                int c_lo, int c_hi, uint64_t* d2_full, uint32_t parity, const float* alpha_ptr) {
  // rebase on the first channel of the range: the loop below then indexes from a literal 0, which is what lets ptxas
  // fold every address into base + immediate * stride (with a run-time start it kept 60 channel indices on the stack)
  ubase += (size_t)c_lo * HW;
  dbase += (size_t)c_lo * HW;
  d2_taddr += (uint32_t)c_lo;
  const int Cn = c_hi - c_lo;
  float ua[32], ub[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) ua[e] = (live && Cn > 0) ? __ldcs(ubase + (size_t)e * HW) : 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) ub[e] = (live && Cn > 32) ? __ldcs(ubase + (size_t)(32 + e) * HW) : 0.f;
  mbar_wait(d2_full, parity);
  tc_fence_after();
  const float alpha = *alpha_ptr;
  for (int c0 = 0; c0 < Cn; c0 += 64) {
    float d[32];
    tmem_ld_x32(d2_taddr + c0, d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (live) __stcs(dbase + (size_t)(c0 + e) * HW, fmaf(alpha, ua[e], d[e]));
    }
    if (c0 + 64 < Cn) {
#pragma unroll
      for (int e = 0; e < 32; ++e) ua[e] = live ? __ldcs(ubase + (size_t)(c0 + 64 + e) * HW) : 0.f;
    }
    if (c0 + 32 >= Cn) break;   // an odd number of 32-channel groups
    tmem_ld_x32(d2_taddr + c0 + 32, d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (live) __stcs(dbase + (size_t)(c0 + 32 + e) * HW, fmaf(alpha, ub[e], d[e]));
    }
    if (c0 + 96 < Cn) {
#pragma unroll
      for (int e = 0; e < 32; ++e) ub[e] = live ? __ldcs(ubase + (size_t)(c0 + 96 + e) * HW) : 0.f;
    }
  }
}

template <int NP, int OP>
__global__ void __launch_bounds__(BT_THREADS, 1)
head_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const BwdTcArgs a, const float* __restrict__ wtc,
                   const float* __restrict__ w2g) {
  static_assert(BT_FG_COL + 4 * NP <= BT_D2_COL, "TMEM column budget");
  constexpr int NR = 2 * OP;                   // stored parameter rows (head_pack_tc_kernel) = K extent of MMA2
  extern __shared__ __align__(1024) unsigned char smem[];
  const int C = a.C;
  const BtSmem L = bt_smem_layout(NP, OP, C);
  float* sW = reinterpret_cast<float*>(smem);
  float* sW2 = reinterpret_cast<float*>(smem + L.w2_off);  // [2][NR/4][C][4]: transposed planes for MMA2
  float* ring = reinterpret_cast<float*>(smem + L.ring_off);
  const int NST = L.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full = bars;
  uint64_t* empty = bars + BT_STAGES;
  uint64_t* a_full = bars + 2 * BT_STAGES;   // [2]
  uint64_t* a_empty = a_full + 2;            // [2]
  uint64_t* facc_full = a_empty + 2;         // [2] MMA1 done for a tile of that parity
  uint64_t* g_full = facc_full + 2;          // [2] G (and alpha) of a tile are in TMEM / smem
  uint64_t* g_empty = g_full + 2;            // [2] MMA2 has consumed G: the region may take the S,T of tile i+2
  uint64_t* d2_full = g_empty + 2;           // MMA2 done
  uint64_t* d2_empty = d2_full + 1;          // output warps have read D2
  uint64_t* alpha_free = d2_empty + 1;       // [2] output warps have read sAlpha of a tile of that parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
  float* sCls = reinterpret_cast<float*>(smem + L.cls_off);
  float* sN2 = reinterpret_cast<float*>(smem + L.n2_off);      // [2][128]
  float* sAlpha = reinterpret_cast<float*>(smem + L.alpha_off);  // [2][128]
  float* sRed = reinterpret_cast<float*>(smem + L.red_off);  // [4 warps][3][OP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int n4 = 2 * NR * C / 4;
    const float4* src = reinterpret_cast<const float4*>(wtc);
    float4* dst = reinterpret_cast<float4*>(sW);
    for (int i = threadIdx.x; i < n4; i += BT_THREADS) dst[i] = src[i];
    const float4* src2 = reinterpret_cast<const float4*>(w2g);
    float4* dst2 = reinterpret_cast<float4*>(sW2);
    for (int i = threadIdx.x; i < n4; i += BT_THREADS) dst2[i] = src2[i];
    const float* csrc = wtc + (size_t)2 * NR * C;
    for (int i = threadIdx.x; i < 4 * OP; i += BT_THREADS) sCls[(i % OP) * 4 + i / OP] = csrc[i];
    for (int i = threadIdx.x; i < 4 * 3 * OP; i += BT_THREADS) sRed[i] = 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1);
      mbar_init(&facc_full[i], 1); mbar_init(&g_full[i], 4); mbar_init(&g_empty[i], 1);
      mbar_init(&alpha_free[i], 4);
    }
    mbar_init(d2_full, 1);   mbar_init(d2_empty, 8);   // derivative + output warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int HW = a.HW;
  const int cpt = C / BT_BK;
  // output pass: derivative warpgroup [0, c_split), output warpgroup [c_split, C).  A quarter, not half: the derivative
  // math itself takes about as long as the output pass of half a tile (r1p: with an even split the derivative warpgroup
  // never waited and the output warpgroup idled 29 % of the time)
#ifndef HALO_BT_SPLIT_DIV
#define HALO_BT_SPLIT_DIV 128
#endif
  const int c_split = (C / HALO_BT_SPLIT_DIV) * 32;
  const int my_tiles = bt_my_tiles(a.total_tiles, blockIdx.x, gridDim.x);
  const uint32_t w_hi = smem_u32(sW), w_lo = smem_u32(sW + (size_t)NR * C);

  if (warp == 0) {
    // =================== TMA producer ===================
    {
      const uint64_t pol = l2_policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = bt_tile_of(i, blockIdx.x, gridDim.x);
        const int n = tile / a.tiles_per_img;
        const int p0 = (tile - n * a.tiles_per_img) * BT_BM;
        for (int j = 0; j < cpt; ++j) {
          mbar_wait(&empty[s], ph ^ 1u);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full[s], BT_STAGE_FLOATS * 4);
            tma_load_2d_hint(ring + (size_t)s * BT_STAGE_FLOATS, &tmap, p0, n * C + j * BT_BK, &full[s], pol);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =================== MMA1 issuer: S, T (one main + one correction accumulator) ===================
    // warp-wide loop, tcgen05 instructions on an elected lane (see elect_one_sync in tc_common.cuh)
    {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t w_hi = __shfl_sync(0xffffffffu, smem_u32(sW), 0), w_lo = w_hi + (uint32_t)NR * C * 4;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(BT_BM >> 4) << 24);
      const uint32_t lbo = NR * 16, sbo = 128;
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1;
        const uint32_t d_main = tb + BT_FG_COL + b * 2 * NP, d_corr = d_main + NP;
        mbar_wait(&g_empty[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);   // MMA2 of tile i-2 has read the region
        tc_fence_after();
        for (int j = 0; j < cpt; ++j) {
          const int ca = i * cpt + j;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&a_full[h], (uint32_t)ca & 1u);
            tc_fence_after();
            const uint32_t a_col = tb + BT_A_COL + h * 2 * BT_HK;
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < BT_HK / 8; ++ks) {
                const uint32_t koff = (uint32_t)((j * BT_BK + h * BT_HK + ks * 8) / 4) * lbo;
                const uint64_t b_hi = make_b_desc(w_hi + koff, lbo, sbo);
                const uint64_t b_lo = make_b_desc(w_lo + koff, lbo, sbo);
                const uint32_t first = (j == 0 && h == 0 && ks == 0) ? 0u : 1u;
                tc_mma_tf32_ts(d_main, a_col + ks * 8, b_hi, idesc, first);
                tc_mma_tf32_ts(d_corr, a_col + BT_HK + ks * 8, b_hi, idesc, first);
                tc_mma_tf32_ts(d_corr, a_col + ks * 8, b_lo, idesc, 1u);
              }
              tc_commit(&a_empty[h]);
            }
            __syncwarp();
          }
        }
        if (elect_one_sync()) tc_commit(&facc_full[b]);
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // =================== MMA2 issuer: D2[128 x C] = G . W ===================
    {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      // D=f32, A=B=tf32, both K-major, N=C, M=128.  B rows = channels (16 B apart), K = n: chunks of 4 n are C*16 B apart
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(BT_BM >> 4) << 24);
      const uint32_t lbo = (uint32_t)C * 16, sbo = 128;
      const uint32_t w2_hi = __shfl_sync(0xffffffffu, smem_u32(sW2), 0), w2_lo = w2_hi + (uint32_t)NR * C * 4;
      const uint32_t d2 = tb + BT_D2_COL;
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1;
        mbar_wait(&g_full[b], (uint32_t)(i >> 1) & 1u);
        mbar_wait(d2_empty, ((uint32_t)i & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t g_col = tb + BT_FG_COL + b * 2 * NP;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < NR / 8; ++ks) {
            const uint64_t b_hi = make_b_desc(w2_hi + 2 * ks * lbo, lbo, sbo);
            const uint64_t b_lo = make_b_desc(w2_lo + 2 * ks * lbo, lbo, sbo);
            const uint32_t g_hi = g_col + ks * 8, g_lo = g_hi + NP;
            tc_mma_tf32_ts(d2, g_hi, b_hi, idesc, ks == 0 ? 0u : 1u);
            tc_mma_tf32_ts(d2, g_lo, b_hi, idesc, 1u);
            tc_mma_tf32_ts(d2, g_hi, b_lo, idesc, 1u);
          }
          tc_commit(&g_empty[b]);
          tc_commit(d2_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =================== converters ===================
    const int wq = warp & 3, m = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      unsigned long long n2 = 0ull;
      for (int j = 0; j < cpt; ++j) {
        const int ca = i * cpt + j;
        mbar_wait(&full[s], ph);
        const float* src = ring + (size_t)s * BT_STAGE_FLOATS + m;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&a_empty[h], ((uint32_t)ca & 1u) ^ 1u);
          tc_fence_after();
          uint32_t hi[BT_HK], lo[BT_HK];
          tc_split16(src + h * BT_HK * BT_BM, BT_BM, hi, lo, n2);
          const uint32_t taddr = tmem_base + lane_addr + BT_A_COL + h * 2 * BT_HK;
          tmem_st_x16(taddr, hi);
          tmem_st_x16(taddr + BT_HK, lo);
          // sN2[i & 1] was last read by the derivative warps of tile i-2, which finished before MMA2 of tile i-2, which
          // MMA1 of this tile waited for (g_empty) before consuming the A buffers this loop has been refilling
          if (j == cpt - 1 && h == 1) sN2[(i & 1) * BT_BM + m] = n2_of(n2);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[h]);
        }
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NST) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // =================== derivative warps (thread = pixel) ===================
    const int wq = warp & 3, m = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const HeadConsts& hc = a.hc;   // operands straight from the constant bank: a register copy pushed this role past 128 registers
    const int O = a.O;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      const int tile = bt_tile_of(i, blockIdx.x, gridDim.x);
      const int n = tile / a.tiles_per_img;
      const int p = (tile - n * a.tiles_per_img) * BT_BM + m;
      const bool live = (p < HW);
      const float* dl = a.dlogits + (size_t)n * O * HW + p;
      float* gpl = a.G + (size_t)n * 2 * OP * HW + p;
      float Gn[4];   // upstream gradients of the next class group (in flight under the math of the current one)
#pragma unroll
      for (int e = 0; e < 4; ++e) Gn[e] = (live && e < O) ? __ldg(dl + (size_t)e * HW) : 0.f;
      mbar_wait(&facc_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const PixelScalarGrads ps = tangent_scalar_grads(sN2[b * BT_BM + m], hc);
      const uint32_t fg = tmem_base + lane_addr + BT_FG_COL + b * 2 * NP;   // main | correction -> G hi | G lo
      float g_gamma = 0.f, g_t2 = 0.f, g_om = 0.f;
#pragma unroll 1
      for (int k0 = 0; k0 < OP; k0 += 4) {
        float Sm[4], Sc[4], Tm[4], Tc[4], Gc[4];
        tmem_ld_x4(fg + k0, Sm);
        tmem_ld_x4(fg + OP + k0, Tm);
        tmem_ld_x4(fg + NP + k0, Sc);
        tmem_ld_x4(fg + NP + OP + k0, Tc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          Gc[e] = Gn[e];
          const int kn = k0 + 4 + e;
          Gn[e] = (live && kn < O) ? __ldg(dl + (size_t)kn * HW) : 0.f;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float gS[4], gT[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + e;
          float d_pp = 0.f, d_an = 0.f, d_pa = 0.f;
          gS[e] = gT[e] = 0.f;
          if (k < O) {   // warp-uniform
            const float4 cl = reinterpret_cast<const float4*>(sCls)[k];
            mlr_logit_grad(Gc[e], Sm[e] + Sc[e], Tm[e] + Tc[e], ps, cl.x, cl.y, cl.z, cl.w, hc, gS[e], gT[e], g_gamma, g_t2,
                           g_om, d_pp, d_an, d_pa);
            // class scalars: fixed-order warp reduction into this warp's slot
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              d_pp += __shfl_xor_sync(0xffffffffu, d_pp, o);
              d_an += __shfl_xor_sync(0xffffffffu, d_an, o);
              d_pa += __shfl_xor_sync(0xffffffffu, d_pa, o);
            }
            if (lane == 0) {
              sRed[(wq * 3 + 0) * OP + k] += d_pp;
              sRed[(wq * 3 + 1) * OP + k] += d_an;
              sRed[(wq * 3 + 2) * OP + k] += d_pa;
            }
          }
        }
        // fp32 G planes for the weight-gradient kernel
        if (live) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            gpl[(size_t)(k0 + e) * HW] = gS[e];
            gpl[(size_t)(OP + k0 + e) * HW] = gT[e];
          }
        }
        // G -> TMEM in place of the S, T it was derived from: hi over the main columns, lo over the correction columns
        uint32_t sh[4], sl[4], th[4], tl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sh[e] = cvt_rna_tf32(gS[e]);
          sl[e] = __float_as_uint(gS[e] - __uint_as_float(sh[e]));
          th[e] = cvt_rna_tf32(gT[e]);
          tl[e] = __float_as_uint(gT[e] - __uint_as_float(th[e]));
        }
        tmem_st_x4(fg + k0, sh);
        tmem_st_x4(fg + OP + k0, th);
        tmem_st_x4(fg + NP + k0, sl);
        tmem_st_x4(fg + NP + OP + k0, tl);
      }
      mbar_wait(&alpha_free[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);   // the output warps have read alpha of tile i-2
      sAlpha[b * BT_BM + m] = 2.f * (g_gamma * ps.dgam + g_t2 * ps.dt2 + g_om * ps.dom);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_full[b]);
      // second job of this warpgroup: the output pass of the lower channels of the same tile (see bt_output_range)
      bt_output_range(a.feat + (size_t)n * C * HW + p, a.dfeat + (size_t)n * C * HW + p, (size_t)HW, live,
                      tmem_base + lane_addr + BT_D2_COL, 0, c_split, d2_full, (uint32_t)i & 1u, &sAlpha[b * BT_BM + m]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
    }
  } else if (warp >= 12) {
    // =================== output warps (thread = pixel): du = D2 + alpha * u ===================
    const int wq = warp & 3, m = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = bt_tile_of(i, blockIdx.x, gridDim.x);
      const int n = tile / a.tiles_per_img;
      const int p = (tile - n * a.tiles_per_img) * BT_BM + m;
      const bool live = (p < HW);
      const float* ubase = a.feat + (size_t)n * C * HW + p;
      float* dbase = a.dfeat + (size_t)n * C * HW + p;
      // channels [c_split, C); the derivative warpgroup takes [0, c_split) of the same tile
      bt_output_range(ubase, dbase, (size_t)HW, live, tmem_base + lane_addr + BT_D2_COL, c_split, C, d2_full, (uint32_t)i & 1u,
                      &sAlpha[(i & 1) * BT_BM + m]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&alpha_free[i & 1]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * OP; i += BT_THREADS) {
    float s = 0.f;
    for (int w = 0; w < 4; ++w) s += sRed[w * 3 * OP + i];
    a.cls_part[(size_t)blockIdx.x * 3 * OP + i] = s;
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BT_TMEM_COLS) : "memory");
  }
}

// transposed parameter planes for MMA2: [2 (hi,lo)][NR/4][C][4], NR = 2*OP, value(c, n) = Wt[c][n] of the CUDA-core pack
__global__ void head_pack_bwd_planes_kernel(const float* __restrict__ std_pack, float* __restrict__ w2, int C, int OP) {
  const int NR = 2 * OP, total = NR * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n4 = i / (C * 4), rem = i - n4 * C * 4;
    const int ch = rem >> 2, n = n4 * 4 + (rem & 3);
    const float w = std_pack[(size_t)ch * NR + n];
    const uint32_t h = cvt_rna_tf32(w);
    w2[i] = __uint_as_float(h);
    w2[(size_t)total + i] = __uint_as_float(cvt_rna_tf32(w - __uint_as_float(h)));
  }
}

int head_pack_bwd_planes_launch(const float* std_pack, float* w2, int C, int OP, cudaStream_t st) {
  head_pack_bwd_planes_kernel<<<(2 * OP * C + 255) / 256, 256, 0, st>>>(std_pack, w2, C, OP);
  return launch_status("head_pack_bwd_planes_kernel");
}

// ---- host side -----------------------------------------------------------------------------------------
bool head_bwd_tc_supported(int C, int O, int H, int W, const void* feat, const void* dfeat) {
  if (C % BT_BK != 0 || C > 256 || C < 2 * BT_BK) return false;  // >= 2 pipeline stages per tile (sN2 hand-over, see converters)
  const int NP = round_up(2 * head_op_pad(O), 16);
  if (BT_FG_COL + 4 * NP > BT_D2_COL) return false;  // O <= 24
  if (((long long)H * W) % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(feat) & 15) != 0) return false;
  const BtSmem L = bt_smem_layout(NP, head_op_pad(O), C);
  if (L.stages < 2 || L.total > 227 * 1024) return false;
  return get_encode_fn() != nullptr;
}

int head_bwd_tc_grid(int N, int HW) {
  const int tiles = ((HW + BT_BM - 1) / BT_BM) * N;
  const int runs = (tiles + BT_RUN - 1) / BT_RUN;
  int g = sm_count();
  return g < runs ? g : runs;
}

template <int NP, int OP>
static int launch_bwd_tc(const CUtensorMap& tmap, const BwdTcArgs& a, const float* wtc, const float* w2, size_t smem, int grid,
                         cudaStream_t st) {
  HALO_CUDA(cudaFuncSetAttribute(head_bwd_tc_kernel<NP, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_bwd_tc_kernel<NP, OP><<<grid, BT_THREADS, smem, st>>>(tmap, a, wtc, w2);
  return launch_status("head_bwd_tc_kernel");
}

// wtc: tensor-core parameter pack (head_pack_tc_kernel layout); G, cls_part: as in head_bwd.cu
int head_bwd_tc_launch(const float* feat, const float* dlogits, float* dfeat, float* G, float* cls_part, const float* std_pack,
                       const float* wtc, float* w2, float c, int N, int C, int O, int H, int W, int grid, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return HALO_ERR_CUDA;
  }
  const int OP = head_op_pad(O), NP = round_up(2 * OP, 16), HW = H * W;
  {
    int rc = head_pack_bwd_planes_launch(std_pack, w2, C, OP, st);
    if (rc) return rc;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)N * C};
  const cuuint64_t gstride[1] = {(cuuint64_t)HW * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BT_BM, (cuuint32_t)BT_BK};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return HALO_ERR_CUDA;
  }
  BwdTcArgs a;
  a.feat = feat; a.dlogits = dlogits; a.dfeat = dfeat; a.G = G; a.cls_part = cls_part;
  a.N = N; a.C = C; a.O = O; a.HW = HW;
  a.tiles_per_img = (HW + BT_BM - 1) / BT_BM;
  a.total_tiles = a.tiles_per_img * N;
  a.hc = make_head_consts(c);
  const BtSmem L = bt_smem_layout(NP, OP, C);
  switch (OP) {
    case 4: return launch_bwd_tc<16, 4>(tmap, a, wtc, w2, L.total, grid, st);
    case 8: return launch_bwd_tc<16, 8>(tmap, a, wtc, w2, L.total, grid, st);
    case 12: return launch_bwd_tc<32, 12>(tmap, a, wtc, w2, L.total, grid, st);
    case 16: return launch_bwd_tc<32, 16>(tmap, a, wtc, w2, L.total, grid, st);
    case 20: return launch_bwd_tc<48, 20>(tmap, a, wtc, w2, L.total, grid, st);
    default: return launch_bwd_tc<48, 24>(tmap, a, wtc, w2, L.total, grid, st);
  }
}

}  // namespace halo
