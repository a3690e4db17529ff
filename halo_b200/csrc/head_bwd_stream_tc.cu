// K4s -- streaming backward of the fused head on the Blackwell tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// The forward saved the contractions S_k = <u,-p_k>, T_k = <u,a_hat_k> and |u|^2 of every pixel (164 B per pixel at
// O = 19; halo_head_fwd `saved`).  The analytic derivative of the epilogue then needs NO feature data, so the backward
// reads the features exactly ONCE, as a stream of [128 channels x 32 pixels] boxes that feeds both contractions:
//
//   du[px][c]  = alpha[px] * u[px][c] + sum_n G[px][n] * W[n][c]          G = [gS | gT]   (per 128-pixel tile)
//   dW^T[c][n] = sum_px u[c][px] * G[n][px]                                                (accumulated over all tiles)
//
// Round 1 ran these as two kernels that read u three times (MMA1 recompute, the alpha*u pass through L2, the weight
// gradient): 33.5 GB of DRAM traffic for 21.1 GB algorithmic per BASELINE configs[4] step.  Here: 1.08 GB saved planes
// + 0.5 GB dlogits + 6.7 GB features + 6.7 GB du.
//
// One persistent 512-thread CTA per SM, per 128-pixel tile i:
//   warps 8-11 / 12-15  two DERIVATIVE warpgroups on alternate tiles (thread = pixel = TMEM lane): saved S, T, |u|^2 and
//                 dlogits -> gS, gT (kept in registers), alpha, class-scalar partials; once the previous tile has released
//                 the G buffers: G as TF32 hi/lo into TMEM (A operand of the du GEMM) and into shared memory in the
//                 K-major SWIZZLE_128B layout [n][px] (B operand of the dW GEMM).  The same warpgroup then runs the OUTPUT
//                 pass of its tile: warp q takes the stages of pixel chunk q -- D2 from TMEM (its own lanes) + alpha*u with
//                 u read from the stage in shared memory -> coalesced 128-byte du stores.
//   warp 1        du GEMM issuer: D2[128 px x 128 ch] = G . W2 per channel block (3xTF32, A = G from TMEM, B = transposed
//                 parameter planes resident in shared memory), two D2 buffers.
//   warp 0        TMA producer: [128 ch x 32 px] SWIZZLE_128B boxes, channel block major, 6-stage ring at C=256 / O=19.
//   warps 4-7     CONVERTERS (thread = channel = TMEM lane): the stage row -> TF32 hi/lo -> TMEM A buffers; every BS_DRAIN
//                 tiles they add the dW accumulators into this CTA's partial (bounds the round-toward-zero bias of
//                 tcgen05.mma, profiles/r1_tc_numerics.md).
//   warp 2        dW GEMM issuer: acc[128 ch x NP] += U[128 ch x 16 px] . G[16 px x NP] (3xTF32, A from TMEM, B = G in smem).
// Outputs: du; one dW partial [2*OP][CP] and one class-scalar partial [3][OP] per CTA, reduced in a fixed order by
// head_bwd_finalize_kernel (bitwise reproducible on a given device).
#include <cuda.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"
#include "tc_common.cuh"

namespace halo {

constexpr int BS_BM = 128;                        // pixels per tile
constexpr int BS_CHUNK = 32;                      // pixels per stage (128-byte rows)
constexpr int BS_NQ = BS_BM / BS_CHUNK;           // pixel chunks per tile
constexpr int BS_THREADS = 512;
constexpr int BS_STAGE_BYTES = 128 * BS_CHUNK * 4;  // 16 KB: [128 ch x 32 px] (8 KB used when C = 64)
constexpr int BS_MAX_STAGES = 8;
#ifndef HALO_BS_DRAIN
#define HALO_BS_DRAIN 8
#endif
constexpr int BS_DRAIN = HALO_BS_DRAIN;           // tiles per dW accumulator chain (48 MMAs per tile and channel block: 384 per chain)
// TMEM columns (512): G hi|lo [0,96) | A buffer 0 [96,128) | D2 x2 [128,384) | A buffer 1 [384,416) | dW acc x2 [416,512)
constexpr int BS_G_COL = 0, BS_A_COL0 = 96, BS_D2_COL = 128, BS_A_COL1 = 384, BS_ACC_COL = 416;

struct BsArgs {
  const float* dlogits;
  const float* saved;   // [N][2*OP+1][HW]
  float* dfeat;
  float* dw_part;       // [grid][2*OP][CP]
  float* cls_part;      // [grid][3][OP]
  int N, C, CP, O, HW, tiles_per_img, total_tiles, nblk, cb;   // nblk channel blocks of cb (64 | 128) channels
  HeadConsts hc;
};

struct BsSmem {
  size_t g_off, ring_off, bar_off, tmem_off, cls_off, red_off, total;
  int stages, g_plane;   // g_plane: bytes of one G plane (hi or lo): BS_NQ chunks of NP rows of 128 bytes
};
__host__ __device__ inline BsSmem bs_smem_layout(int NP, int OP, int C) {
  BsSmem L;
  const size_t w_bytes = (size_t)2 * (2 * OP) * C * 4;   // transposed parameter planes, hi and lo
  L.g_off = (w_bytes + 1023) / 1024 * 1024;
  L.g_plane = BS_NQ * NP * 128;
  L.ring_off = (L.g_off + 2 * (size_t)L.g_plane + 1023) / 1024 * 1024;
  const size_t tail = 2048;
  const size_t budget = (size_t)227 * 1024;
  int st = (budget > L.ring_off + tail) ? (int)((budget - L.ring_off - tail) / BS_STAGE_BYTES) : 0;
  L.stages = st > BS_MAX_STAGES ? BS_MAX_STAGES : st;
  L.bar_off = L.ring_off + (size_t)L.stages * BS_STAGE_BYTES;
  const int nbars = 2 * BS_MAX_STAGES + 3 + 4 + 4 + 2;
  L.tmem_off = L.bar_off + (size_t)nbars * 8;
  L.cls_off = (L.tmem_off + 16 + 15) / 16 * 16;
  L.red_off = L.cls_off + (size_t)4 * OP * 4;
  L.total = L.red_off + (size_t)8 * 3 * OP * 4;
  return L;
}

__device__ __forceinline__ uint64_t bs_desc_sw128(uint32_t smem_addr) {   // K-major SWIZZLE_128B: 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int NP, int OP>
__global__ void __launch_bounds__(BS_THREADS, 1)
head_bwd_stream_kernel(const __grid_constant__ CUtensorMap tmap, const BsArgs a, const float* __restrict__ w2g,
                       const float* __restrict__ cls_g) {
  constexpr int NR = 2 * OP;
  static_assert(2 * NP <= BS_A_COL0 && BS_ACC_COL + 2 * NP <= 512, "TMEM column budget");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int C = a.C, nblk = a.nblk, cb = a.cb;
  const BsSmem L = bs_smem_layout(NP, OP, C);
  float* sW2 = reinterpret_cast<float*>(smem);                    // [2][NR/4][C][4]
  unsigned char* sG = smem + L.g_off;                             // [2 (hi,lo)][BS_NQ][NP rows][128 B], SW128
  unsigned char* ring = smem + L.ring_off;
  const int NST = L.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full = bars;                          // [8] TMA bytes landed
  uint64_t* empty = bars + BS_MAX_STAGES;         // [8] 4 converter warps + the output warp of the stage's pixel chunk
  uint64_t* g_ready = bars + 2 * BS_MAX_STAGES;   // G of a tile is in TMEM and shared memory
  uint64_t* g_tmem_free = g_ready + 1;            // the du GEMMs of a tile have read G from TMEM
  uint64_t* g_smem_free = g_ready + 2;            // the dW GEMMs of a tile have read G from shared memory
  uint64_t* d2_full = g_ready + 3;                // [2]
  uint64_t* d2_empty = d2_full + 2;               // [2] the four output warps have read the buffer
  uint64_t* a_full = d2_empty + 2;                // [2]
  uint64_t* a_empty = a_full + 2;                 // [2]
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
  // Number of stages the producer has issued so far.  The output warps visit only the stages of their own tiles and pixel
  // chunk, so they do NOT see every phase of a ring slot's `full` barrier; a parity wait is then only sound once the slot's
  // previous use is known to have completed -- which is exactly what "the producer has issued this stage" implies (it waited
  // for the slot's `empty` phase, i.e. for converters that had waited for the previous landing).  Without this gate a warp
  // that races ahead (dead lanes of a ragged tile store nothing) could pass on the stale parity, release the slot early and
  // let the next load overwrite rows a slower consumer had not read yet (seen as rare wrong dP/dA or dfeat, round 2).
  volatile uint32_t* issued = tmem_slot + 1;
  float* sCls = reinterpret_cast<float*>(smem + L.cls_off);
  float* sRed = reinterpret_cast<float*>(smem + L.red_off);       // [8 derivative warps][3][OP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int n4 = 2 * NR * C / 4;
    const float4* src = reinterpret_cast<const float4*>(w2g);
    float4* dst = reinterpret_cast<float4*>(sW2);
    for (int i = threadIdx.x; i < n4; i += BS_THREADS) dst[i] = src[i];
    // class constants per class PAIR: two float4 {pp0, pp1, an0, an1} {pa0, pa1, Bk0, Bk1} (operands of the packed derivative)
    for (int i = threadIdx.x; i < 4 * OP; i += BS_THREADS) {
      const int q = i / OP, k = i % OP;
      sCls[(k >> 1) * 8 + q * 2 + (k & 1)] = cls_g[i];
    }
    for (int i = threadIdx.x; i < 8 * 3 * OP; i += BS_THREADS) sRed[i] = 0.f;
    // rows NR..NP-1 of the G planes are addressed by the dW GEMM (N = NP) and never written: keep them finite
    float4* gz = reinterpret_cast<float4*>(sG);
    for (int i = threadIdx.x; i < 2 * L.g_plane / 16; i += BS_THREADS) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    *issued = 0u;
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 5); }
    mbar_init(g_ready, 4); mbar_init(g_tmem_free, 1); mbar_init(g_smem_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 4);
      mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int HW = a.HW;
  const int my_tiles = ((int)blockIdx.x < a.total_tiles) ? (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // Register budget by role (setmaxnreg, warpgroup-wide, issued at the top of each role's region so that no control-flow
  // merge follows it): the control warps give registers back, the derivative warpgroups -- 40 parked gradients + the
  // class math + the output pass -- take 152 instead of the 128 a 512-thread CTA gets (128 x 80 + 128 x 128 + 256 x 152 = 65 536).
  if (warp < 4) {
#ifndef HALO_BS_NO_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
#endif
  if (warp == 0) {
    // =================== TMA producer ===================
    int s = 0;
    uint32_t ph = 0, n_issued = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int n = tile / a.tiles_per_img;
      const int p0 = (tile - n * a.tiles_per_img) * BS_BM;
      for (int g = 0; g < nblk; ++g) {
        for (int q = 0; q < BS_NQ; ++q) {
          mbar_wait(&empty[s], ph ^ 1u);
          ++n_issued;
          if (elect_one_sync()) {
            *issued = n_issued;      // the slot's previous phase is complete: parity waits on full[s] are sound from here on
            mbar_arrive_expect_tx(&full[s], (uint32_t)(cb * BS_CHUNK * 4));
            tma_load_2d(ring + (size_t)s * BS_STAGE_BYTES, &tmap, p0 + q * BS_CHUNK, n * C + g * 128, &full[s]);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =================== du GEMM issuer: D2[128 px x cb] = G . W2 per channel block ===================
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cb >> 3) << 17) | ((uint32_t)(BS_BM >> 4) << 24);
    const uint32_t lbo = (uint32_t)C * 16, sbo = 128;
    const uint32_t w2_hi = __shfl_sync(0xffffffffu, smem_u32(sW2), 0), w2_lo = w2_hi + (uint32_t)NR * C * 4;
    const uint32_t g_col = tb + BS_G_COL;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(g_ready, (uint32_t)i & 1u);
      for (int g = 0; g < nblk; ++g) {
        const int bc = i * nblk + g, db = bc & 1;
        mbar_wait(&d2_empty[db], (((uint32_t)bc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d2 = tb + BS_D2_COL + db * 128;
        const uint32_t wb = (uint32_t)g * 128u * 16u;   // first channel row of the block inside every K chunk
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < NR / 8; ++ks) {
            const uint64_t b_hi = make_b_desc(w2_hi + wb + 2 * ks * lbo, lbo, sbo);
            const uint64_t b_lo = make_b_desc(w2_lo + wb + 2 * ks * lbo, lbo, sbo);
            const uint32_t g_hi = g_col + ks * 8, g_lo = g_hi + NP;
            tc_mma_tf32_ts(d2, g_hi, b_hi, idesc, ks == 0 ? 0u : 1u);
            tc_mma_tf32_ts(d2, g_lo, b_hi, idesc, 1u);
            tc_mma_tf32_ts(d2, g_hi, b_lo, idesc, 1u);
          }
          tc_commit(&d2_full[db]);
          if (g == nblk - 1) tc_commit(g_tmem_free);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // =================== dW GEMM issuer: acc[g][128 ch x NP] += U . G ===================
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t g_hi0 = __shfl_sync(0xffffffffu, smem_u32(sG), 0), g_lo0 = g_hi0 + (uint32_t)L.g_plane;
    long long hc = 0;   // hand-over counter -> A buffer, phase
    for (int i = 0; i < my_tiles; ++i) {
      const int ci = i % BS_DRAIN;
      if (ci == 0) {
        mbar_wait(acc_empty, (((uint32_t)(i / BS_DRAIN)) & 1u) ^ 1u);   // the previous chain has been drained
        tc_fence_after();
      }
      mbar_wait(g_ready, (uint32_t)i & 1u);
      for (int g = 0; g < nblk; ++g) {
        const uint32_t acc = tb + BS_ACC_COL + g * NP;
        for (int q = 0; q < BS_NQ; ++q) {
          const uint32_t gq_hi = g_hi0 + (uint32_t)(q * NP * 128), gq_lo = g_lo0 + (uint32_t)(q * NP * 128);
#pragma unroll
          for (int h = 0; h < 2; ++h, ++hc) {
            const int bf = (int)(hc & 1);
            mbar_wait(&a_full[bf], (uint32_t)(hc >> 1) & 1u);
            tc_fence_after();
            const uint32_t a_col = tb + (bf ? BS_A_COL1 : BS_A_COL0);
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t koff = (uint32_t)(h * 16 + ks * 8) * 4;   // byte offset of the K slice inside the 128-byte rows
                const uint64_t b_hi = bs_desc_sw128(gq_hi + koff);
                const uint64_t b_lo = bs_desc_sw128(gq_lo + koff);
                const uint32_t first = (ci == 0 && q == 0 && h == 0 && ks == 0) ? 0u : 1u;
                tc_mma_tf32_ts(acc, a_col + ks * 8, b_hi, idesc, first);
                tc_mma_tf32_ts(acc, a_col + 16 + ks * 8, b_hi, idesc, 1u);
                tc_mma_tf32_ts(acc, a_col + ks * 8, b_lo, idesc, 1u);
              }
              tc_commit(&a_empty[bf]);
            }
            __syncwarp();
          }
        }
      }
      if (elect_one_sync()) {
        tc_commit(g_smem_free);
        if (ci == BS_DRAIN - 1 || i == my_tiles - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  }
  } else if (warp < 8) {
    // =================== converters (thread = channel = TMEM lane) ===================
    const int wq = warp & 3, ch = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const bool have = (wq * 32 < cb);       // warp-uniform: C = 64 leaves the upper two warps without channels
    float* out = a.dw_part + (size_t)blockIdx.x * NR * a.CP + ch;
    int s = 0;
    uint32_t ph = 0;
    long long hc = 0;
    for (int i = 0; i < my_tiles; ++i) {
      for (int g = 0; g < nblk; ++g) {
        for (int q = 0; q < BS_NQ; ++q) {
          mbar_wait(&full[s], ph);
          float4 v[8];
          if (have) {
            const unsigned char* row = ring + (size_t)s * BS_STAGE_BYTES + ch * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ (ch & 7)) << 4));
          }
#pragma unroll
          for (int h = 0; h < 2; ++h, ++hc) {
            const int bf = (int)(hc & 1);
            mbar_wait(&a_empty[bf], ((uint32_t)(hc >> 1) & 1u) ^ 1u);
            if (have) {
              tc_fence_after();
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 f = v[h * 4 + j];
                const uint32_t u0 = __float_as_uint(f.x), u1 = __float_as_uint(f.y), u2 = __float_as_uint(f.z), u3 = __float_as_uint(f.w);
                const uint32_t h0 = (u0 + 0x1000u) & 0xffffe000u, h1 = (u1 + 0x1000u) & 0xffffe000u;
                const uint32_t h2 = (u2 + 0x1000u) & 0xffffe000u, h3 = (u3 + 0x1000u) & 0xffffe000u;
                unsigned long long l01, l23;
                asm("sub.f32x2 %0, %1, %2;" : "=l"(l01) : "l"(pack_f32x2(u0, u1)), "l"(pack_f32x2(h0, h1)));
                asm("sub.f32x2 %0, %1, %2;" : "=l"(l23) : "l"(pack_f32x2(u2, u3)), "l"(pack_f32x2(h2, h3)));
                hi[4 * j + 0] = h0; hi[4 * j + 1] = h1; hi[4 * j + 2] = h2; hi[4 * j + 3] = h3;
                asm("mov.b64 {%0, %1}, %2;" : "=r"(lo[4 * j + 0]), "=r"(lo[4 * j + 1]) : "l"(l01));
                asm("mov.b64 {%0, %1}, %2;" : "=r"(lo[4 * j + 2]), "=r"(lo[4 * j + 3]) : "l"(l23));
              }
              const uint32_t taddr = tmem_base + lane_addr + (bf ? BS_A_COL1 : BS_A_COL0);
              tmem_st_x16(taddr, hi);
              tmem_st_x16(taddr + 16, lo);
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&a_full[bf]);
              // The stage is released only here, after the conversion has CONSUMED every loaded register.  Releasing it
              // right behind the eight LDS.128 (nothing had read their results yet) let the arrive overtake the loads once
              // in ~1e9 hand-overs: the next TMA load then rewrote rows this warp had not fetched, and dP / dA came out
              // wrong in a few channels of one CTA (round 2, tools/bwd_catch.py; any added use of v[] hid it).
              if (h == 1) mbar_arrive(&empty[s]);
            }
          }
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
      if ((i % BS_DRAIN) == BS_DRAIN - 1 || i == my_tiles - 1) {
        // drain: add the chain's accumulators into this CTA's partial (fixed order: bitwise reproducible)
        const int chain = i / BS_DRAIN;
        mbar_wait(acc_full, (uint32_t)chain & 1u);
        tc_fence_after();
        if (have) {
          for (int g = 0; g < nblk; ++g) {
            const uint32_t taddr = tmem_base + lane_addr + BS_ACC_COL + g * NP;
            float* o = out + g * 128;
            float prev[NR];   // the partial so far: all loads in flight together (one L2 round trip per drain, not NR)
#pragma unroll
            for (int k = 0; k < NR; ++k) prev[k] = (chain == 0) ? 0.f : __ldcg(o + (size_t)k * a.CP);
#pragma unroll
            for (int c8 = 0; c8 < NR / 8; ++c8) {
              float m8[8];
              tmem_ld_x8(taddr + c8 * 8, m8);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int e = 0; e < 8; ++e) __stcg(o + (size_t)(c8 * 8 + e) * a.CP, prev[c8 * 8 + e] + m8[e]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
    if (my_tiles == 0 && have) {
      for (int g = 0; g < nblk; ++g)
        for (int k = 0; k < NR; ++k) out[(size_t)k * a.CP + g * 128] = 0.f;
    }
  } else if (warp >= 8) {
    // =================== derivative warpgroups (thread = pixel = TMEM lane), then the output pass of the tile ===========
#ifndef HALO_BS_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
#endif
    const int wg = (warp - 8) >> 2;                 // even / odd tiles
    const int wq = warp & 3, m = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const HeadConsts& hc = a.hc;
    const int O = a.O;
    const int SVR = 2 * OP + 1;
    float* red = sRed + (size_t)(warp - 8) * 3 * OP;
    // byte offset of this lane's pixel inside a 128-byte SW128 row whose (row & 7) == r:  ((lane>>2) ^ r) << 4 | (lane&3) << 2
    const uint32_t lane_q = (uint32_t)lane >> 2, lane_e = ((uint32_t)lane & 3u) << 2;
    // The first loads of a tile (|u|^2 and the first class group) are issued one own-tile ahead -- before the output pass of
    // the previous one -- so the derivative math never starts by waiting a DRAM round trip (r2e profile: 10 % of this role).
    const size_t hw = (size_t)HW;
    float n2 = 0.f, Sn[4], Tn[4], Gn[4];
    const float *pG = nullptr, *pS = nullptr, *pT = nullptr;
    auto first_loads = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int n = tile / a.tiles_per_img;
      const int p = (tile - n * a.tiles_per_img) * BS_BM + m;
      // dead lanes of a ragged tile read the image's last pixel (valid memory) and get zero upstream gradients, so every
      // load is unpredicated and every quantity they produce is an exact zero
      const int pc = (p < HW) ? p : HW - 1;
      pG = a.dlogits + (size_t)n * O * HW + pc;          // class k0 .. of the group being prefetched
      pS = a.saved + (size_t)n * SVR * HW + pc;
      pT = pS + (size_t)OP * HW;
      n2 = __ldg(pS + (size_t)(2 * OP) * HW);
#pragma unroll
      for (int e = 0; e < 4; ++e) {      // rows of padded classes (k >= O) were never written by the forward: not read
        Sn[e] = (e < O) ? __ldcs(pS + e * hw) : 0.f;
        Tn[e] = (e < O) ? __ldcs(pT + e * hw) : 0.f;
        Gn[e] = (e < O) ? __ldcs(pG + e * hw) : 0.f;
      }
    };
    if (wg < my_tiles) first_loads(wg);
    for (int i = wg; i < my_tiles; i += 2) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int n = tile / a.tiles_per_img;
      const int p = (tile - n * a.tiles_per_img) * BS_BM + m;
      const bool live = (p < HW);
#ifdef HALO_BS_NO_PREFETCH
      if (i != wg) first_loads(i);
#endif
      const PixelScalarGrads ps = tangent_scalar_grads(n2, hc);
      float2 g_gamma2 = make_float2(0.f, 0.f), g_t22 = g_gamma2, g_om2 = g_gamma2;
      float keepS[OP], keepT[OP];
#ifndef HALO_BS_NO_L2_PREFETCH
      // The in-loop loads below run one class group ahead (~400 cycles of math), less than a DRAM round trip under load
      // (r2i profile: 15 % of all stall samples sat on them).  So the rows of this warpgroup's NEXT tile are pulled into L2
      // while this tile is differentiated -- one prefetch per row, no registers held -- and the loads then cost an L2 hit.
      // Measured (batch 8, profiles/r2_k4.md): C = 128 backward 2.16 -> 1.98 ms, C = 64 1.60 -> 1.59 ms, but C = 256
      // 3.03 -> 3.24 ms (the feature stream already saturates the memory system there), hence C <= 128 only.
      const bool pf = (i + 2 < my_tiles) && (C <= 128);          // warp-uniform
      long long dS = 0, dG = 0;                    // element offsets from this tile's pointers to the next own tile's
      if (pf) {
        const int tile2 = tile + 2 * (int)gridDim.x;
        const int n2i = tile2 / a.tiles_per_img;
        const int p2 = (tile2 - n2i * a.tiles_per_img) * BS_BM + m;
        const int pc = live ? p : HW - 1, pc2 = (p2 < HW) ? p2 : HW - 1;
        dS = (long long)(n2i - n) * SVR * HW + (pc2 - pc);
        dG = (long long)(n2i - n) * O * HW + (pc2 - pc);
      }
#endif
#pragma unroll 1
      for (int k0 = 0; k0 < OP; k0 += 4) {
        float Sc[4], Tc[4], Gc[4];
        pS += 4 * hw; pT += 4 * hw; pG += 4 * hw;
        const bool more = (k0 + 4 < OP);          // warp-uniform: the saved planes hold OP rows, dlogits only O
#ifndef HALO_BS_NO_L2_PREFETCH
        if (pf && more && (lane == 0 || lane == 31)) {      // a warp's 32 pixels of a row are one or two 128-byte lines
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if ((k0 + 4 + e < OP - 3) || (k0 + 4 + e < O)) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pS + e * hw + dS));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pT + e * hw + dS));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pG + e * hw + dG));
            }
          }
        }
#endif
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // padded classes (k >= O) carry zeros, and so does the upstream gradient of dead lanes: every quantity derived
          // from them is then an exact zero
          Sc[e] = Sn[e]; Tc[e] = Tn[e]; Gc[e] = live ? Gn[e] : 0.f;
          if (more) {
            const bool valid = (k0 + 4 + e < OP - 3) || (k0 + 4 + e < O);   // warp-uniform, a compile-time truth but for 3 classes
            Sn[e] = valid ? __ldcs(pS + e * hw) : 0.f;
            Tn[e] = valid ? __ldcs(pT + e * hw) : 0.f;
            Gn[e] = valid ? __ldcs(pG + e * hw) : 0.f;
          }
        }
        float gS[4], gT[4];
        float cs[16];     // class-scalar partials of the group: [class e][d_pp, d_an, d_pa], padded to 16
#pragma unroll
        for (int e = 12; e < 16; ++e) cs[e] = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {             // two class pairs on the packed fp32x2 pipe
          const int kp = (k0 >> 1) + j;
          const float4 c0 = reinterpret_cast<const float4*>(sCls)[2 * kp], c1 = reinterpret_cast<const float4*>(sCls)[2 * kp + 1];
          float2 gs2, gt2, dpp = make_float2(0.f, 0.f), dan = dpp, dpa = dpp;
          mlr_logit_grad2(make_float2(Gc[2 * j], Gc[2 * j + 1]), make_float2(Sc[2 * j], Sc[2 * j + 1]),
                          make_float2(Tc[2 * j], Tc[2 * j + 1]), ps, make_float2(c0.x, c0.y), make_float2(c0.z, c0.w),
                          make_float2(c1.x, c1.y), make_float2(c1.z, c1.w), hc, gs2, gt2, g_gamma2, g_t22, g_om2, dpp, dan, dpa);
          gS[2 * j] = gs2.x; gS[2 * j + 1] = gs2.y; gT[2 * j] = gt2.x; gT[2 * j + 1] = gt2.y;
          cs[6 * j + 0] = dpp.x; cs[6 * j + 1] = dan.x; cs[6 * j + 2] = dpa.x;
          cs[6 * j + 3] = dpp.y; cs[6 * j + 4] = dan.y; cs[6 * j + 5] = dpa.y;
        }
        // warp reduction of the 12 partials by recursive halving: at every step a lane hands half of its values to its
        // partner and adds the partner's other half, so 8 + 4 + 2 + 1 + 1 shuffles replace 12 x 5 (fixed order: bitwise
        // reproducible).  Afterwards lanes 2j and 2j+1 both hold the warp total of value j.
        {
          const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
          float r8[8], r4[4], r2[2], r1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float send = b4 ? cs[j] : cs[j + 8];
            r8[j] = (b4 ? cs[j + 8] : cs[j]) + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float send = b3 ? r8[j] : r8[j + 4];
            r4[j] = (b3 ? r8[j + 4] : r8[j]) + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float send = b2 ? r4[j] : r4[j + 2];
            r2[j] = (b2 ? r4[j + 2] : r4[j]) + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          {
            const float send = b1 ? r2[0] : r2[1];
            r1 = (b1 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
          const int j = lane >> 1;             // value index: class e = j / 3 of the group, scalar t = j % 3
          const int e = j / 3, t = j - 3 * e;
          if ((lane & 1) == 0 && j < 12 && k0 + e < O) red[t * OP + k0 + e] += r1;
        }
        // park the group's gradients in statically indexed registers until the G buffers are free (see below)
#pragma unroll
        for (int gi = 0; gi < OP / 4; ++gi) {
          if (k0 == 4 * gi) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { keepS[4 * gi + e] = gS[e]; keepT[4 * gi + e] = gT[e]; }
          }
        }
      }
      const float g_gamma = g_gamma2.x + g_gamma2.y, g_t2 = g_t22.x + g_t22.y, g_om = g_om2.x + g_om2.y;
      const float alpha = live ? 2.f * (g_gamma * ps.dgam + g_t2 * ps.dt2 + g_om * ps.dom) : 0.f;
      // The G buffers are single: wait until the GEMMs of the previous tile (the other warpgroup's) have read them.
      // (Handing G over earlier -- TMEM first, shared memory chunk by chunk behind per-chunk barriers -- was tried in round 2:
      // no gain, and intermittently wrong dP / dA on ragged images; profiles/r2_k4.md.)
      uint32_t sh[OP], sl[OP], th[OP], tl[OP];
#pragma unroll
      for (int k = 0; k < OP; ++k) {
        sh[k] = cvt_rna_tf32(keepS[k]);
        sl[k] = __float_as_uint(keepS[k] - __uint_as_float(sh[k]));
        th[k] = cvt_rna_tf32(keepT[k]);
        tl[k] = __float_as_uint(keepT[k] - __uint_as_float(th[k]));
      }
      mbar_wait(g_tmem_free, ((uint32_t)i & 1u) ^ 1u);
      mbar_wait(g_smem_free, ((uint32_t)i & 1u) ^ 1u);
      tc_fence_after();
      {
        const uint32_t fg = tmem_base + lane_addr + BS_G_COL;
#pragma unroll
        for (int k4 = 0; k4 < OP / 4; ++k4) {
          tmem_st_x4(fg + 4 * k4, *reinterpret_cast<uint32_t(*)[4]>(&sh[4 * k4]));
          tmem_st_x4(fg + OP + 4 * k4, *reinterpret_cast<uint32_t(*)[4]>(&th[4 * k4]));
          tmem_st_x4(fg + NP + 4 * k4, *reinterpret_cast<uint32_t(*)[4]>(&sl[4 * k4]));
          tmem_st_x4(fg + NP + OP + 4 * k4, *reinterpret_cast<uint32_t(*)[4]>(&tl[4 * k4]));
        }
        // shared-memory copy for the dW GEMM: chunk wq, row n, this lane's pixel (conflict-free: a warp writes one
        // 128-byte row per instruction)
        unsigned char* gq_hi = sG + (size_t)wq * NP * 128;
        unsigned char* gq_lo = gq_hi + L.g_plane;
#pragma unroll
        for (int k = 0; k < OP; ++k) {
          const uint32_t o_s = (uint32_t)k * 128u + (((lane_q ^ (uint32_t)(k & 7)) << 4) | lane_e);
          const uint32_t o_t = (uint32_t)(OP + k) * 128u + (((lane_q ^ (uint32_t)((OP + k) & 7)) << 4) | lane_e);
          *reinterpret_cast<uint32_t*>(gq_hi + o_s) = sh[k];
          *reinterpret_cast<uint32_t*>(gq_lo + o_s) = sl[k];
          *reinterpret_cast<uint32_t*>(gq_hi + o_t) = th[k];
          *reinterpret_cast<uint32_t*>(gq_lo + o_t) = tl[k];
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores of G -> visible to the dW GEMM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_ready);
#ifndef HALO_BS_NO_PREFETCH
      if (i + 2 < my_tiles) first_loads(i + 2);     // in flight under the output pass below
#endif

      // ---- output pass of this tile: warp wq owns pixel chunk wq of every channel block ----
      float* dbase = a.dfeat + (size_t)n * C * HW + p;
      for (int g = 0; g < nblk; ++g) {
        const int sc = (i * nblk + g) * BS_NQ + wq;
        const int s = sc % NST;
        const int bc = i * nblk + g, db = bc & 1;
        while (*issued <= (uint32_t)sc) {}      // see `issued` above
        mbar_wait(&full[s], (uint32_t)(sc / NST) & 1u);
        mbar_wait(&d2_full[db], ((uint32_t)bc >> 1) & 1u);
        tc_fence_after();
        // shared-memory address of this lane's pixel in row 0 of the stage; row c adds c*128 and swaps the 16-byte unit by
        // (c & 7) -- c0 is a multiple of 32, so the swap depends on e alone and every offset below is a literal
        const uint32_t st = smem_u32(ring + (size_t)s * BS_STAGE_BYTES) + lane_e;
        const uint32_t d2 = tmem_base + lane_addr + BS_D2_COL + db * 128;
        float* pd = dbase + (size_t)(g * 128) * HW;
        for (int c0 = 0; c0 < cb; c0 += 32) {
          float uv[32], d[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const uint32_t addr = st + (uint32_t)(c0 + e) * 128u + ((lane_q ^ (uint32_t)(e & 7)) << 4);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(uv[e]) : "r"(addr));
          }
          tmem_ld_x32(d2 + c0, d);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (live) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              __stcs(pd, fmaf(alpha, uv[e], d[e]));
              pd += HW;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&empty[s]);
          mbar_arrive(&d2_empty[db]);
        }
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * OP; i += BS_THREADS) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sRed[w * 3 * OP + i];
    a.cls_part[(size_t)blockIdx.x * 3 * OP + i] = s;
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------
bool head_bwd_stream_supported(int C, int O, int H, int W, const void* feat, const void* dfeat) {
  return head_bwd_stream_shape_ok(C, O, H, W, feat, dfeat) && get_encode_fn() != nullptr;
}

bool head_bwd_stream_shape_ok(int C, int O, int H, int W, const void* feat, const void* dfeat) {
  if (C != 64 && C != 128 && C != 256) return false;
  const int OP = head_op_pad(O), NP = round_up(2 * OP, 16);
  if (NP > 48) return false;                                  // O <= 24: TMEM column budget
  if (((long long)H * W) % 4 != 0) return false;              // TMA global stride must be a multiple of 16 bytes
  if ((reinterpret_cast<uintptr_t>(feat) & 15) != 0 || (reinterpret_cast<uintptr_t>(dfeat) & 3) != 0) return false;
  const BsSmem L = bs_smem_layout(NP, OP, C);
  return L.stages >= 3 && L.total <= 227 * 1024;
}

int head_bwd_stream_grid(int N, int HW) {
  const long long tiles = (long long)N * ((HW + BS_BM - 1) / BS_BM);
  const long long g = sm_count();
  return (int)(g < tiles ? g : tiles);
}

template <int NP, int OP>
static int launch_bs(const CUtensorMap& tmap, const BsArgs& a, const float* w2, const float* cls, size_t smem, int grid,
                     cudaStream_t st) {
  HALO_CUDA(cudaFuncSetAttribute(head_bwd_stream_kernel<NP, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_bwd_stream_kernel<NP, OP><<<grid, BS_THREADS, smem, st>>>(tmap, a, w2, cls);
  return launch_status("head_bwd_stream_kernel");
}

// std_pack: head_pack_kernel layout (Wt[CPAD][2*OP] + cls[4][OP]); w2: scratch for the transposed planes
int head_bwd_stream_launch(const float* feat, const float* dlogits, const float* saved, float* dfeat, float* dw_part,
                           float* cls_part, const float* std_pack, float* w2, float c, int N, int C, int CPAD, int O, int H, int W,
                           int CP, int grid, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return HALO_ERR_CUDA;
  }
  const int OP = head_op_pad(O), NP = round_up(2 * OP, 16), HW = H * W;
  int rc = head_pack_bwd_planes_launch(std_pack, w2, C, OP, st);
  if (rc) return rc;
  BsArgs a;
  a.dlogits = dlogits; a.saved = saved; a.dfeat = dfeat; a.dw_part = dw_part; a.cls_part = cls_part;
  a.N = N; a.C = C; a.CP = CP; a.O = O; a.HW = HW;
  a.tiles_per_img = (HW + BS_BM - 1) / BS_BM;
  a.total_tiles = a.tiles_per_img * N;
  a.cb = C < 128 ? C : 128;
  a.nblk = C / a.cb;
  a.hc = make_head_consts(c);
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)N * C};
  const cuuint64_t gstride[1] = {(cuuint64_t)HW * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BS_CHUNK, (cuuint32_t)a.cb};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (features, SW128) failed (%d)", (int)cr);
    return HALO_ERR_CUDA;
  }
  const BsSmem L = bs_smem_layout(NP, OP, C);
  const float* cls = std_pack + (size_t)CPAD * 2 * OP;
  switch (OP) {
    case 4: return launch_bs<16, 4>(tmap, a, w2, cls, L.total, grid, st);
    case 8: return launch_bs<16, 8>(tmap, a, w2, cls, L.total, grid, st);
    case 12: return launch_bs<32, 12>(tmap, a, w2, cls, L.total, grid, st);
    case 16: return launch_bs<32, 16>(tmap, a, w2, cls, L.total, grid, st);
    case 20: return launch_bs<48, 20>(tmap, a, w2, cls, L.total, grid, st);
    default: return launch_bs<48, 24>(tmap, a, w2, cls, L.total, grid, st);
  }
}

}  // namespace halo
