// K1-TC -- fused Poincare-ball head, forward, Blackwell tensor-core path (tcgen05 + TMEM + TMA), sm_100a.
//
// Why: the contraction [pixels x C] . [C x 2O] costs 2*C*(2O+1) ~ 20 kFLOP per 1 KB of features; the fp32 FMA
// pipe tops out at ~40-59 % of the HBM roofline (profiles/r1_k1_cuda_core.md), so the contraction moves to the
// 5th-gen tensor cores.  Plain TF32 (10-bit mantissa) would break the 1e-5 parity bound, so each operand is
// split u = hi + lo (both TF32-representable, round-to-nearest) and three MMAs accumulate in fp32:
//        D += U_hi.W_hi + U_lo.W_hi + U_hi.W_lo            (the dropped lo.lo term is ~2^-22 relative)
//
// Pipeline of one persistent CTA (one per SM, 512 threads):
//   warps 0, 2        TMA producers (warp-wide loop, TMA on an elected lane; one per pixel warpgroup): [32 channels x 128 pixels] fp32 boxes of
//                     the NCHW feature tensor (viewed as a 2-D [N*C, H*W] tensor) into that warpgroup's 3-deep
//                     shared-memory ring (16 KB per stage), completion by mbarrier tx-count.
//   warps 4-7, 8-11   two CONVERTER warpgroups working on alternate (adjacent) 128-pixel tiles; thread = pixel = TMEM
//                     lane: read the pixel's channel values of a stage (conflict-free, pixels contiguous), accumulate
//                     |u|^2, split hi/lo, tcgen05.st both into the A-operand columns of TMEM.
//   warps 12-15       EPILOGUE warpgroup, every tile in order: tcgen05.ld the accumulator columns of the pixel and run
//                     the same register epilogue as the CUDA-core kernel (Mobius algebra, asinh, radius, softmax
//                     entropy); the converters never wait for it (profiles/r1_k1_tc.md: a warpgroup that converts AND
//                     finishes its tile leaves the SM issue slots half empty).
//   warps 1, 3        MMA issuers (warp-wide loop, tcgen05 on an elected lane; one per pixel warpgroup): tcgen05.mma.kind::tf32, M=128 (pixels) x
//                     N=NP (2*OP padded to 16) x K=8,
//                     A from TMEM, B (class parameters, hi and lo planes) from shared memory, D in TMEM;
//                     tcgen05.commit releases A buffers / publishes accumulators through mbarriers.
// The class parameters (<= 96 KB) stay resident in shared memory for the life of the CTA.
#include <cuda.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"
#include "tc_common.cuh"

namespace halo {

constexpr int TC_BM = 128;      // pixels per tile = MMA M = TMEM lanes
constexpr int TC_BK = 32;       // channels per pipeline stage
#ifndef HALO_TC_WG_STAGES
#define HALO_TC_WG_STAGES 3
#endif
constexpr int TC_WG_STAGES = HALO_TC_WG_STAGES;  // shared-memory stages per pixel warpgroup (each warpgroup has its own ring + producer)
constexpr int TC_STAGES = TC_WG_STAGES * 2;
constexpr int TC_STAGE_FLOATS = TC_BK * TC_BM;
constexpr int TC_NWG = 2;       // converter warpgroups (each owns a TMA ring, an MMA issuer, A buffers and accumulators)
constexpr int TC_THREADS = 128 + 128 * TC_NWG + 128;  // control warps + converters + one epilogue warpgroup
// A second epilogue warpgroup (NEPI = 2: 640 threads, epilogue warpgroup e finishes the tiles of converter warpgroup e) is
// used for C <= 128: the epilogue costs the same per pixel whatever C is, so with a quarter / half of the conversion work
// per pixel the single epilogue warpgroup paces the kernel (C = 64, the channel count of every shipped HALO config: 44 % of
// the HBM roofline with one, see profiles/r2_k1.md).  At C = 256 the SM's total issue bandwidth is the limit and one is as good.
constexpr int tc_threads(int nepi) { return 128 + 128 * TC_NWG + 128 * nepi; }
constexpr int TC_TMEM_COLS = 512;
#ifndef HALO_TC_HK
#define HALO_TC_HK 16
#endif
// channels per A-operand buffer (one converter -> MMA hand-over).  8 (four buffers per warpgroup, finer hand-overs) was
// measured slower than 16 (two buffers): 4 080 vs 4 250 GB/s sustained -- the barrier traffic costs more than the depth wins.
constexpr int TC_HK = HALO_TC_HK;
constexpr int TC_NHAND = TC_BK / TC_HK;             // hand-overs per pipeline stage
#ifndef HALO_TC_NBUF
#define HALO_TC_NBUF (TC_BK / TC_HK)
#endif
constexpr int TC_NBUF = HALO_TC_NBUF;               // A buffers per warpgroup, used round-robin by hand-over
constexpr int TC_ACOLS = 2 * TC_HK;                 // per A buffer: TC_HK hi + TC_HK lo columns
constexpr int TC_ACC_COL0 = TC_NWG * TC_NBUF * TC_ACOLS;  // = 128: accumulators start after the A buffers
constexpr int TC_NACC_MAX = 4;  // partial accumulators per tile (shortens the in-TMEM accumulation chains)

// i-th tile of CTA b (G CTAs): the NWG warpgroups of a CTA take ADJACENT 128-pixel tiles (2b, 2b+1, then +2G ...), so a
// CTA reads 1 KB runs of every channel row within a short window and neighbouring CTAs continue the same DRAM pages.
__device__ __forceinline__ int tc_tile_of(int i, int b, int G) { return (i / TC_NWG) * (TC_NWG * G) + b * TC_NWG + (i % TC_NWG); }

struct TcSmemLayout {
  size_t w_bytes, ring_off, bar_off, tmem_off, cls_off, n2_off, total;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int NP, int OP, int C) {
  TcSmemLayout L;
  (void)NP;
  L.w_bytes = (size_t)2 * (2 * OP) * C * 4;  // only the 2*OP real rows are stored (see head_pack_tc_kernel)
  L.ring_off = (L.w_bytes + 1023) / 1024 * 1024;
  L.bar_off = L.ring_off + (size_t)TC_STAGES * TC_STAGE_FLOATS * 4;
  const int nbars = 2 * TC_STAGES + TC_NWG * TC_NBUF * 2 + TC_NWG * 2;
  L.tmem_off = L.bar_off + (size_t)nbars * 8;
  L.cls_off = (L.tmem_off + 16 + 15) / 16 * 16;
  L.n2_off = L.cls_off + (size_t)4 * OP * 4;
  L.total = L.n2_off + (size_t)TC_NWG * TC_BM * 4;  // |u|^2 of the tile each converter warpgroup just finished
  return L;
}

// Accumulator plan (measured on B200, tools/tc_numerics.py, profiles/r1_tc_numerics.md): every tcgen05.mma rounds
// its fp32 accumulator toward zero, a systematic ~2^-25.5 relative shrink per instruction, so the error of a dot
// product grows LINEARLY with the number of MMAs chained on one accumulator.  Therefore
//   * the dominant hi.hi products go to NMAIN "main" accumulators used round-robin by pipeline stage
//     (C=256: 8 stages x 4 MMAs over 3 accumulators -> chains of <= 12 MMAs instead of 96), and
//   * the two cross terms lo.hi + hi.lo (2^-11 of the main term, so their rounding is harmless) share one
//     "correction" accumulator;
// the epilogue adds the NMAIN+1 partial results in fp32 registers.
template <int NP, int OP, int NMAIN, int NEPI>
__global__ void __launch_bounds__(tc_threads(NEPI), 1)
head_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const HeadArgs a, const float* __restrict__ wtc) {
  constexpr int NACC = NMAIN + 1;  // accumulator NMAIN is the correction accumulator
  static_assert(TC_ACC_COL0 + TC_NWG * NACC * NP <= TC_TMEM_COLS, "TMEM column budget");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int C = a.C;
  const TcSmemLayout L = tc_smem_layout(NP, OP, C);
  constexpr int NR = 2 * OP;                   // stored rows of the B operand; the MMA's rows NR..NP-1 alias the next K slice
  float* sW = reinterpret_cast<float*>(smem);  // [2][C/4][NR][4]: hi plane then lo plane
  float* ring = reinterpret_cast<float*>(smem + L.ring_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_STAGES;
  uint64_t* a_full = bars + 2 * TC_STAGES;               // [NWG][2]
  uint64_t* a_empty = a_full + TC_NWG * TC_NBUF;         // [NWG][NBUF]
  uint64_t* acc_full = a_empty + TC_NWG * TC_NBUF;       // [NWG]
  uint64_t* acc_empty = acc_full + TC_NWG;               // [NWG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
  float* sCls = reinterpret_cast<float*>(smem + L.cls_off);
  float* sN2 = reinterpret_cast<float*>(smem + L.n2_off);  // [NWG][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ----
  {
    const int n4 = 2 * NR * C / 4;
    const float4* src = reinterpret_cast<const float4*>(wtc);
    float4* dst = reinterpret_cast<float4*>(sW);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    // [4][OP] -> per class PAIR two float4 {pp0, pp1, an0, an1} {pa0, pa1, Bk0, Bk1} (operands of the packed epilogue)
    const float* csrc = wtc + (size_t)2 * NR * C;
    for (int i = threadIdx.x; i < 4 * OP; i += blockDim.x) {
      const int q = i / OP, k = i % OP;
      sCls[(k >> 1) * 8 + q * 2 + (k & 1)] = csrc[i];
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);     // one elected lane per converter warp
    }
    for (int i = 0; i < TC_NWG * TC_NBUF; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);   // tcgen05.commit
    }
    for (int g = 0; g < TC_NWG; ++g) {
      mbar_init(&acc_full[g], 1);  // tcgen05.commit
      mbar_init(&acc_empty[g], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy weight stores -> visible to the MMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int HW = a.HW;
  const int cpt = C / TC_BK;  // chunks (pipeline stages) per tile
  // number of i with tc_tile_of(i) < total_tiles: full rounds of NWG tiles plus the tail of the last round
  int my_tiles = 0;
  {
    const int per_round = TC_NWG * gridDim.x;
    const int rounds = a.total_tiles / per_round, rem = a.total_tiles - rounds * per_round;
    const int tail = rem - (int)blockIdx.x * TC_NWG;
    my_tiles = rounds * TC_NWG + (tail <= 0 ? 0 : (tail >= TC_NWG ? TC_NWG : tail));
  }

  // NEPI = 2: 640 threads start with 96 registers each; the roles then trade them (setmaxnreg is warpgroup-wide and sits
  // at the top of each role's region, no control-flow merge behind it): control 40, converters 80, epilogue 136
  // (128 x 40 + 256 x 80 + 256 x 136 = 60 416 <= 640 x 96).
  if (warp < 4) {
  if (NEPI == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 || warp == 2) {
    // =================== TMA producers: warp 0 feeds warpgroup 0, warp 2 feeds warpgroup 1 ===================
    // Each warpgroup owns a private ring (stages + barriers): a barrier then has exactly one producer and one
    // consumer advancing in lock-step, which the 1-bit mbarrier phase parity requires.
    // (warp-wide loop, the TMA instruction on an elected lane: see elect_one_sync in tc_common.cuh)
    const int g = __shfl_sync(0xffffffffu, warp >> 1, 0);
    {
      for (int i = g; i < my_tiles; i += TC_NWG) {
        const int it = i / TC_NWG;
        const int tile = tc_tile_of(i, blockIdx.x, gridDim.x);
        const int n = tile / a.tiles_per_img;
        const int p0 = (tile - n * a.tiles_per_img) * TC_BM;
        for (int j = 0; j < cpt; ++j) {
          const int q = it * cpt + j;                       // stage counter of this warpgroup
          const int s = g * TC_WG_STAGES + q % TC_WG_STAGES;
          const uint32_t ph = (uint32_t)(q / TC_WG_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full[s], TC_STAGE_FLOATS * 4);
            tma_load_2d(ring + (size_t)s * TC_STAGE_FLOATS, &tmap, p0, n * C + j * TC_BK, &full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // =================== MMA issuers: warp 1 serves warpgroup 0, warp 3 serves warpgroup 1 ===================
    // One issuing thread per pixel warpgroup, each strictly in order for ITS warpgroup: the two conversion
    // streams then overlap on the tensor pipe instead of being serialised behind a single in-order issuer
    // (tcgen05.commit tracks the MMAs of the executing thread only, and the two streams touch disjoint TMEM).
    // The loop runs warp-wide (converged) and only the tcgen05 instructions are predicated on an elected lane; warp
    // index and TMEM base go through a shuffle so that ptxas sees every MMA operand as warp-uniform.
    const int g = __shfl_sync(0xffffffffu, warp >> 1, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N=NP, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t w_hi = __shfl_sync(0xffffffffu, smem_u32(sW), 0), w_lo = w_hi + (uint32_t)NR * C * 4;
      const uint32_t lbo = NR * 16, sbo = 128;
      const uint32_t d_corr = tb + TC_ACC_COL0 + (g * NACC + NMAIN) * NP;
      for (int i = g; i < my_tiles; i += TC_NWG) {
        const int it = i / TC_NWG;
        mbar_wait(&acc_empty[g], ((uint32_t)it & 1u) ^ 1u);
        tc_fence_after();
        for (int j = 0; j < cpt; ++j) {
          const int ca = it * cpt + j;                      // stage counter of this warpgroup
          const uint32_t d_main = tb + TC_ACC_COL0 + (g * NACC + (j % NMAIN)) * NP;
#pragma unroll
          for (int h = 0; h < TC_NHAND; ++h) {
            const int hc = ca * TC_NHAND + h, bf = hc % TC_NBUF;   // hand-over counter -> buffer, phase
            mbar_wait(&a_full[g * TC_NBUF + bf], (uint32_t)(hc / TC_NBUF) & 1u);
            tc_fence_after();
            const uint32_t a_col = tb + (g * TC_NBUF + bf) * TC_ACOLS;
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < TC_HK / 8; ++ks) {
                const uint32_t koff = (uint32_t)((j * TC_BK + h * TC_HK + ks * 8) / 4) * lbo;  // byte offset of the K-slice
                const uint64_t b_hi = make_b_desc(w_hi + koff, lbo, sbo);
                const uint64_t b_lo = make_b_desc(w_lo + koff, lbo, sbo);
                const bool first_k = (h == 0 && ks == 0);
                tc_mma_tf32_ts(d_main, a_col + ks * 8, b_hi, idesc, (j < NMAIN && first_k) ? 0u : 1u);   // hi . hi
                tc_mma_tf32_ts(d_corr, a_col + TC_HK + ks * 8, b_hi, idesc, (j == 0 && first_k) ? 0u : 1u);  // lo . hi
                tc_mma_tf32_ts(d_corr, a_col + ks * 8, b_lo, idesc, 1u);                                 // hi . lo
              }
              tc_commit(&a_empty[g * TC_NBUF + bf]);
            }
            __syncwarp();
          }
        }
        if (elect_one_sync()) tc_commit(&acc_full[g]);
        __syncwarp();
      }
    }
  }
  } else if (warp >= 4 && warp < 4 + 4 * TC_NWG) {
    // =================== converter warpgroups (thread = pixel = TMEM lane) ===================
    if (NEPI == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    const int g = (warp - 4) >> 2;
    const int wq = warp & 3;                   // TMEM lane quarter this warp may touch
    const int m = wq * 32 + lane;              // pixel row inside the tile
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    for (int i = g; i < my_tiles; i += TC_NWG) {
      const int it = i / TC_NWG;
      unsigned long long n2 = 0ull;  // two fp32 partial sums of |u|^2
      for (int j = 0; j < cpt; ++j) {
        const int ca = it * cpt + j;                        // stage counter of this warpgroup
        const int s = g * TC_WG_STAGES + ca % TC_WG_STAGES;
        mbar_wait(&full[s], (uint32_t)(ca / TC_WG_STAGES) & 1u);
        const float* src = ring + (size_t)s * TC_STAGE_FLOATS + m;
#pragma unroll
        for (int h = 0; h < TC_NHAND; ++h) {
          const int hc = ca * TC_NHAND + h, bf = hc % TC_NBUF;
          mbar_wait(&a_empty[g * TC_NBUF + bf], ((uint32_t)(hc / TC_NBUF) & 1u) ^ 1u);
          tc_fence_after();
          uint32_t hi[TC_HK], lo[TC_HK];
          tc_split<TC_HK>(src + h * TC_HK * TC_BM, TC_BM, hi, lo, n2);
          const uint32_t taddr = tmem_base + lane_addr + (g * TC_NBUF + bf) * TC_ACOLS;
          tmem_st(taddr, hi);
          tmem_st(taddr + TC_HK, lo);
          // |u|^2 travels to the epilogue warpgroup through shared memory; it is published by the release of the
          // tile's last a_full arrival (-> MMA issuer -> tcgen05.commit -> acc_full acquire in the epilogue)
          if (j == cpt - 1 && h == TC_NHAND - 1) sN2[g * TC_BM + m] = n2_of(n2);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[g * TC_NBUF + bf]);
        }
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
  } else if (warp >= 4 + 4 * TC_NWG) {
    // =================== epilogue warpgroup(s): NEPI = 1 every tile in order; NEPI = 2 warpgroup e the tiles of converter
    // warpgroup e (accumulators, |u|^2 and their barriers are per converter warpgroup already) ===================
    if (NEPI == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int epi = (warp - (4 + 4 * TC_NWG)) >> 2;
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const HeadConsts hc = a.hc;
    static_assert(NEPI == 1 || NEPI == TC_NWG, "one epilogue warpgroup, or one per converter warpgroup");
    for (int i = (NEPI == 1 ? 0 : epi); i < my_tiles; i += NEPI) {
      const int g = i % TC_NWG, it = i / TC_NWG;
      const int tile = tc_tile_of(i, blockIdx.x, gridDim.x);
      const int n = tile / a.tiles_per_img;
      const int p = (tile - n * a.tiles_per_img) * TC_BM + m;
      mbar_wait(&acc_full[g], (uint32_t)it & 1u);
      tc_fence_after();
      const float n2 = sN2[g * TC_BM + m];
      // the contractions as class PAIRS (float2 = one 64-bit register pair = one operand of FFMA2 / FADD2 / FMUL2)
      float2 S2[OP / 2], T2[OP / 2];
#pragma unroll
      for (int j = 0; j < OP / 2; ++j) S2[j] = T2[j] = make_float2(0.f, 0.f);
      static_assert((2 * OP) % 8 == 0, "OP is a multiple of 4");
#pragma unroll
      for (int qa = 0; qa < NACC; ++qa) {
        if (qa < cpt || qa == NMAIN) {  // main accumulators that received no stage (C < NMAIN*32) hold garbage
          const uint32_t taddr = tmem_base + lane_addr + TC_ACC_COL0 + (g * NACC + qa) * NP;
          float buf[2 * OP];
#pragma unroll
          for (int c8 = 0; c8 < (2 * OP) / 8; ++c8) tmem_ld_x8(taddr + c8 * 8, *reinterpret_cast<float(*)[8]>(&buf[c8 * 8]));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < OP / 2; ++j) {
            S2[j] = __fadd2_rn(S2[j], make_float2(buf[2 * j], buf[2 * j + 1]));
            T2[j] = __fadd2_rn(T2[j], make_float2(buf[OP + 2 * j], buf[OP + 2 * j + 1]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[g]);

      const bool live = (p < HW);
      if (a.saved != nullptr && live) {
        // training forward: keep the contractions for the streaming backward (head_bwd_stream_tc.cu), which then reads
        // the features ONCE (164 B per pixel saved here against a second 4*C-byte pass over u there)
        float* sv = a.saved + (size_t)n * (2 * OP + 1) * HW + p;
#pragma unroll
        for (int j = 0; j < OP / 2; ++j) {
          if (2 * j < a.O) {
            __stcs(sv + (size_t)(2 * j) * HW, S2[j].x);
            __stcs(sv + (size_t)(OP + 2 * j) * HW, T2[j].x);
          }
          if (2 * j + 1 < a.O) {
            __stcs(sv + (size_t)(2 * j + 1) * HW, S2[j].y);
            __stcs(sv + (size_t)(OP + 2 * j + 1) * HW, T2[j].y);
          }
        }
        __stcs(sv + (size_t)(2 * OP) * HW, n2);
      }
      if (a.logits == nullptr && a.radius == nullptr && a.pixunc == nullptr && a.label == nullptr && a.stats == nullptr)
        continue;   // contraction-only launch (backward called without saved planes)
      const PixelScalars ps = tangent_scalars(n2, hc);
      float2 l2[OP / 2];
#pragma unroll
      for (int j = 0; j < OP / 2; ++j) {
        l2[j] = make_float2(-3.0e38f, -3.0e38f);   // padded classes: skipped (warp-uniform), invisible to the softmax below
        if (2 * j < OP - 3 || 2 * j < a.O) {
          const float4 c0 = reinterpret_cast<const float4*>(sCls)[2 * j], c1 = reinterpret_cast<const float4*>(sCls)[2 * j + 1];
          l2[j] = mlr_logit2(S2[j], T2[j], ps, make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y),
                             make_float2(c1.z, c1.w), hc);
          if (2 * j + 1 >= a.O) l2[j].y = -3.0e38f;   // the odd half of the last pair when O is odd
        }
      }
      float l[OP];
#pragma unroll
      for (int j = 0; j < OP / 2; ++j) { l[2 * j] = l2[j].x; l[2 * j + 1] = l2[j].y; }
#ifdef HALO_TC_VARIANTS
      if (a.debug_raw) {  // numerics probe: expose the raw contractions <u,a_hat_k> instead of the logits
#pragma unroll
        for (int j = 0; j < OP / 2; ++j) { l[2 * j] = T2[j].x; l[2 * j + 1] = T2[j].y; }
      }
#endif
      const float r = (a.norm_mode == HALO_NORM_EUCLID) ? ps.xnorm : ps.radius;
      const size_t pix = (size_t)n * HW + p;
      if (live) {
        if (a.logits != nullptr) {
#pragma unroll
          for (int k = 0; k < OP; ++k)
            if (k < a.O) __stcs(a.logits + ((size_t)n * a.O + k) * HW + p, l[k]);
        }
        if (a.radius != nullptr) __stcs(a.radius + pix, r);
      }
      if (a.pixunc != nullptr || a.label != nullptr) {
        int gtv = 255;
        if (a.gt != nullptr && live) gtv = a.gt[pix];
        float unc;
        int lab = 0;
        if (a.label == nullptr && a.pixunc_mode == HALO_PIXUNC_ENTROPY) unc = softmax_entropy_only2<OP>(l2, hc);
        else softmax_stats<OP>(l, a.O, hc, a.pixunc_mode, a.label_mode, gtv, unc, lab);
        if (live) {
          if (a.pixunc != nullptr) __stcs(a.pixunc + pix, unc);
          if (a.label != nullptr) a.label[pix] = (uint8_t)lab;
        }
      }
      if (a.stats != nullptr) {
        const float rmin = warp_min(live ? r : __int_as_float(0x7f800000));
        const float rmax = warp_max(live ? r : 0.f);
        if (lane == 0) {
          atomicMin(reinterpret_cast<int*>(a.stats + 4 * n + 0), __float_as_int(rmin));
          atomicMax(reinterpret_cast<int*>(a.stats + 4 * n + 1), __float_as_int(rmax));
        }
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// ---- class parameters in the tensor-core operand layout ---------------------------------------------------
// in : std pack  Wt[CPAD][2*OP] + cls[4][OP]     (head_pack_kernel)
// out: W planes  [2 (hi,lo)][C/4][NR][4] fp32 with TF32-representable values (round-to-nearest split), NR = 2*OP,
//      followed by cls[4][OP].  Row n of the B operand: n < OP -> -P_n ; OP <= n < 2*OP -> a_hat_{n-OP}.
//      The MMA runs with N = NP (NR rounded up to 16): its rows NR..NP-1 fall on the first rows of the next K slice
//      (finite values) and only feed accumulator columns that nobody reads -- a column of D depends on its own row
//      of B alone -- so the padding rows are not stored (16 KB of shared memory at C=256, O=19).
__global__ void head_pack_tc_kernel(const float* __restrict__ std_pack, float* __restrict__ wtc, int C, int CPAD, int OP) {
  const int NR = 2 * OP;
  const int total = NR * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k4 = i / (NR * 4);
    const int rem = i - k4 * NR * 4;
    const int nrow = rem >> 2, kk = rem & 3;
    const int ch = k4 * 4 + kk;
    const float w = std_pack[(size_t)ch * NR + nrow];
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(w));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(w - __uint_as_float(h)));
    wtc[i] = __uint_as_float(h);
    wtc[(size_t)total + i] = __uint_as_float(l);
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < 4 * OP; i += blockDim.x) wtc[(size_t)2 * total + i] = std_pack[(size_t)CPAD * NR + i];
  }
}

// ---- host side ----------------------------------------------------------------------------------------
int head_tc_np(int O) {
  const int OP = head_op_pad(O);
  return round_up(2 * OP, 16);
}

bool head_tc_supported(int feat_kind, int C, int O, int H, int W, const void* feat) {
  return head_tc_shape_ok(feat_kind, C, O, H, W, feat) && get_encode_fn() != nullptr;
}

bool head_tc_shape_ok(int feat_kind, int C, int O, int H, int W, const void* feat) {
  if (feat_kind != HALO_FEAT_TANGENT_F32) return false;
  if (C % TC_BK != 0 || C > 256 || C < TC_BK) return false;
  if (O > 32) return false;
  if (((long long)H * W) % 4 != 0) return false;          // TMA global stride must be a multiple of 16 bytes
  if ((reinterpret_cast<uintptr_t>(feat) & 15) != 0) return false;
  const TcSmemLayout L = tc_smem_layout(head_tc_np(O), head_op_pad(O), C);
  return L.total <= 225 * 1024;
}

int head_pack_tc_launch(const float* std_pack, float* wtc, int C, int CPAD, int O, cudaStream_t st) {
  const int OP = head_op_pad(O);
  head_pack_tc_kernel<<<(2 * OP * C + 255) / 256, 256, 0, st>>>(std_pack, wtc, C, CPAD, OP);
  return launch_status("head_pack_tc_kernel");
}

size_t head_tc_pack_floats(int O, int C) { return (size_t)2 * head_tc_np(O) * C + 4 * head_op_pad(O); }

template <int NP, int OP>
static int launch_tc(const CUtensorMap& tmap, const HeadArgs& a, const float* wtc, size_t smem, int grid, cudaStream_t st) {
  // as many main accumulators as TMEM allows next to the A buffers (at most one per pipeline stage of C=256)
  constexpr int FIT = (TC_TMEM_COLS - TC_ACC_COL0) / (TC_NWG * NP) - 1;
  constexpr int NMAIN = FIT > 8 ? 8 : FIT;
  static_assert(NMAIN >= 1, "TMEM budget");
  static const int force = [] { const char* e = getenv("HALO_TC_NEPI"); return e ? atoi(e) : 0; }();   // A/B knob: 1 | 2
  const bool two = force ? (force == 2) : (a.C <= 128);
  if (two) {
    HALO_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<NP, OP, NMAIN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_fwd_tc_kernel<NP, OP, NMAIN, 2><<<grid, tc_threads(2), smem, st>>>(tmap, a, wtc);
  } else {
    HALO_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<NP, OP, NMAIN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_fwd_tc_kernel<NP, OP, NMAIN, 1><<<grid, tc_threads(1), smem, st>>>(tmap, a, wtc);
  }
  return launch_status("head_fwd_tc_kernel");
}

int head_fwd_tc_launch(HeadArgs a, const float* std_pack, float* wtc, cudaStream_t st) {
  const int OP = head_op_pad(a.O), NP = head_tc_np(a.O);
  head_pack_tc_kernel<<<(2 * OP * a.C + 255) / 256, 256, 0, st>>>(std_pack, wtc, a.C, a.CPAD, OP);
  int rc = launch_status("head_pack_tc_kernel");
  if (rc) return rc;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return HALO_ERR_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.HW, (cuuint64_t)a.N * a.C};
  const cuuint64_t gstride[1] = {(cuuint64_t)a.HW * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BM, (cuuint32_t)TC_BK};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(a.feat), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return HALO_ERR_CUDA;
  }
  a.tiles_per_img = (a.HW + TC_BM - 1) / TC_BM;
  a.total_tiles = a.tiles_per_img * a.N;
  const TcSmemLayout L = tc_smem_layout(NP, OP, a.C);
  int grid = sm_count();
  if (grid > (a.total_tiles + TC_NWG - 1) / TC_NWG) grid = (a.total_tiles + TC_NWG - 1) / TC_NWG;
  switch (OP) {
    case 4: return launch_tc<16, 4>(tmap, a, wtc, L.total, grid, st);
    case 8: return launch_tc<16, 8>(tmap, a, wtc, L.total, grid, st);
    case 12: return launch_tc<32, 12>(tmap, a, wtc, L.total, grid, st);
    case 16: return launch_tc<32, 16>(tmap, a, wtc, L.total, grid, st);
    case 20: return launch_tc<48, 20>(tmap, a, wtc, L.total, grid, st);
    case 24: return launch_tc<48, 24>(tmap, a, wtc, L.total, grid, st);
    case 28: return launch_tc<64, 28>(tmap, a, wtc, L.total, grid, st);
    default: return launch_tc<64, 32>(tmap, a, wtc, L.total, grid, st);
  }
}

}  // namespace halo
