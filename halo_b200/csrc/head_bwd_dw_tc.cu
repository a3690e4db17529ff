// K4b-TC -- weight gradient of the fused head backward on the Blackwell tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   dW^T[c][n] = sum_px u[c][px] * G[n][px]          c < C channels, n < NR = 2*OP rows of G = [gS | gT]
//
// The reduction runs over PIXELS, so here the pixel axis is the MMA K dimension and the channel is the TMEM lane:
//   D[128 ch x NP] += A[128 ch x 8 px] . B[8 px x NP]       3xTF32: main += Uhi.Ghi ; corr += Ulo.Ghi + Uhi.Glo
//   A (features)  from TMEM: converter threads (thread = channel = lane) read their channel's 32 pixels of a
//                 [128 ch x 32 px] TMA box -- 128-byte rows, SWIZZLE_128B, so the 8 x LDS.128 of a row are
//                 bank-conflict free although every lane reads a different row -- split hi/lo, tcgen05.st.
//   B (G planes)  from shared memory: [NR rows x 32 px] TMA box, SWIZZLE_128B = the canonical K-major SW128 operand
//                 layout; a splitter warpgroup rounds it to TF32 in place and writes the remainder plane beside it.
// One persistent 512-thread CTA per SM:
//   warp 0        TMA producer (features + G per 32-pixel chunk into a 4-5 stage ring)
//   warps 1, 3    MMA issuers for channel block 0 / 1 (128 channels each)
//   warps 4-7, 8-11  converter warpgroups of channel block 0 / 1
//   warps 12-15   G splitter
// Accumulators stay in TMEM across chunks.  Every tcgen05.mma rounds its accumulator toward zero (profiles/
// r1_tc_numerics.md), a bias that grows with the chain length, so every DW_DRAIN units (64 chunks = 256 MMAs on the
// main accumulator, <= 5e-6 relative) the converter threads add the accumulators into fp32 registers and the chain
// restarts.  Output: one partial [NR][CP] per CTA, reduced in a fixed order by head_bwd_finalize_kernel.
#include <cuda.h>

#include "common.cuh"
#include "head_common.cuh"
#include "head_tc.cuh"
#include "tc_common.cuh"

namespace halo {

constexpr int DW_CHUNK = 32;                 // pixels per pipeline stage (128-byte rows)
#ifndef HALO_DW_UNIT_CHUNKS
#define HALO_DW_UNIT_CHUNKS 4
#endif
constexpr int DW_UNIT_CHUNKS = HALO_DW_UNIT_CHUNKS;   // consecutive chunks per work unit (4: 512 contiguous bytes of every row)
constexpr int DW_DRAIN = 16;                 // units per accumulator chain
constexpr int DW_TC_THREADS = 512;
constexpr int DW_U_BOX_BYTES = 128 * DW_CHUNK * 4;   // 16 KB: one channel block of one chunk
#ifndef HALO_DW_NBUF
#define HALO_DW_NBUF 2
#endif
constexpr int DW_NBUF = HALO_DW_NBUF;        // A buffers per channel block, used round-robin by 16-pixel hand-over
constexpr int DW_A_COL = 0;                  // A buffers: [wg][DW_NBUF] x 32 columns (16 hi + 16 lo)
constexpr int DW_ACC_COL = 2 * DW_NBUF * 32; // accumulators: [wg] x (main NP | corr NP)

struct DwTcArgs {
  float* dw_part;   // [grid][NR][CP]
  int N, C, CP, HW, nwg, units_per_img;
  long long total_units;
  int stages, g_bytes;   // g_bytes: NP rows of 128 bytes reserved per G buffer (the MMA addresses NP rows)
};

__host__ __device__ inline size_t dw_stage_bytes(int nwg, int g_bytes) { return (size_t)nwg * DW_U_BOX_BYTES + 2 * (size_t)g_bytes; }

// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_b_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

template <int NP, int OP>
__global__ void __launch_bounds__(DW_TC_THREADS, 1)
head_bwd_dw_tc_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_g, const DwTcArgs a) {
  constexpr int NR = 2 * OP;
  static_assert(DW_ACC_COL + 2 * 2 * NP <= 512, "TMEM column budget");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nwg = a.nwg, NST = a.stages;
  const size_t stage_bytes = dw_stage_bytes(nwg, a.g_bytes);
  unsigned char* ring = smem;                                           // stage: [u block 0][u block 1][G hi][G lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NST * stage_bytes);
  uint64_t* full = bars;               // [8]  TMA bytes landed
  uint64_t* empty = bars + 8;          // [8]  converters have read the features, MMAs have read G
  uint64_t* g_ready = bars + 16;       // [8]  splitter has written G hi / lo
  uint64_t* a_full = bars + 24;                      // [2][DW_NBUF]
  uint64_t* a_empty = a_full + 2 * DW_NBUF;          // [2][DW_NBUF]
  uint64_t* acc_full = a_empty + 2 * DW_NBUF;        // [2]
  uint64_t* acc_empty = acc_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4 * nwg + nwg);   // one lane per converter warp + one tcgen05.commit per MMA issuer
      mbar_init(&g_ready[s], 4);
    }
    for (int i = 0; i < 2 * DW_NBUF; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long my_units = ((long long)blockIdx.x < a.total_units) ? (a.total_units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long my_chunks = my_units * DW_UNIT_CHUNKS;

  if (warp == 0) {
    // =================== TMA producer ===================
    {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < my_units; ++i) {
        const long long unit = blockIdx.x + i * gridDim.x;
        const int n = (int)(unit / a.units_per_img);
        const int p0 = (int)(unit - (long long)n * a.units_per_img) * (DW_CHUNK * DW_UNIT_CHUNKS);
        for (int q = 0; q < DW_UNIT_CHUNKS; ++q) {
          mbar_wait(&empty[s], ph ^ 1u);
          if (elect_one_sync()) {
            unsigned char* st = ring + (size_t)s * stage_bytes;
            mbar_arrive_expect_tx(&full[s], (uint32_t)(nwg * DW_U_BOX_BYTES + NR * 128));
            for (int g = 0; g < nwg; ++g) tma_load_2d(st + (size_t)g * DW_U_BOX_BYTES, &tmap_u, p0 + q * DW_CHUNK, n * a.C + g * 128, &full[s]);
            tma_load_2d(st + (size_t)nwg * DW_U_BOX_BYTES, &tmap_g, p0 + q * DW_CHUNK, n * NR, &full[s]);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // =================== MMA issuers (warp 1: channels 0-127, warp 3: channels 128-255) ===================
    // warp-wide loop, tcgen05 instructions on an elected lane (see elect_one_sync in tc_common.cuh)
    const int g = __shfl_sync(0xffffffffu, warp >> 1, 0);
    if (g < nwg) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t ring_u32 = __shfl_sync(0xffffffffu, smem_u32(ring), 0);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t d_main = tb + DW_ACC_COL + g * 2 * NP, d_corr = d_main + NP;
      int s = 0;
      uint32_t ph = 0;
      long long ca = 0;
      for (long long i = 0; i < my_units; ++i) {
        const bool chain_start = (i % DW_DRAIN) == 0;
        if (chain_start) {
          mbar_wait(&acc_empty[g], ((uint32_t)(i / DW_DRAIN) & 1u) ^ 1u);   // the previous chain has been drained
          tc_fence_after();
        }
        for (int q = 0; q < DW_UNIT_CHUNKS; ++q, ++ca) {
          mbar_wait(&g_ready[s], ph);
          const uint32_t g_hi = ring_u32 + (uint32_t)(s * stage_bytes) + (uint32_t)(nwg * DW_U_BOX_BYTES), g_lo = g_hi + (uint32_t)a.g_bytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const long long hc = ca * 2 + h;                    // hand-over counter -> buffer, phase
            const int bf = (int)(hc % DW_NBUF);
            mbar_wait(&a_full[g * DW_NBUF + bf], (uint32_t)(hc / DW_NBUF) & 1u);
            tc_fence_after();
            const uint32_t a_col = tb + DW_A_COL + (g * DW_NBUF + bf) * 32;
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t koff = (uint32_t)(h * 16 + ks * 8) * 4;   // byte offset of the K slice inside the 128-byte rows
                const uint64_t b_hi = make_b_desc_sw128(g_hi + koff);
                const uint64_t b_lo = make_b_desc_sw128(g_lo + koff);
                const uint32_t first = (chain_start && q == 0 && h == 0 && ks == 0) ? 0u : 1u;
                tc_mma_tf32_ts(d_main, a_col + ks * 8, b_hi, idesc, first);
                tc_mma_tf32_ts(d_corr, a_col + 16 + ks * 8, b_hi, idesc, first);
                tc_mma_tf32_ts(d_corr, a_col + ks * 8, b_lo, idesc, 1u);
              }
              tc_commit(&a_empty[g * DW_NBUF + bf]);
              if (h == 1) tc_commit(&empty[s]);
            }
            __syncwarp();
          }
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        if ((i % DW_DRAIN) == DW_DRAIN - 1 || i == my_units - 1) {
          if (elect_one_sync()) tc_commit(&acc_full[g]);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =================== converter warpgroups (thread = channel = TMEM lane) ===================
    const int g = (warp - 4) >> 2;
    if (g < nwg) {
      const int wq = warp & 3, ch = wq * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
      float acc[NR];
#pragma unroll
      for (int k = 0; k < NR; ++k) acc[k] = 0.f;
      int s = 0;
      uint32_t ph = 0;
      long long ca = 0;
      for (long long i = 0; i < my_units; ++i) {
        for (int q = 0; q < DW_UNIT_CHUNKS; ++q, ++ca) {
          mbar_wait(&full[s], ph);
          const unsigned char* row = ring + (size_t)s * stage_bytes + (size_t)g * DW_U_BOX_BYTES + ch * 128;
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + (((j ^ (ch & 7))) << 4));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const long long hc = ca * 2 + h;
            const int bf = (int)(hc % DW_NBUF);
            mbar_wait(&a_empty[g * DW_NBUF + bf], ((uint32_t)(hc / DW_NBUF) & 1u) ^ 1u);
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
            const uint32_t taddr = tmem_base + lane_addr + DW_A_COL + (g * DW_NBUF + bf) * 32;
            tmem_st_x16(taddr, hi);
            tmem_st_x16(taddr + 16, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&a_full[g * DW_NBUF + bf]);
              if (h == 1) mbar_arrive(&empty[s]);   // this warp's rows of the stage are in registers / TMEM
            }
          }
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        if ((i % DW_DRAIN) == DW_DRAIN - 1 || i == my_units - 1) {
          // drain: add the chain's accumulators (main + correction) into the fp32 registers
          mbar_wait(&acc_full[g], (uint32_t)(i / DW_DRAIN) & 1u);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + DW_ACC_COL + g * 2 * NP;
#pragma unroll
          for (int c8 = 0; c8 < NR / 8; ++c8) {
            float m8[8], c8v[8];
            tmem_ld_x8(taddr + c8 * 8, m8);
            tmem_ld_x8(taddr + NP + c8 * 8, c8v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[c8 * 8 + e] += m8[e] + c8v[e];
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[g]);
        }
      }
      float* out = a.dw_part + (size_t)blockIdx.x * NR * a.CP + g * 128 + ch;
#pragma unroll
      for (int k = 0; k < NR; ++k) out[(size_t)k * a.CP] = acc[k];
    }
  } else if (warp >= 12) {
    // =================== G splitter: TF32 hi in place, remainder plane beside it ===================
    const int t = threadIdx.x - 12 * 32;
    int s = 0;
    uint32_t ph = 0;
    for (long long ca = 0; ca < my_chunks; ++ca) {
      mbar_wait(&full[s], ph);
      float4* ghi = reinterpret_cast<float4*>(ring + (size_t)s * stage_bytes + (size_t)nwg * DW_U_BOX_BYTES);
      float4* glo = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(ghi) + a.g_bytes);
      for (int i = t; i < NR * 8; i += 128) {
        const float4 f = ghi[i];
        float4 h, l;
        h.x = __uint_as_float(cvt_rna_tf32(f.x)); h.y = __uint_as_float(cvt_rna_tf32(f.y));
        h.z = __uint_as_float(cvt_rna_tf32(f.z)); h.w = __uint_as_float(cvt_rna_tf32(f.w));
        l.x = f.x - h.x; l.y = f.y - h.y; l.z = f.z - h.z; l.w = f.w - h.w;
        ghi[i] = h;
        glo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_ready[s]);
      if (++s == NST) { s = 0; ph ^= 1u; }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------
bool head_bwd_dw_tc_supported(int C, int O, int H, int W, const void* feat, const void* G) {
  if (C != 128 && C != 256) return false;
  if (O > 32) return false;
  if (((long long)H * W) % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(feat) & 15) != 0 || (reinterpret_cast<uintptr_t>(G) & 15) != 0) return false;
  return get_encode_fn() != nullptr;
}

int head_bwd_dw_tc_grid(int N, int HW) {
  const long long units = (long long)N * ((HW + DW_CHUNK * DW_UNIT_CHUNKS - 1) / (DW_CHUNK * DW_UNIT_CHUNKS));
  long long g = sm_count();
  return (int)(g < units ? g : units);
}

template <int NP, int OP>
static int launch_dw_tc(const CUtensorMap& tu, const CUtensorMap& tg, const DwTcArgs& a, size_t smem, int grid, cudaStream_t st) {
  HALO_CUDA(cudaFuncSetAttribute(head_bwd_dw_tc_kernel<NP, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_bwd_dw_tc_kernel<NP, OP><<<grid, DW_TC_THREADS, smem, st>>>(tu, tg, a);
  return launch_status("head_bwd_dw_tc_kernel");
}

// G: [N][2*OP][HW] fp32 planes; dw_part: [grid][2*OP][CP]
int head_bwd_dw_tc_launch(const float* feat, const float* G, float* dw_part, int N, int C, int O, int H, int W, int CP, int grid,
                          cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return HALO_ERR_CUDA;
  }
  const int OP = head_op_pad(O), NR = 2 * OP, NP = round_up(NR, 16), HW = H * W;
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap tu, tg;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)N * C};
    const cuuint64_t gstride[1] = {(cuuint64_t)HW * 4};
    const cuuint32_t box[2] = {(cuuint32_t)DW_CHUNK, 128u};
    CUresult cr = enc(&tu, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (features, SW128) failed (%d)", (int)cr);
      return HALO_ERR_CUDA;
    }
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)N * NR};
    const cuuint64_t gstride[1] = {(cuuint64_t)HW * 4};
    const cuuint32_t box[2] = {(cuuint32_t)DW_CHUNK, (cuuint32_t)NR};
    CUresult cr = enc(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(G), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (G planes, SW128) failed (%d)", (int)cr);
      return HALO_ERR_CUDA;
    }
  }
  DwTcArgs a;
  a.dw_part = dw_part;
  a.N = N; a.C = C; a.CP = CP; a.HW = HW;
  a.nwg = C / 128;
  a.units_per_img = (HW + DW_CHUNK * DW_UNIT_CHUNKS - 1) / (DW_CHUNK * DW_UNIT_CHUNKS);
  a.total_units = (long long)N * a.units_per_img;
  a.g_bytes = NP * 128;
  const size_t stage = dw_stage_bytes(a.nwg, a.g_bytes);
  int stages = (int)(((size_t)227 * 1024 - 1024) / stage);
  if (stages > 8) stages = 8;
  a.stages = stages;
  const size_t smem = (size_t)stages * stage + (24 + 4 * DW_NBUF + 4 + 1) * 8 + 16;
  switch (OP) {
    case 4: return launch_dw_tc<16, 4>(tu, tg, a, smem, grid, st);
    case 8: return launch_dw_tc<16, 8>(tu, tg, a, smem, grid, st);
    case 12: return launch_dw_tc<32, 12>(tu, tg, a, smem, grid, st);
    case 16: return launch_dw_tc<32, 16>(tu, tg, a, smem, grid, st);
    case 20: return launch_dw_tc<48, 20>(tu, tg, a, smem, grid, st);
    case 24: return launch_dw_tc<48, 24>(tu, tg, a, smem, grid, st);
    case 28: return launch_dw_tc<64, 28>(tu, tg, a, smem, grid, st);
    default: return launch_dw_tc<64, 32>(tu, tg, a, smem, grid, st);
  }
}

}  // namespace halo
