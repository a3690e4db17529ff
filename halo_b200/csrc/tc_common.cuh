// PTX wrappers shared by the tcgen05 kernels (forward head_fwd_tc.cu, backward head_bwd_tc.cu): mbarrier, TMA,
// tcgen05 fences / commit / MMA / TMEM load-store, shared-memory matrix descriptors, tensor-map encoder lookup.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace halo {

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Blocking wait on a phase parity.  The suspend-time hint lets the hardware park the thread until the phase flips
// instead of re-issuing the probe every few tens of nanoseconds: in the r1e profile 39 % of all executed
// instructions were TRYWAIT/BRA pairs of the producer / issuer / waiting pixel warps (issue slots and power).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// same, with an L2 cache policy (createpolicy): the backward reads every feature tile a second time a few
// microseconds later (alpha*u in the output warps), so its lines are asked to stay (evict_last) until then
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One elected lane of a fully converged warp (the CUTLASS idiom).  The MMA-issuer warps run their loops warp-wide and
// predicate only the tcgen05 instructions on this: under `if (lane == 0) { loop }` ptxas cannot prove the operands
// uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~13 instructions per MMA on one
// lane: the issuer, not the tensor pipe, paced the converters -- profiles/r1_k1.md r1n).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred;
}
// D[tmem] (+)= A[tmem] . B[smem]     kind::tf32, cta_group::1
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// round-to-nearest (ties away) to TF32 with two integer ops; ptxas expands cvt.rna.tf32.f32 into ~6 instructions
// on sm_100a (profiles/r1_k1.md).  Sign-magnitude bits: adding half an ulp of the 10-bit mantissa to the
// magnitude and clearing the 13 low bits rounds correctly for both signs (inf/NaN inputs stay non-finite).
__device__ __forceinline__ uint32_t cvt_rna_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// 16 feature values of one pixel (channel stride `stride` floats) -> TF32 hi (round-to-nearest) and the exact
// remainder lo = u - hi, plus |u|^2 accumulated in two fp32 lanes.  The remainder is NOT re-rounded: it has <= 13
// significant bits and the tensor core reads the top 10 of them (absolute error <= 2^-21 |u|, unbiased in sign) --
// measured equal to the rounded variant in tools/tc_numerics.py.  The two fp32 subtractions / multiply-adds of a
// channel pair issue as one packed FADD2 / FFMA2 (sm_100a), so a pair costs 2 LDS + 2 IADD + 2 LOP + FADD2 + FFMA2
// instead of 14 instructions.
__device__ __forceinline__ unsigned long long pack_f32x2(uint32_t a, uint32_t b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
template <int NK>
__device__ __forceinline__ void tc_split(const float* __restrict__ src, int stride, uint32_t (&hi)[NK], uint32_t (&lo)[NK],
                                         unsigned long long& n2acc) {
#pragma unroll
  for (int k = 0; k < NK; k += 2) {
    const uint32_t u0 = __float_as_uint(src[k * stride]), u1 = __float_as_uint(src[(k + 1) * stride]);
    const unsigned long long uu = pack_f32x2(u0, u1);
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(n2acc) : "l"(uu));
    const uint32_t h0 = (u0 + 0x1000u) & 0xffffe000u, h1 = (u1 + 0x1000u) & 0xffffe000u;
    unsigned long long ll;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(ll) : "l"(uu), "l"(pack_f32x2(h0, h1)));
    hi[k] = h0;
    hi[k + 1] = h1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo[k]), "=r"(lo[k + 1]) : "l"(ll));
  }
}
__device__ __forceinline__ void tc_split16(const float* __restrict__ src, int stride, uint32_t (&hi)[16], uint32_t (&lo)[16],
                                           unsigned long long& n2acc) {
  tc_split<16>(src, stride, hi, lo, n2acc);
}
__device__ __forceinline__ float n2_of(unsigned long long n2acc) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(n2acc));
  return __uint_as_float(a) + __uint_as_float(b);
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[8]) { tmem_st_x8(taddr, v); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st_x16(taddr, v); }
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle ("interleaved" canonical layout, see head_tc.cuh)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}


// x32 TMEM load (32 consecutive columns of this thread's lane)
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}


}  // namespace halo
