// K3 -- budgeted greedy selection with square neighbourhood suppression, exact parallel form of
// select_pixels_to_label (core/active/build.py:27-64).
//
// The sequential loop picks, n_regions times, the arg-max of `score` (ties: smallest w, then smallest h --
// nested torch.max keeps the first index), stops at -inf, and sets a (2m+1)^2 window to -inf.  Picks are
// therefore made in strictly decreasing order of the composite key
//        ckey = (orderable(score) << LB) | (HW-1 - (w*H + h))
// and a pixel is picked iff no EARLIER-ordered pick lies within Chebyshev distance m.  One CTA per image:
//   1. MSB-first radix descent over ckey finds the largest key range [lo, bound] whose live population fits
//      the shared-memory candidate list (a "chunk"); massive ties descend into the index bits automatically;
//   2. the chunk is gathered and bitonic-sorted in shared memory (descending ckey);
//   3. one warp walks the sorted list 32 candidates at a time: a suppression bitmap answers "already
//      masked?", the survivors of a group are resolved in order with ballots, every pick marks its window;
//   4. chunks repeat (bound := lo-1) until n_regions picks were made or no live candidate remains;
//   5. all threads replay the pick list onto score / active / selected / active_mask (idempotent writes).
#include "common.cuh"

namespace halo {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr size_t SEL_LIST_BYTES = 64 * 1024;
constexpr size_t SEL_SMEM_MAX = 220 * 1024;

template <typename S>
struct SelTraits;
template <>
struct SelTraits<float> {
  typedef unsigned long long Comp;
  static constexpr int KB = 32;
  __device__ static __forceinline__ Comp key(float x) {
    x = x + 0.0f;  // -0.0 -> +0.0 (torch.max treats them as equal)
    unsigned b = __float_as_uint(x);
    if (x != x) return 0xffffffffull;  // NaN wins every max (torch.max propagates NaN)
    return (Comp)((b & 0x80000000u) ? ~b : (b | 0x80000000u));
  }
  __device__ static __forceinline__ bool dead(float x) { return x == -INFINITY; }
};
template <>
struct SelTraits<double> {
  typedef unsigned __int128 Comp;
  static constexpr int KB = 64;
  __device__ static __forceinline__ Comp key(double x) {
    x = x + 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    if (x != x) return (Comp)0xffffffffffffffffull;
    return (Comp)((b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull));
  }
  __device__ static __forceinline__ bool dead(double x) { return x == -(double)INFINITY; }
};

template <typename S>
struct SelArgs {
  S* score;
  uint8_t* active;
  uint8_t* selected;
  uint8_t* active_mask;
  const uint8_t* gt;
  int* n_picked;
  int* picks;          // [N][n_regions]
  unsigned* gbitmap;   // [N][words] or NULL when the bitmap lives in shared memory
  int n_regions, a_r, m_r, H, W, LB, cap, words;
};

// Visit every pixel of one image plane as f(p, value).  16-byte vector loads, 4 independent loads in flight per
// thread (the scans are pure streaming passes; without the batching they are latency-bound at ~1 load/thread).
template <typename S, typename F>
__device__ __forceinline__ void scan_plane(const S* __restrict__ plane, int HW, int tid, F f) {
  constexpr int V = 16 / (int)sizeof(S);
  constexpr int UN = 4;
  struct __align__(16) Vec { S e[V]; };
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(plane) & 15) == 0) {
    const int nvec = HW / V;
    const Vec* vp = reinterpret_cast<const Vec*>(plane);
    for (int base = tid; base < nvec; base += SEL_THREADS * UN) {
      Vec v[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int idx = base + j * SEL_THREADS;
        if (idx < nvec) v[j] = vp[idx];
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int idx = base + j * SEL_THREADS;
        if (idx < nvec) {
#pragma unroll
          for (int e = 0; e < V; ++e) f(idx * V + e, v[j].e[e]);
        }
      }
    }
    done = nvec * V;
  }
  for (int p = done + tid; p < HW; p += SEL_THREADS) f(p, plane[p]);
}

struct SelShared {
  unsigned hist[SEL_BINS];
  unsigned warp_tot[32];
  int sel_bin;
  unsigned cum;
  unsigned total;
  unsigned cnt;
  int npicks;
  int grp_p[32];  // pixel index of the picks of the current group, in pick order
};

// address-space-specific accessors of the suppression bitmap: shared-memory atomics (ATOMS) when it fits next to the
// candidate list, global atomics otherwise -- a generic pointer would make every mark a slow generic ATOM
template <bool GBM>
struct Bitmap {
  unsigned* base;  // shared (GBM=false) or global (GBM=true)
  __device__ __forceinline__ bool test(int p) const {
    if (GBM) return (__ldcg(base + (p >> 5)) >> (p & 31)) & 1u;
    return (*(volatile unsigned*)(base + (p >> 5)) >> (p & 31)) & 1u;
  }
  __device__ __forceinline__ void set_bits(int word, unsigned msk) const {
    if (GBM) atomicOr(base + word, msk);
    else atomicOr(reinterpret_cast<unsigned*>(__cvta_shared_to_generic(__cvta_generic_to_shared(base + word))), msk);
  }
};

template <typename S, bool GBM>
__global__ void __launch_bounds__(SEL_THREADS, 1) select_kernel(const SelArgs<S> a) {
  typedef SelTraits<S> TR;
  typedef typename TR::Comp Comp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Comp* list = reinterpret_cast<Comp*>(smem_raw);
  unsigned* sbitmap = reinterpret_cast<unsigned*>(smem_raw + SEL_LIST_BYTES);
  __shared__ SelShared sh;

  const int img = blockIdx.x;
  const int H = a.H, W = a.W, HW = H * W, LB = a.LB, TB = TR::KB + LB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  S* score = a.score + (size_t)img * HW;
  unsigned* bitmap = GBM ? a.gbitmap + (size_t)img * a.words : sbitmap;
  const Bitmap<GBM> bm{bitmap};
  int* picks = a.picks + (size_t)img * a.n_regions;
  const Comp lin_mask = (((Comp)1) << LB) - 1;
  const unsigned LINMAX = (unsigned)(HW - 1);

  for (int i = tid; i < a.words; i += SEL_THREADS) bitmap[i] = 0u;
  for (int i = tid; i < a.n_regions; i += SEL_THREADS) picks[i] = -1;
  if (tid == 0) sh.npicks = 0;
  __syncthreads();

  Comp bound = ~(Comp)0;
  if (TB < (int)sizeof(Comp) * 8) bound = ((((Comp)1) << TB) - 1);
  int npicks = 0;
  bool more = (a.n_regions > 0);

  while (more) {
    // ---------------- 1. radix descent: choose lo so that |{live, lo <= ckey <= bound}| <= cap ----------------
    Comp prefix = 0, lo = 0;
    unsigned accepted = 0;
    int level = 0;
    bool exhausted = false;  // nothing alive at or below `bound`
    bool covers_all = false; // chunk reaches down to ckey 0: no later chunk can exist
    while (true) {
      const int hi_bit = TB - SEL_BITS * level;            // bits [shift, hi_bit) form this level's digit
      const int shift = hi_bit > SEL_BITS ? hi_bit - SEL_BITS : 0;
      const int nb = hi_bit - shift;
      const unsigned dmask = (1u << nb) - 1u;
      for (int i = tid; i < SEL_BINS; i += SEL_THREADS) sh.hist[i] = 0u;
      __syncthreads();
      scan_plane<S>(score, HW, tid, [&](int p, S v) {
        if (TR::dead(v)) return;
        if ((bitmap[p >> 5] >> (p & 31)) & 1u) return;
        const int h = p / W, w = p - h * W;
        const Comp ck = (TR::key(v) << LB) | (Comp)(LINMAX - (unsigned)(w * H + h));
        if (ck > bound) return;
        if (level > 0 && (ck >> hi_bit) != prefix) return;
        atomicAdd(&sh.hist[(unsigned)(ck >> shift) & dmask], 1u);
      });
      __syncthreads();
      // suffix scan over bins (top bin first): thread t owns reversed indices 2t, 2t+1
      const int b_hi = SEL_BINS - 1 - 2 * tid, b_lo = b_hi - 1;
      const unsigned c_hi = sh.hist[b_hi], c_lo = sh.hist[b_lo];
      unsigned incl = c_hi + c_lo;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) sh.warp_tot[warp] = incl;
      if (tid == 0) sh.sel_bin = -1;
      __syncthreads();
      unsigned base = 0;
      for (int wv = 0; wv < warp; ++wv) base += sh.warp_tot[wv];
      const unsigned before = accepted + base + incl - (c_hi + c_lo);  // population strictly above bin b_hi
      // crossing bin: first bin (from the top) at which the running total would exceed cap
      if (before <= (unsigned)a.cap && before + c_hi > (unsigned)a.cap) { sh.sel_bin = b_hi; sh.cum = before; }
      else if (before + c_hi <= (unsigned)a.cap && before + c_hi + c_lo > (unsigned)a.cap) { sh.sel_bin = b_lo; sh.cum = before + c_hi; }
      if (tid == SEL_THREADS - 1) sh.total = accepted + base + incl;
      __syncthreads();
      const int sel_bin = sh.sel_bin;
      const unsigned total = sh.total;
      if (sel_bin < 0) {  // everything under this prefix fits
        accepted = total;
        lo = (level > 0) ? (prefix << hi_bit) : (Comp)0;
        covers_all = (level == 0);
        exhausted = (total == 0);
        break;
      }
      const unsigned cum = sh.cum;
      if (cum > 0 && (cum >= (unsigned)(a.cap - a.cap / 8) || shift == 0)) {  // fill the list to >= 7/8 before accepting
        accepted = cum;
        lo = (((prefix << nb) + (Comp)(sel_bin + 1)) << shift);
        break;
      }
      accepted = cum;
      prefix = (prefix << nb) | (Comp)sel_bin;
      ++level;
      __syncthreads();
    }
    if (exhausted) break;

    // ---------------- 2. gather + sort (descending ckey) ----------------
    if (tid == 0) sh.cnt = 0u;
    __syncthreads();
    scan_plane<S>(score, HW, tid, [&](int p, S v) {
      if (TR::dead(v)) return;
      if ((bitmap[p >> 5] >> (p & 31)) & 1u) return;
      const int h = p / W, w = p - h * W;
      const Comp ck = (TR::key(v) << LB) | (Comp)(LINMAX - (unsigned)(w * H + h));
      if (ck > bound || ck < lo) return;
      const unsigned pos = atomicAdd(&sh.cnt, 1u);
      if (pos < (unsigned)a.cap) list[pos] = ck;
    });
    __syncthreads();
    const int cnt = (int)min(sh.cnt, (unsigned)a.cap);
    int n2 = 32;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + tid; i < n2; i += SEL_THREADS) list[i] = (Comp)0;
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2; i += SEL_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const Comp x = list[i], y = list[ixj];
            const bool desc = ((i & kk) == 0);
            if (desc ? (x < y) : (x > y)) { list[i] = y; list[ixj] = x; }
          }
        }
        __syncthreads();
      }
    }

    // ---------------- 3. ordered greedy walk (warp 0) ----------------
    if (warp == 0) {
      const int m = a.m_r;
      if (m == 0) {
        // pixel mode: nothing suppresses anything else -> the sorted prefix IS the pick sequence
        const int take = min(cnt, a.n_regions - npicks);
        for (int i = lane; i < take; i += 32) {
          const unsigned lin = LINMAX - (unsigned)(list[i] & lin_mask);
          const int w = (int)(lin / (unsigned)H), h = (int)(lin - (unsigned)w * H);
          const int p = h * W + w;
          picks[npicks + i] = p;
          bm.set_bits(p >> 5, 1u << (p & 31));
        }
        npicks += take;
      } else {
        for (int base = 0; base < cnt && npicks < a.n_regions; base += 32) {
          // lane i holds the i-th best remaining candidate of this group
          const int i = base + lane;
          int h = -1000000, w = -1000000;
          bool alive = false;
          if (i < cnt) {
            const unsigned lin = LINMAX - (unsigned)(list[i] & lin_mask);
            w = (int)(lin / (unsigned)H);
            h = (int)(lin - (unsigned)w * H);
            alive = !bm.test(h * W + w);           // not inside the window of a pick of an earlier group
          }
          const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
          if (alive_mask == 0u) continue;
          // conflict mask: earlier lanes of the group within Chebyshev distance m
          unsigned conf = 0u;
#pragma unroll 8
          for (int j = 0; j < 31; ++j) {
            const int hj = __shfl_sync(0xffffffffu, h, j), wj = __shfl_sync(0xffffffffu, w, j);
            if (j < lane && abs(h - hj) <= m && abs(w - wj) <= m) conf |= 1u << j;
          }
          // resolve in order (identical in every lane): picked iff alive and no earlier PICKED lane conflicts
          unsigned picked = 0u;
          unsigned todo = alive_mask;
          while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned cj = __shfl_sync(0xffffffffu, conf, j);
            if ((cj & picked) == 0u) picked |= 1u << j;
          }
          // budget: keep only the first (n_regions - npicks) picks of the group
          const int room = a.n_regions - npicks;
          if (__popc(picked) > room) {
            unsigned keep = 0u, t = picked;
            for (int r = 0; r < room; ++r) { keep |= t & (0u - t); t &= t - 1; }
            picked = keep;
          }
          const int npk = __popc(picked);
          if ((picked >> lane) & 1u) {
            const int rank = __popc(picked & ((1u << lane) - 1u));
            picks[npicks + rank] = h * W + w;
            sh.grp_p[rank] = h * W + w;
          }
          __syncwarp();
          // mark the (2m+1)^2 windows of all picks of the group: (pick, row) tasks spread over the lanes
          const int rows = 2 * m + 1;
          for (int t = lane; t < npk * rows; t += 32) {
            const int k = t / rows, ry = t - k * rows;
            const int pj = sh.grp_p[k];
            const int hj = pj / W, wj = pj - hj * W;
            const int y = hj - m + ry;
            if (y < 0 || y >= H) continue;
            const int x0 = max(wj - m, 0), x1 = min(wj + m, W - 1);
            const int p0 = y * W + x0, p1 = y * W + x1;
            for (int wd = p0 >> 5; wd <= (p1 >> 5); ++wd) {
              const int blo = max(p0 - (wd << 5), 0), bhi = min(p1 - (wd << 5), 31);
              const unsigned msk = (bhi == 31 ? 0xffffffffu : ((1u << (bhi + 1)) - 1u)) & ~((1u << blo) - 1u);
              bm.set_bits(wd, msk);
            }
          }
          npicks += npk;
          __syncwarp();
        }
      }
      if (lane == 0) sh.npicks = npicks;
    }
    __syncthreads();
    npicks = sh.npicks;
    if (npicks >= a.n_regions || covers_all || lo == (Comp)0) more = false;
    else bound = lo - 1;
    __syncthreads();
  }

  // ---------------- 5. replay the picks onto the four planes (build.py:45-62) ----------------
  __syncthreads();
  npicks = sh.npicks;
  if (tid == 0) a.n_picked[img] = npicks;
  {
    const int m = a.m_r, ar = a.a_r;
    const S ninf = (S)(-INFINITY);
    uint8_t* act = a.active + (size_t)img * HW;
    uint8_t* sel = a.selected + (size_t)img * HW;
    uint8_t* msk = a.active_mask + (size_t)img * HW;
    const uint8_t* gt = a.gt + (size_t)img * HW;
    const int rows_m = 2 * m + 1, rows_a = 2 * ar + 1;
    const long long work_m = (long long)npicks * rows_m;
    for (long long t = tid; t < work_m; t += SEL_THREADS) {
      const int pi = (int)(t / rows_m), ry = (int)(t - (long long)pi * rows_m);
      const int p = picks[pi];
      const int h = p / W, w = p - h * W;
      const int y = h - m + ry;
      if (y < 0 || y >= H) continue;
      const int x0 = max(w - m, 0), x1 = min(w + m, W - 1);
      for (int x = x0; x <= x1; ++x) {
        score[y * W + x] = ninf;
        act[y * W + x] = 1;
      }
    }
    const long long work_a = (long long)npicks * rows_a;
    for (long long t = tid; t < work_a; t += SEL_THREADS) {
      const int pi = (int)(t / rows_a), ry = (int)(t - (long long)pi * rows_a);
      const int p = picks[pi];
      const int h = p / W, w = p - h * W;
      const int y = h - ar + ry;
      if (y < 0 || y >= H) continue;
      const int x0 = max(w - ar, 0), x1 = min(w + ar, W - 1);
      for (int x = x0; x <= x1; ++x) {
        sel[y * W + x] = 1;
        msk[y * W + x] = gt[y * W + x];
      }
    }
  }
}

static int sel_lb(int HW) {
  int lb = 1;
  while ((1LL << lb) < (long long)HW) ++lb;
  return lb;
}

template <typename S>
static int select_launch(S* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask, const uint8_t* gt,
                         int n_regions, int active_radius, int mask_radius, int* n_picked, int* picks, int N, int H,
                         int W, void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(score && active && selected && active_mask && gt && n_picked, "halo_select: NULL pointer");
  HALO_CHECK_ARG(N > 0 && H > 0 && W > 0, "halo_select: bad dims");
  HALO_CHECK_ARG(n_regions >= 0 && active_radius >= 0 && mask_radius >= 0, "halo_select: negative budget / radius");
  HALO_CHECK_ARG((long long)H * W < (1LL << 28), "halo_select: image too large");
  const size_t need = halo_select_workspace_bytes(N, H, W, n_regions);
  if (!ws || ws_bytes < need) {
    set_error("halo_select: workspace %zu < %zu bytes", ws_bytes, need);
    return HALO_ERR_WORKSPACE;
  }
  const int HW = H * W;
  SelArgs<S> a;
  a.score = score; a.active = active; a.selected = selected; a.active_mask = active_mask; a.gt = gt;
  a.n_picked = n_picked;
  a.n_regions = n_regions; a.a_r = active_radius; a.m_r = mask_radius; a.H = H; a.W = W; a.LB = sel_lb(HW);
  a.cap = (int)(SEL_LIST_BYTES / sizeof(typename SelTraits<S>::Comp));
  a.words = (HW + 31) / 32;
  const size_t picks_bytes = (((size_t)N * (n_regions > 0 ? n_regions : 1) * sizeof(int)) + 255) / 256 * 256;
  a.picks = picks ? picks : (int*)ws;
  const size_t bitmap_bytes = (size_t)a.words * 4;
  size_t smem = SEL_LIST_BYTES;
  if (SEL_LIST_BYTES + bitmap_bytes <= SEL_SMEM_MAX) {
    a.gbitmap = nullptr;
    smem += bitmap_bytes;
  } else {
    a.gbitmap = (unsigned*)((unsigned char*)ws + picks_bytes);
  }
  if (a.gbitmap) {
    HALO_CUDA(cudaFuncSetAttribute(select_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_kernel<S, true><<<N, SEL_THREADS, smem, (cudaStream_t)stream>>>(a);
  } else {
    HALO_CUDA(cudaFuncSetAttribute(select_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_kernel<S, false><<<N, SEL_THREADS, smem, (cudaStream_t)stream>>>(a);
  }
  return launch_status("select_kernel");
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_select_workspace_bytes(int N, int H, int W, int n_regions) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const size_t picks_bytes = (((size_t)N * (n_regions > 0 ? n_regions : 1) * sizeof(int)) + 255) / 256 * 256;
  const size_t words = ((size_t)H * W + 31) / 32;
  return picks_bytes + (size_t)N * words * 4 + 256;
}

extern "C" int halo_select_f32(float* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask,
                               const uint8_t* gt, int n_regions, int active_radius, int mask_radius, int* n_picked,
                               int* picks, int N, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream) {
  return select_launch<float>(score, active, selected, active_mask, gt, n_regions, active_radius, mask_radius,
                              n_picked, picks, N, H, W, ws, ws_bytes, stream);
}

extern "C" int halo_select_f64(double* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask,
                               const uint8_t* gt, int n_regions, int active_radius, int mask_radius, int* n_picked,
                               int* picks, int N, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream) {
  return select_launch<double>(score, active, selected, active_mask, gt, n_regions, active_radius, mask_radius,
                               n_picked, picks, N, H, W, ws, ws_bytes, stream);
}
