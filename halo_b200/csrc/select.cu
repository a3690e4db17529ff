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
constexpr int SEL_TEAM_WARPS = 4;  // warps that walk the sorted candidate list together
constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr size_t SEL_LIST_BYTES = 64 * 1024;  // candidate list: 8 192 fp32-score keys (4 096 fp64-score keys); power of two (bitonic)
constexpr size_t SEL_SMEM_MAX = 212 * 1024;  // dynamic part: 227 KB opt-in limit minus the static SelShared (~9.1 KB) and slack

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
  int n_regions, a_r, m_r, H, W, LB, cap, words, keep_score;
  unsigned long long* prof;  // HALO_SEL_PROFILE builds: cycles per phase, summed over CTAs
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
  int grp_h[32], grp_w[32];  // coordinates of the picks of the current group, in pick order
  unsigned conf_part[SEL_TEAM_WARPS][32];
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

// fast exact p -> (h, w) without an integer division
__device__ __forceinline__ void split_hw(int p, int W, float invW, int& h, int& w) {
  if (p >= (1 << 24)) {  // float(p) would not be exact
    h = p / W;
    w = p - h * W;
    return;
  }
  h = __float2int_rd(((float)p + 0.5f) * invW);
  w = p - h * W;
  if (w < 0) { --h; w += W; }
  else if (w >= W) { ++h; w -= W; }
}

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
  const float invW = 1.0f / (float)W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  S* score = a.score + (size_t)img * HW;
  unsigned* bitmap = GBM ? a.gbitmap + (size_t)img * a.words : sbitmap;
  const Bitmap<GBM> bm{bitmap};
  int* picks = a.picks + (size_t)img * a.n_regions;
  const Comp lin_mask = (((Comp)1) << LB) - 1;
  const unsigned LINMAX = (unsigned)(HW - 1);
  // index part of the composite key of pixel p (larger = earlier in the reference's tie-break: smallest w, then h)
  auto lin_of = [&](int p) -> Comp {
    int h, w;
    split_hw(p, W, invW, h, w);
    return (Comp)(LINMAX - (unsigned)(w * H + h));
  };

#ifdef HALO_SEL_PROFILE
  long long t_prev = clock64();
#define SEL_PHASE(k) do { if (tid == 0) { const long long t_now = clock64(); atomicAdd(a.prof + (k), (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
#else
#define SEL_PHASE(k) do { } while (0)
#endif
  for (int i = tid; i < a.words; i += SEL_THREADS) bitmap[i] = 0u;
  for (int i = tid; i < a.n_regions; i += SEL_THREADS) picks[i] = -1;
  if (tid == 0) sh.npicks = 0;
  __syncthreads();

  Comp bound = ~(Comp)0;
  if (TB < (int)sizeof(Comp) * 8) bound = ((((Comp)1) << TB) - 1);
  int npicks = 0;
  bool more = (a.n_regions > 0);
  // The first chunk estimates its threshold on a row sample (every SAMPLE-th row): ~1/16 of a pass instead of 2-3 full
  // histogram passes.  The gather that follows is exact whatever the estimate; only an overflow of the candidate list
  // (sampling error, or heavy ties) sends the chunk back through the exact full-image descent.
  constexpr int SAMPLE = 16;
  const bool can_sample = (H >= 4 * SAMPLE) && (HW >= 64 * a.cap);
  bool sampled = can_sample;

  while (more) {
    // ---------------- 1. radix descent: choose lo so that |{live, lo <= ckey <= bound}| <= cap ----------------
    const Comp bound_s = bound >> LB, bound_l = bound & lin_mask;
    const unsigned cap_eff = sampled ? (unsigned)(a.cap - a.cap / 8) / SAMPLE : (unsigned)a.cap;
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
      auto visit = [&](int p, S v) {
        if (TR::dead(v)) return;
        if ((bitmap[p >> 5] >> (p & 31)) & 1u) return;
        const Comp sk = TR::key(v);
        if (sk > bound_s) return;
        unsigned digit;
        if (shift >= LB && sk < bound_s) {          // common case: the index bits are not needed
          if (level > 0 && (sk >> (hi_bit - LB)) != prefix) return;
          digit = (unsigned)(sk >> (shift - LB)) & dmask;
        } else {
          const Comp ck = (sk << LB) | lin_of(p);
          if (ck > bound) return;
          if (level > 0 && (ck >> hi_bit) != prefix) return;
          digit = (unsigned)(ck >> shift) & dmask;
        }
        atomicAdd(&sh.hist[digit], 1u);
      };
      if (sampled) {
        for (int r = SAMPLE / 2; r < H; r += SAMPLE)
          for (int x = tid; x < W; x += SEL_THREADS) visit(r * W + x, score[r * W + x]);
      } else {
        scan_plane<S>(score, HW, tid, visit);
      }
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
      // crossing bin: first bin (from the top) at which the running total would exceed the capacity
      if (before <= cap_eff && before + c_hi > cap_eff) { sh.sel_bin = b_hi; sh.cum = before; }
      else if (before + c_hi <= cap_eff && before + c_hi + c_lo > cap_eff) { sh.sel_bin = b_lo; sh.cum = before + c_hi; }
      if (tid == SEL_THREADS - 1) sh.total = accepted + base + incl;
      __syncthreads();
      const int sel_bin = sh.sel_bin;
      const unsigned total = sh.total;
      if (sel_bin < 0) {  // everything under this prefix fits
        accepted = total;
        lo = (level > 0) ? (prefix << hi_bit) : (Comp)0;
        covers_all = (level == 0) && !sampled;
        exhausted = (total == 0) && !sampled;
        break;
      }
      const unsigned cum = sh.cum;
      if (cum > 0 && (cum >= cap_eff - cap_eff / 8 || shift == 0)) {  // fill the list to >= 7/8 before accepting
        accepted = cum;
        lo = (((prefix << nb) + (Comp)(sel_bin + 1)) << shift);
        break;
      }
      accepted = cum;
      prefix = (prefix << nb) | (Comp)sel_bin;
      ++level;
      __syncthreads();
    }
    SEL_PHASE(0);
    if (exhausted) break;

    // ---------------- 2. gather + sort (descending ckey) ----------------
    if (tid == 0) sh.cnt = 0u;
    __syncthreads();
    {
      const Comp lo_s = lo >> LB;
      scan_plane<S>(score, HW, tid, [&](int p, S v) {
        if (TR::dead(v)) return;
        const Comp sk = TR::key(v);
        if (sk < lo_s || sk > bound_s) return;            // the vast majority of pixels leave here
        if ((bitmap[p >> 5] >> (p & 31)) & 1u) return;
        const Comp ck = (sk << LB) | lin_of(p);
        if (ck > bound || ck < lo) return;
        const unsigned pos = atomicAdd(&sh.cnt, 1u);       // (a warp-aggregated atomic was measured: +66 % gather time -- the
        if (pos < (unsigned)a.cap) list[pos] = ck;         //  scan is bound by instructions per pixel, not by this counter)
      });
    }
    __syncthreads();
    SEL_PHASE(1);
    if (sh.cnt > (unsigned)a.cap) {
      // only a sampled threshold can overflow the list: redo this chunk with the exact descent
      __syncthreads();
      sampled = false;
      continue;
    }
    const bool was_sampled = sampled;
    sampled = can_sample;  // the next chunk (if any) estimates its threshold on the sample again
    const int cnt = (int)sh.cnt;
    int n2 = 32;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + tid; i < n2; i += SEL_THREADS) list[i] = (Comp)0;
    __syncthreads();
    // bitonic sort, descending; one compare-exchange PAIR per loop trip (pair t -> element i with a zero at bit j)
    for (int kk = 2; kk <= n2; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (n2 >> 1); t += SEL_THREADS) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int ixj = i | j;
          const Comp x = list[i], y = list[ixj];
          const bool desc = ((i & kk) == 0);
          if (desc ? (x < y) : (x > y)) { list[i] = y; list[ixj] = x; }
        }
        __syncthreads();
      }
    }

    SEL_PHASE(2);
    // ---------------- 3. ordered greedy walk (a team of 4 warps in lock-step) ----------------
    // Every team warp walks the same sorted list (identical registers), so decisions need no broadcast; the
    // work that parallelises -- the pairwise conflict test and the marking of the suppression windows -- is split
    // over the 128 team threads, with named-barrier hand-offs (bar 1) between the phases of a group.
    if (warp < SEL_TEAM_WARPS) {
      const int m = a.m_r;
      const int tl = tid;  // 0 .. 32*SEL_TEAM_WARPS-1
      auto team_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * SEL_TEAM_WARPS) : "memory"); };
      if (m == 0) {
        // pixel mode: nothing suppresses anything else -> the sorted prefix IS the pick sequence
        const int take = min(cnt, a.n_regions - npicks);
        for (int i = tl; i < take; i += 32 * SEL_TEAM_WARPS) {
          const unsigned lin = LINMAX - (unsigned)(list[i] & lin_mask);
          const int w = (int)(lin / (unsigned)H), h = (int)(lin - (unsigned)w * H);
          const int p = h * W + w;
          picks[npicks + i] = p;
          bm.set_bits(p >> 5, 1u << (p & 31));
        }
        npicks += take;
      } else {
        const int rows = 2 * m + 1;
        for (int base = 0; base < cnt && npicks < a.n_regions; base += 32) {
          // lane i holds the i-th best remaining candidate of this group
          const int i = base + lane;
          int h = -30000, w = -30000;
          bool alive = false;
          if (i < cnt) {
            const unsigned lin = LINMAX - (unsigned)(list[i] & lin_mask);
            w = (int)(lin / (unsigned)H);
            h = (int)(lin - (unsigned)w * H);
            alive = !bm.test(h * W + w);           // not inside the window of a pick of an earlier group
          }
          const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
          if (alive_mask == 0u) continue;          // uniform over the team (same data in every warp)
          // conflict mask: earlier ALIVE lanes of the group within Chebyshev distance m; warp q tests lanes j = q mod 4
          const int packed = alive ? ((h << 16) | (w & 0xffff)) : (int)0x80008000;
          unsigned conf = 0u;
#pragma unroll
          for (int jj = 0; jj < 32 / SEL_TEAM_WARPS; ++jj) {
            const int j = jj * SEL_TEAM_WARPS + warp;
            const int pj = __shfl_sync(0xffffffffu, packed, j);
            const int hj = pj >> 16, wj = (int)(short)(pj & 0xffff);
            if (j < lane && abs(h - hj) <= m && abs(w - wj) <= m) conf |= 1u << j;
          }
          sh.conf_part[warp][lane] = conf;
          team_sync();
          conf = 0u;
#pragma unroll
          for (int q = 0; q < SEL_TEAM_WARPS; ++q) conf |= sh.conf_part[q][lane];
          unsigned picked = alive_mask;
          if (__any_sync(0xffffffffu, alive && conf != 0u)) {
            // resolve in order (identical in every lane): picked iff alive and no earlier PICKED lane conflicts
            picked = 0u;
            unsigned todo = alive_mask;
            while (todo) {
              const int j = __ffs(todo) - 1;
              todo &= todo - 1;
              const unsigned cj = __shfl_sync(0xffffffffu, conf, j);
              if ((cj & picked) == 0u) picked |= 1u << j;
            }
          }
          // budget: keep only the first (n_regions - npicks) picks of the group
          const int room = a.n_regions - npicks;
          if (__popc(picked) > room) {
            unsigned keep = 0u, t = picked;
            for (int r = 0; r < room; ++r) { keep |= t & (0u - t); t &= t - 1; }
            picked = keep;
          }
          const int npk = __popc(picked);
          if (warp == 0 && ((picked >> lane) & 1u)) {
            const int rank = __popc(picked & ((1u << lane) - 1u));
            picks[npicks + rank] = h * W + w;
            sh.grp_h[rank] = h;
            sh.grp_w[rank] = w;
          }
          team_sync();
          // mark the (2m+1)^2 windows of all picks of the group: (pick, row) tasks spread over the team
          for (int t = tl; t < npk * rows; t += 32 * SEL_TEAM_WARPS) {
            const int k = t / rows, ry = t - k * rows;
            const int hj = sh.grp_h[k], wj = sh.grp_w[k];
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
          team_sync();                             // marks visible to the next group's bitmap test
        }
      }
      if (tid == 0) sh.npicks = npicks;
    }
    __syncthreads();
    SEL_PHASE(3);
    npicks = sh.npicks;
    if (npicks >= a.n_regions || (covers_all && !was_sampled) || lo == (Comp)0) more = false;
    else bound = lo - 1;
    __syncthreads();
  }

  // ---------------- 5. replay the picks onto the four planes (build.py:45-62) ----------------
  __syncthreads();
  SEL_PHASE(4);
  npicks = sh.npicks;
  if (tid == 0) a.n_picked[img] = npicks;
  {
    const int ar = a.a_r;
    const S ninf = (S)(-INFINITY);
    uint8_t* act = a.active + (size_t)img * HW;
    uint8_t* sel = a.selected + (size_t)img * HW;
    uint8_t* msk = a.active_mask + (size_t)img * HW;
    const uint8_t* gt = a.gt + (size_t)img * HW;
    // score / active windows: the suppression bitmap IS the union of the (2m+1)^2 windows of all picks, so the
    // masked planes are written by one coalesced streaming pass over the bitmap instead of scattered window stores
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(score) & (4 * sizeof(S) - 1)) == 0) &&
                     ((reinterpret_cast<uintptr_t>(act) & 3) == 0);
    if (a.keep_score && (HW % 32 == 0) && ((reinterpret_cast<uintptr_t>(act) & 15) == 0)) {
      // acquisition path (the score plane is scratch): one bitmap word = 32 pixels = one 32-byte sector of `active`,
      // written whole (read-merge-write only for partly covered words) instead of byte stores at every window edge
      for (int wi = tid; wi < a.words; wi += SEL_THREADS) {
        const unsigned word = GBM ? __ldcg(bitmap + wi) : bitmap[wi];
        if (word == 0u) continue;
        uint4* ap = reinterpret_cast<uint4*>(act + (size_t)wi * 32);
        if (word == 0xffffffffu) {
          const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
          ap[0] = ones;
          ap[1] = ones;
        } else {
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const unsigned hw16 = (word >> (16 * hq)) & 0xffffu;
            if (hw16 == 0u) continue;
            uint4 o = ap[hq];
            unsigned* ow = reinterpret_cast<unsigned*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const unsigned m1 = (((hw16 >> (4 * e)) & 0xfu) * 0x00204081u) & 0x01010101u;   // 4 bits -> 4 bytes of 0 / 1
              ow[e] = (ow[e] & ~(m1 * 0xffu)) | m1;
            }
            ap[hq] = o;
          }
        }
      }
    } else if (vec) {
      struct __align__(4 * sizeof(S)) S4 { S v[4]; };
      S4 ninf4;
      ninf4.v[0] = ninf4.v[1] = ninf4.v[2] = ninf4.v[3] = ninf;
      for (int q = tid; q < HW / 4; q += SEL_THREADS) {
        const int p0 = q << 2;
        const unsigned word = GBM ? __ldcg(bitmap + (p0 >> 5)) : bitmap[p0 >> 5];
        const unsigned bits = (word >> (p0 & 31)) & 0xfu;
        if (bits == 0xfu) {
          if (!a.keep_score) *reinterpret_cast<S4*>(score + p0) = ninf4;
          *reinterpret_cast<unsigned*>(act + p0) = 0x01010101u;
        } else if (bits) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if ((bits >> e) & 1u) {
              if (!a.keep_score) score[p0 + e] = ninf;
              act[p0 + e] = 1;
            }
        }
      }
    } else {
      for (int p0 = tid; p0 < HW; p0 += SEL_THREADS) {
        const unsigned word = GBM ? __ldcg(bitmap + (p0 >> 5)) : bitmap[p0 >> 5];
        if ((word >> (p0 & 31)) & 1u) {
          if (!a.keep_score) score[p0] = ninf;
          act[p0] = 1;
        }
      }
    }
    // selected / active_mask windows (radius a_r, a few pixels per pick): direct
    const int rows_a = 2 * ar + 1;
    const long long work_a = (long long)npicks * rows_a;
    for (long long t = tid; t < work_a; t += SEL_THREADS) {
      const int pi = (int)(t / rows_a), ry = (int)(t - (long long)pi * rows_a);
      const int p = picks[pi];
      int h, w;
      split_hw(p, W, invW, h, w);
      const int y = h - ar + ry;
      if (y < 0 || y >= H) continue;
      const int x0 = max(w - ar, 0), x1 = min(w + ar, W - 1);
      for (int x = x0; x <= x1; ++x) {
        sel[y * W + x] = 1;
        msk[y * W + x] = gt[y * W + x];
      }
    }
  }
  __syncthreads();
  SEL_PHASE(5);
}

static int sel_lb(int HW) {
  int lb = 1;
  while ((1LL << lb) < (long long)HW) ++lb;
  return lb;
}

template <typename S>
static int select_launch(S* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask, const uint8_t* gt,
                         int n_regions, int active_radius, int mask_radius, int flags, int* n_picked, int* picks, int N,
                         int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream) {
  HALO_CHECK_ARG(score && active && selected && active_mask && gt && n_picked, "halo_select: NULL pointer");
  HALO_CHECK_ARG(N > 0 && H > 0 && W > 0, "halo_select: bad dims");
  HALO_CHECK_ARG(n_regions >= 0 && active_radius >= 0 && mask_radius >= 0, "halo_select: negative budget / radius");
  HALO_CHECK_ARG((long long)H * W < (1LL << 28) && H < 32768 && W < 32768, "halo_select: image too large");
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
  a.keep_score = (flags & HALO_SELECT_KEEP_SCORE) ? 1 : 0;
  a.cap = (int)(SEL_LIST_BYTES / sizeof(typename SelTraits<S>::Comp));
  a.words = (HW + 31) / 32;
  a.prof = (unsigned long long*)((unsigned char*)ws + need - 128);  // 6 counters in the slack at the end of ws
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
  return (picks_bytes + (size_t)N * words * 4 + 255) / 256 * 256 + 256;
}

extern "C" int halo_select_f32(float* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask,
                               const uint8_t* gt, int n_regions, int active_radius, int mask_radius, int flags,
                               int* n_picked, int* picks, int N, int H, int W, void* ws, size_t ws_bytes,
                               halo_stream_t stream) {
  return select_launch<float>(score, active, selected, active_mask, gt, n_regions, active_radius, mask_radius, flags,
                              n_picked, picks, N, H, W, ws, ws_bytes, stream);
}

extern "C" int halo_select_f64(double* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask,
                               const uint8_t* gt, int n_regions, int active_radius, int mask_radius, int flags,
                               int* n_picked, int* picks, int N, int H, int W, void* ws, size_t ws_bytes,
                               halo_stream_t stream) {
  return select_launch<double>(score, active, selected, active_mask, gt, n_regions, active_radius, mask_radius, flags,
                               n_picked, picks, N, H, W, ws, ws_bytes, stream);
}
