// Round deltas -- compact form of what one selection round wrote into the label masks (the only exchange
// between the image shards is the final all-gather of pick counts and masks).  A round labels the (2a+1)^2 windows
// around its picks (build.py:58-62), at most 5 % of the pixels, so the shards exchange
//     picks [N][cap] i32 (h*W+w, -1 padded)  +  lab [N][cap][(2a+1)^2] u8 (gt of the window, 255 outside the image)
// = 13 B per pick for 3x3 windows (59 KB per 1280x640 image at 4 552 picks) instead of the 819 KB mask plane, and every
// rank applies the gathered deltas to its replica of the pool's masks.
#include "common.cuh"

namespace halo {

__global__ void round_delta_pack_kernel(const int* __restrict__ picks, const int* __restrict__ n_picked,
                                        const uint8_t* __restrict__ gt, uint8_t* __restrict__ lab, int N, int cap, int H, int W,
                                        int r) {
  const int k = 2 * r + 1, k2 = k * k;
  const long long total = (long long)N * cap * k2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(t % k2);
    const long long pi = t / k2;
    const int i = (int)(pi % cap), n = (int)(pi / cap);
    uint8_t v = 255;
    const int p = (i < n_picked[n]) ? picks[pi] : -1;
    if (p >= 0) {
      const int h = p / W + e / k - r, w = p % W + e % k - r;
      if (h >= 0 && h < H && w >= 0 && w < W) v = gt[((size_t)n * H + h) * W + w];
    }
    lab[t] = v;
  }
}

__global__ void round_delta_apply_kernel(uint8_t* __restrict__ masks, const int* __restrict__ row_image,
                                         const int* __restrict__ picks, const int* __restrict__ n_picked,
                                         const uint8_t* __restrict__ lab, int rows, int cap, int H, int W, int r) {
  const int k = 2 * r + 1, k2 = k * k;
  const long long total = (long long)rows * cap * k2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(t % k2);
    const long long pi = t / k2;
    const int i = (int)(pi % cap), j = (int)(pi / cap);
    const int img = row_image[j];
    if (img < 0 || i >= n_picked[j]) continue;
    const int p = picks[pi];
    if (p < 0) continue;
    const int h = p / W + e / k - r, w = p % W + e % k - r;
    // active_mask[window] = gt[window]; an unlabeled gt pixel (255) leaves the replica as it is (it holds 255 there
    // in every consistent state), so a padded / clipped entry can never erase an earlier round's label
    const uint8_t v = lab[t];
    if (v != 255 && h >= 0 && h < H && w >= 0 && w < W) masks[((size_t)img * H + h) * W + w] = v;
  }
}

// ---- packed rows: ONE buffer per shard, so the round ends with ONE all-gather --------------------------------------
// Row of pool image j (row_bytes = halo_round_row_bytes(cap, a), a multiple of 16):
//     [ int32 count ][ int32 picks[cap] ][ uint8 lab[cap][(2a+1)^2] ][ pad ]
// pack writes count, the first `count` picks and their window labels (entries past `count` are left as they are: apply
// never reads them); apply replays rows onto the replicated masks and scatters the counts to n_picked_out[image].
__host__ __device__ inline size_t round_row_bytes(int cap, int r) {
  const size_t k2 = (size_t)(2 * r + 1) * (2 * r + 1);
  return (4 + (size_t)cap * 4 + (size_t)cap * k2 + 15) / 16 * 16;
}

__global__ void round_rows_pack_kernel(const int* __restrict__ picks, const int* __restrict__ n_picked,
                                       const uint8_t* __restrict__ gt, uint8_t* __restrict__ rows, int N, int cap,
                                       int pick_stride, int H, int W, int r, size_t row_bytes) {
  const int k = 2 * r + 1, k2 = k * k;
  const long long total = (long long)N * cap;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % cap), n = (int)(t / cap);
    uint8_t* row = rows + (size_t)n * row_bytes;
    int cnt = n_picked[n];
    cnt = cnt < cap ? cnt : cap;
    if (i == 0) *reinterpret_cast<int*>(row) = cnt;
    if (i >= cnt) continue;
    const int p = picks[(size_t)n * pick_stride + i];
    reinterpret_cast<int*>(row + 4)[i] = p;
    uint8_t* lab = row + 4 + (size_t)cap * 4 + (size_t)i * k2;
    const int h0 = p / W - r, w0 = p % W - r;
    for (int e = 0; e < k2; ++e) {
      const int h = h0 + e / k, w = w0 + e % k;
      lab[e] = (p >= 0 && h >= 0 && h < H && w >= 0 && w < W) ? gt[((size_t)n * H + h) * W + w] : (uint8_t)255;
    }
  }
}

__global__ void round_rows_apply_kernel(uint8_t* __restrict__ masks, const int* __restrict__ row_image,
                                        const uint8_t* __restrict__ rows, int* __restrict__ n_picked_out, int n_rows, int cap,
                                        int H, int W, int r, size_t row_bytes) {
  const int k = 2 * r + 1, k2 = k * k;
  const long long total = (long long)n_rows * cap;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % cap), j = (int)(t / cap);
    const int img = row_image[j];
    if (img < 0) continue;
    const uint8_t* row = rows + (size_t)j * row_bytes;
    const int cnt = *reinterpret_cast<const int*>(row);
    if (i == 0 && n_picked_out != nullptr) n_picked_out[img] = cnt;
    if (i >= cnt) continue;
    const int p = reinterpret_cast<const int*>(row + 4)[i];
    if (p < 0) continue;
    const uint8_t* lab = row + 4 + (size_t)cap * 4 + (size_t)i * k2;
    const int h0 = p / W - r, w0 = p % W - r;
    for (int e = 0; e < k2; ++e) {
      const int h = h0 + e / k, w = w0 + e % k;
      const uint8_t v = lab[e];   // 255 (unlabeled gt / outside the image) never erases an earlier round's label
      if (v != 255 && h >= 0 && h < H && w >= 0 && w < W) masks[((size_t)img * H + h) * W + w] = v;
    }
  }
}

// Position-sensitive 64-bit checksum: sum over 8-byte little-endian words of word_i * (2*(i + word_offset) + 1) mod 2^64
// (trailing bytes are zero-extended into one last word).  Integer adds commute, so the result does not depend on the
// order the blocks finish in: replicas of the same bytes always agree, on any device.
__global__ void checksum64_kernel(const uint8_t* __restrict__ data, size_t nbytes, unsigned long long word_offset,
                                  unsigned long long* __restrict__ out) {
  const size_t nwords = nbytes / 8;
  unsigned long long acc = 0;
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(data);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
    acc += w[i] * (2ull * (i + word_offset) + 1ull);
  if (blockIdx.x == 0 && threadIdx.x == 0 && (nbytes & 7)) {
    unsigned long long last = 0;
    for (size_t b = 0; b < (nbytes & 7); ++b) last |= (unsigned long long)data[nwords * 8 + b] << (8 * b);
    acc += last * (2ull * (nwords + word_offset) + 1ull);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace halo

using namespace halo;

extern "C" size_t halo_round_row_bytes(int cap, int active_radius) {
  return (cap > 0 && active_radius >= 0) ? round_row_bytes(cap, active_radius) : 0;
}

extern "C" int halo_round_rows_pack(const int* picks, const int* n_picked, const uint8_t* gt, uint8_t* rows, int N, int cap,
                                    int pick_stride, int H, int W, int active_radius, halo_stream_t stream) {
  if (!picks || !n_picked || !gt || !rows || N <= 0 || cap <= 0 || pick_stride < cap || H <= 0 || W <= 0 || active_radius < 0 ||
      (reinterpret_cast<uintptr_t>(rows) & 3)) {
    set_error("halo_round_rows_pack: bad argument");
    return HALO_ERR_BAD_ARG;
  }
  const long long total = (long long)N * cap;
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  round_rows_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(picks, n_picked, gt, rows, N, cap, pick_stride, H, W, active_radius,
                                                               round_row_bytes(cap, active_radius));
  return launch_status("round_rows_pack_kernel");
}

extern "C" int halo_round_rows_apply(uint8_t* masks, const int* row_image, const uint8_t* rows, int* n_picked_out, int n_rows,
                                     int cap, int H, int W, int active_radius, halo_stream_t stream) {
  if (!masks || !row_image || !rows || n_rows <= 0 || cap <= 0 || H <= 0 || W <= 0 || active_radius < 0 ||
      (reinterpret_cast<uintptr_t>(rows) & 3)) {
    set_error("halo_round_rows_apply: bad argument");
    return HALO_ERR_BAD_ARG;
  }
  const long long total = (long long)n_rows * cap;
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  round_rows_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(masks, row_image, rows, n_picked_out, n_rows, cap, H, W,
                                                                active_radius, round_row_bytes(cap, active_radius));
  return launch_status("round_rows_apply_kernel");
}

extern "C" int halo_checksum64(const void* data, size_t nbytes, unsigned long long word_offset, unsigned long long* out,
                               int accumulate, halo_stream_t stream) {
  if (!out || (!data && nbytes) || (reinterpret_cast<uintptr_t>(data) & 7)) {
    set_error("halo_checksum64: bad argument (data must be 8-byte aligned)");
    return HALO_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) HALO_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), st));
  if (nbytes == 0) return HALO_OK;
  const size_t nwords = nbytes / 8 + 1;
  const int grid = (int)((nwords + 255) / 256 < (size_t)sm_count() * 8 ? (nwords + 255) / 256 : (size_t)sm_count() * 8);
  checksum64_kernel<<<grid, 256, 0, st>>>((const uint8_t*)data, nbytes, word_offset, out);
  return launch_status("checksum64_kernel");
}

extern "C" int halo_round_delta_pack(const int* picks, const int* n_picked, const uint8_t* gt, uint8_t* lab, int N, int cap,
                                     int H, int W, int active_radius, halo_stream_t stream) {
  if (!picks || !n_picked || !gt || !lab || N <= 0 || cap <= 0 || H <= 0 || W <= 0 || active_radius < 0) {
    set_error("halo_round_delta_pack: bad argument");
    return HALO_ERR_BAD_ARG;
  }
  const int k2 = (2 * active_radius + 1) * (2 * active_radius + 1);
  const long long total = (long long)N * cap * k2;
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  round_delta_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(picks, n_picked, gt, lab, N, cap, H, W, active_radius);
  return launch_status("round_delta_pack_kernel");
}

extern "C" int halo_round_delta_apply(uint8_t* masks, const int* row_image, const int* picks, const int* n_picked,
                                      const uint8_t* lab, int rows, int cap, int H, int W, int active_radius,
                                      halo_stream_t stream) {
  if (!masks || !row_image || !picks || !n_picked || !lab || rows <= 0 || cap <= 0 || H <= 0 || W <= 0 || active_radius < 0) {
    set_error("halo_round_delta_apply: bad argument");
    return HALO_ERR_BAD_ARG;
  }
  const int k2 = (2 * active_radius + 1) * (2 * active_radius + 1);
  const long long total = (long long)rows * cap * k2;
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  round_delta_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(masks, row_image, picks, n_picked, lab, rows, cap, H, W,
                                                                 active_radius);
  return launch_status("round_delta_apply_kernel");
}
