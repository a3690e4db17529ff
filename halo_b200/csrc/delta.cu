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

}  // namespace halo

using namespace halo;

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
