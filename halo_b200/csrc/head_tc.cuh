// Host-side interface of the tensor-core head kernel (head_fwd_tc.cu).
#pragma once
#include "head_common.cuh"

namespace halo {

// true when the tcgen05/TMA path can run this problem (raw fp32 features, C % 32 == 0, C <= 256, H*W % 4 == 0 ...)
bool head_tc_supported(int feat_kind, int C, int O, int H, int W, const void* feat);
// floats of the tensor-core parameter pack: W planes [2][C/4][NP][4] + cls[4][OP]
size_t head_tc_pack_floats(int O, int C);
// packs the class parameters into the tensor-core layout and launches the kernel on `st`
int head_fwd_tc_launch(HeadArgs a, const float* std_pack, float* wtc, cudaStream_t st);

}  // namespace halo
