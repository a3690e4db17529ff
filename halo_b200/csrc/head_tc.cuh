// Host-side interface of the tensor-core head kernel (head_fwd_tc.cu).
#pragma once
#include "head_common.cuh"

namespace halo {

// true when the tcgen05/TMA path can run this problem (raw fp32 features, C % 32 == 0, C <= 256, H*W % 4 == 0 ...)
bool head_tc_supported(int feat_kind, int C, int O, int H, int W, const void* feat);
bool head_tc_shape_ok(int feat_kind, int C, int O, int H, int W, const void* feat);   // the same without asking the driver
// floats of the tensor-core parameter pack: W planes [2][C/4][NP][4] + cls[4][OP]
size_t head_tc_pack_floats(int O, int C);
// packs the class parameters into the tensor-core layout and launches the kernel on `st`
int head_fwd_tc_launch(HeadArgs a, const float* std_pack, float* wtc, cudaStream_t st);

// std pack (head_pack_kernel) -> tensor-core operand planes
int head_pack_tc_launch(const float* std_pack, float* wtc, int C, int CPAD, int O, cudaStream_t st);

// tensor-core pixel pass of the backward (head_bwd_tc.cu)
bool head_bwd_tc_supported(int C, int O, int H, int W, const void* feat, const void* dfeat);
int head_bwd_tc_grid(int N, int HW);
int head_bwd_tc_launch(const float* feat, const float* dlogits, float* dfeat, float* G, float* cls_part, const float* std_pack,
                       const float* wtc, float* w2, float c, int N, int C, int O, int H, int W, int grid, cudaStream_t st);

// tensor-core weight gradient dW = G^T.U (head_bwd_dw_tc.cu)
bool head_bwd_dw_tc_supported(int C, int O, int H, int W, const void* feat, const void* G);
int head_bwd_dw_tc_grid(int N, int HW);
int head_bwd_dw_tc_launch(const float* feat, const float* G, float* dw_part, int N, int C, int O, int H, int W, int CP, int grid,
                          cudaStream_t st);

// streaming backward (head_bwd_stream_tc.cu): one pass over the features for du AND dW, from the contractions the
// forward saved (HeadArgs::saved)
bool head_bwd_stream_supported(int C, int O, int H, int W, const void* feat, const void* dfeat);
bool head_bwd_stream_shape_ok(int C, int O, int H, int W, const void* feat, const void* dfeat);
int head_bwd_stream_grid(int N, int HW);
int head_bwd_stream_launch(const float* feat, const float* dlogits, const float* saved, float* dfeat, float* dw_part,
                           float* cls_part, const float* std_pack, float* w2, float c, int N, int C, int CPAD, int O, int H, int W,
                           int CP, int grid, cudaStream_t st);
// transposed parameter planes [2 (hi,lo)][2*OP/4][C][4] for the du GEMMs (head_bwd_tc.cu)
int head_pack_bwd_planes_launch(const float* std_pack, float* w2, int C, int OP, cudaStream_t st);

}  // namespace halo
