// Weight preparation kernels (run once in dsb_finalize_weights).
#pragma once
#include "common.cuh"

namespace dsb {

// src fp32 [N][Cin][taps] (conv weight, taps contiguous) -> dst bf16 (or, f16 = 1, fp16 bits) [N][tap*Cin + c]
int pack_weight_launch(const float* src, int N, int Cin, int taps, bf16* dst, cudaStream_t s, int f16 = 0);
// depthwise: dst[tap*C + c] = src[c*src_stride + src_off + tap]
int pack_dw_launch(const float* src, int C, int taps, int src_stride, int src_off, float* dst, cudaStream_t s);
// src [R][Cc] -> dst [Cc][R]
int transpose_launch(const float* src, int R, int Cc, float* dst, cudaStream_t s);
// eval BatchNorm folded to y = conv*scale + shift; conv_bias may be null
int bn_fold_launch(const float* w, const float* b, const float* mean, const float* var, const float* conv_bias, int C,
                   float eps, float* scale, float* shift, cudaStream_t s);
// conv_in (3x3 pad 1, 1->96) o down1 (3x3 stride 4, 96->96) -> w5[25][96], b5[96]
int stem_compose_launch(const float* w_in, const float* b_in, const float* w_d, const float* b_d, float* w5, float* b5,
                        cudaStream_t s);

}  // namespace dsb
