// Fused LN-ed tokens -> fc1 -> GELU -> fc2 -> +residual for C = 96 / 192 (see mlp_fused.cu).
#pragma once
#include "gemm_tc.cuh"

namespace dsb {

struct MlpParams {
    int mode;                  // 0: MLP (GELU between the GEMMs); 1: attention (folded bias + per-head softmax)
    int C, bk;                 // channels; K sub-block of GEMM1 (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B)
    int HW, F;                 // tokens per frame, number of (live) frames
    int f_group, f_used;       // frame remap as in GemmParams (0 = identity)
    const float* b1;           // MLP: fc1 bias [2C];  attention: folded score bias [src frames][64]
    const float* b2;           // [C]
    const float* residual;     // fp32 [src frames][HW][C]
    float* out;                // fp32 [src frames][HW][C]
};

struct MlpOp {
    int mode;
    int C, HW, F, f_group, f_used;
    const bf16* A;             // LayerNormed tokens, bf16 [src frames][HW][C]
    const bf16* W1;            // [2C][C]  (fc1.weight, K-major)
    const bf16* W2;            // [C][2C]  (fc2.weight, K-major)
    const float* b1;
    const float* b2;
    const float* residual;
    float* out;
};

struct MlpLaunch {
    MlpParams p;
    CUtensorMap tmA, tmW1, tmW2;
};

int mlp_fused_lower(const MlpOp& op, MlpLaunch* out);
int mlp_fused_run(const MlpLaunch& l, int num_sms, cudaStream_t stream);

}  // namespace dsb
