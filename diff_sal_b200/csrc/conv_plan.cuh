// Host-side description of one dense contraction (conv / linear) and its lowering to GemmParams + TMA maps.
#pragma once
#include "gemm_tc.cuh"

namespace dsb {

enum ConvKind {
    CONV_3X3 = 0,        // 3x3, stride 1, dilation d, padding d           (A: [F,H,W,Cin])
    CONV_3X3_S2 = 1,     // zero-pad right/bottom by 1, 3x3, stride 2      (A: [F,2H,2W,Cin], space-to-depth view)
    CONV_1X1 = 2,        // 1x1 conv / token linear                        (A: [F,H,W,Cin])
    CONV_TEMPORAL = 3    // (kt,1,1) conv over the first kt of T frames    (A: [F,T,H,W,Cin], F = clips)
};

struct ConvOp {
    int kind;
    int F, H, W;         // OUTPUT extent
    int Cin, N;
    int dilation;        // CONV_3X3 only
    int T, kt;           // CONV_TEMPORAL only
    const bf16* A;       // activations, channels-last bf16
    const bf16* Wt;      // weights [N][taps*Cin], k = tap*Cin + c, tap = dy*3+dx (or t)
    // epilogue (see GemmParams)
    const float* scale;
    const float* shift;
    const float* rowbias;
    const float* residual;
    int act;
    float* out_f32;
    bf16* out_bf16;
    int ldo;             // 0 -> N
    int out_fmul;        // 0 -> 1
    int out_fadd;
    const float* head_w;
    float head_b;
    float* out_head;
    int b_rows_per_frame;   // per-frame B operand (attention): Wt is [F * b_rows_per_frame][K]
    bf16* out_softmax;      // attention-score epilogue (N == 48)
    int f_group, f_used;    // frame remap: F counts USED frames; tile frame tf -> source (tf/f_used)*f_group + tf%f_used
    int out_remap;          // write outputs at the source frame index (else compact)
    float* out2_f32;        // optional second fp32 output with its own frame mapping
    int out2_fmul, out2_fadd;
    int ab_f16;             // A and Wt hold fp16 values (GemmParams::ab_f16)
    int out_f16;            // out_bf16 receives fp16 values (GemmParams::out_f16); not with split-K
    unsigned long long* trace; // GemmParams::trace (tools only)
    int kv_mode, kv_R, kv_C;   // attention-operand output layouts (GemmParams::kv_mode); needs out_bf16, CONV_1X1, no split-K
    int two_cta;            // -1: never, 0: automatic (pairs when there are enough tiles), 1: force
    int halo;               // 1: request the halo-tile path (CONV_3X3, Cin % 64 == 0, N <= 128, enough 8x16 tiles for CTA
                            //    pairs); silently falls back to the per-tap path when the layer does not qualify
    // split-K (opt-in): fp32 scratch for the per-slice partial sums.  Used when the GEMM has too few tiles to fill
    // the machine and a long K loop; the slices are added in a fixed order by splitk_reduce (bitwise reproducible).
    float* split_ws;
    long split_ws_elems;
    int split_frames_nominal;   // > 0: choose the slice count as if F were this (keeps the K partition, and with it the
                                // rounding, independent of the batch size); 0: use F
};

// second pass of a split-K op: out = epilogue(sum over slices of ws[s][row][n]) with the op's own epilogue
struct SplitReduce {
    const float* ws;
    int S;                  // 0: not split
    long slab;              // elements per slice
    int rows, N, HW;
    const float *scale, *shift, *rowbias, *residual;
    int act;
    float* out_f32;
    bf16* out_bf16;
    float* out2_f32;
    int ldo, out_fmul, out_fadd, out2_fmul, out2_fadd;
};

struct ConvLaunch {
    GemmParams p;
    CUtensorMap tmA, tmB;
    SplitReduce split;
};

// Lowers `op`; returns 0 or a negative error.
int conv_lower(const ConvOp& op, ConvLaunch* out);
int conv_run(const ConvLaunch& l, int num_sms, cudaStream_t stream);

}  // namespace dsb
