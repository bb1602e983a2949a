// tcgen05 implicit-GEMM for every dense contraction of the SalUNet denoiser.
//
//   D[m, n] = epilogue( sum_{tap, c} A[pixel(m) + offset(tap), c] * W[n, tap*Cin + c] )
//
// A is a channels-last bf16 activation tensor addressed through a 5-D TMA tensor map, W a K-major bf16 weight
// matrix [N][K] through a 2-D map.  One M tile is a 128-pixel box (bw x bh x bf pixels/frames) so that one TMA box
// load per (tap, 64-channel block) lands exactly one 128-row K-major SWIZZLE_128B operand tile in shared memory;
// out-of-image taps are zero-filled by TMA, which is the convolution's zero padding.  Covers: 3x3 (pad 1),
// 3x3 dilation 2, 3x3 stride 2 with right/bottom pad (via a space-to-depth view of the input), 1x1, the (5,1,1)
// temporal reduction, and plain token linears (bw = 128).
#pragma once
#include "common.cuh"

namespace dsb {

enum GemmAct { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmParams {
    // M tiling (output pixels).  bw * bh * bf == 128, all powers of two.
    int W, H, F;                    // output extent: x in [0,W), y in [0,H), frame in [0,F)
    int bw_log2, bh_log2;           // tile box
    int tiles_x, tiles_y, tiles_f;  // tile grid
    // K loop
    int taps;                       // 1..9
    int cin_blocks;                 // Cin / BK
    int bk;                         // 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
    int ksub;                       // K sub-blocks of bk channels per pipeline stage (amortises barrier round trips)
    int ydim;                       // which tensor-map dimension carries y (2 normally, 3 for the s2d view)
    int tap_off[9][4];              // per tap: coordinate offsets for tensor-map dims 0..3
    // N tiling
    int N, bn;                      // N % bn == 0, bn % 16 == 0, bn <= 256
    int two_cta;                    // 1: CTA pairs (cluster of 2) issue cta_group::2 MMAs with M = 256
    int ab_f16;                     // 1: both operands hold fp16 (not bf16) values -- the output-head GEMMs (DESIGN.md precision policy)
    int out_f16;                    // 1: out_bf16 receives fp16 (not bf16) values (it feeds a GEMM that runs with ab_f16)
    int b_rows_per_frame;           // 0: shared weights; else B rows of frame f start at f * b_rows_per_frame
                                    //    (per-frame attention operands; requires single-frame tiles)
    // frame remap (dead-frame elimination): tile frame tf reads / adds the residual of source frame
    // (tf / f_used) * f_group + tf % f_used (e.g. frames 0..4 of every 9); 0 = identity.  Needs single-frame tiles.
    int f_group, f_used;
    int out_remap;                  // 1: outputs are written at the source frame index too, 0: compact (tf)
    // epilogue: v = act(acc * scale[n] + shift[n] + rowbias[frame][n]) + residual[pix][n]
    const float* scale;             // [N] or null (== 1)
    const float* shift;             // [N] or null (== 0)
    const float* rowbias;           // [F][N] or null
    const float* residual;          // fp32 [pix][N] or null
    int act;
    int epi_transposed;             // 1: stage 32-column chunks through smem for row-coalesced global access
    int halo;                       // 1: halo-tile 3x3 conv -- one TMA box of (8 + 2d) x (16 + 2d) pixels per 64-channel block
    int halo_pw, halo_ph, halo_d;   //    serves all nine taps as row-shifted UMMA descriptors (tile = 8 x 16 pixels, bk = 64)
    int ksplit;                     // > 1: split-K -- work unit = (tile, K slice); slice ks writes its raw fp32 partial to
    long split_stride;              //      out_f32 + ks * split_stride (plain epilogue; splitk_reduce finishes the op)
    float* out_f32;                 // [pix][ldo] or null
    bf16* out_bf16;                 // [pix][ldo] or null
    int ldo;                        // output row pitch in elements
    int out_fmul, out_fadd;         // output frame = f * out_fmul + out_fadd
    float* out2_f32;                // optional second fp32 copy of the output (own frame mapping), e.g. the noise slice
    int out2_fmul, out2_fadd;       //   written both compactly and into frame slot 8 of the 9-frame stage tensor
    // fused 96->1 head (mt_proj + logits): out_head[pix] = sigmoid(sum_n v[n] * head_w[n] + head_b)
    const float* head_w;            // [N] or null
    float head_b;
    float* out_head;
    // attention scores: N == bn == 48 = 2 heads x 18 keys (+12 zero rows); per-head softmax over the 18 keys,
    // written as bf16 probabilities [pix][64] (columns 36..63 zero) -- the A operand of the P.V GEMM
    bf16* out_softmax;
    // folded K / V projections (plan.cu, kernels.cu "projections folded into K / V"): GEMM row m = frame*18 + key j, column
    // n = head*kv_C + c.  kv_mode 1 writes out_bf16 as the score operand [frame][kv_R][kv_C] (row head*18 + j), kv_mode 2
    // as the transposed P.V operand [frame][kv_C][64] (column head*18 + j); padding rows / columns are never written.
    int kv_mode, kv_R, kv_C;
    // per-role clock trace (tools/gemm_trace.py; null in the product path): CTAs 0..3 write, per work unit u < 16,
    // trace[((cta * 3 + role) * 16 + u) * 4 + k] = clock64 at: role 0 (TMA producer) k0 first slot wait, k1 last load issued;
    // role 1 (MMA issuer) k0 accumulator free, k1 first stage landed, k2 last MMA issued; role 2 (epilogue warp 2) k0 start of
    // the wait for the accumulator, k1 accumulator complete, k2 epilogue done.  trace[768 + cta*2 + {0,1}] = kernel start / end.
    unsigned long long* trace;
};

// One-time per-process kernel attribute setup (safe to call repeatedly; called outside stream capture).
int gemm_init();
// Builds tensor maps and launches; returns a cudaError_t-compatible code (0 = ok).
int gemm_launch(const GemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms, cudaStream_t stream);

// Host helper: encode a tensor map (bf16, SWIZZLE_128B when box[0]*2 == 128, SWIZZLE_64B when == 64).
int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

}  // namespace dsb
