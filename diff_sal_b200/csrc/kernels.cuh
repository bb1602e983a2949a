// Memory-bound kernels of the SalUNet denoiser (everything that is not a dense contraction).
// All activations are channels-last: frames [F][H][W][C]; a clip's frames are contiguous (f = b*T + t).
#pragma once
#include "common.cuh"

namespace dsb {

struct QdwTables { const float* wg; const float* wb; const float* wbs; };     // [9][C], [9][C], [C]

struct TembWeights {
    const float* w0; const float* b0;      // temb.dense.0  [384,96]
    const float* w1; const float* b1;      // temb.dense.1  [384,384]
    const float* wp[3]; const float* bp[3];  // res_encoder.i.0.temb_proj [Cout_i,384]
    int cout[3];
};

// t[B] -> per-block temb projections tp_i[B][Cout_i]   (sal_unet.py:15-38,304-307,129)
int temb_launch(const float* t, int B, const TembWeights& w, float* const tp[3], cudaStream_t s);

// conv_in (3x3, pad 1) composed with down1 (pad right/bottom, 3x3, stride 4) = one 5x5 stride-4 conv 1->96
// x[B][224][384] -> h0[B][56][96][96]   (sal_unet.py:292-293)
int stem_launch(const float* x, int B, const float* w5, const float* b5, float* h0, cudaStream_t s);

// GroupNorm(32, eps 1e-6): acc[F][<=64 splits][32][2] (double) receives per-split sums / sums of squares (no atomics,
// bitwise reproducible); gn_apply adds the splits in a fixed order
int gn_stats_launch(const float* x, int F, int HW, int C, double* acc, cudaStream_t s);
// out_act = bf16(swish(GN(x))) ; out_raw = bf16(x) (optional)
int gn_apply_launch(const float* x, int F, int HW, int C, const double* acc, const float* gamma, const float* beta,
                    bf16* out_act, bf16* out_raw, cudaStream_t s);

// the two passes above in one launch (a thread-block cluster per frame, statistics exchanged through DSMEM)
int gn_fused_launch(const float* x, int F, int HW, int C, const float* gamma, const float* beta, bf16* out_act, bf16* out_raw,
                    cudaStream_t s);

// bilinear x2, align_corners=False: fp32 [F][H][W][C] -> bf16 [F][2H][2W][C]
int upsample2x_launch(const float* x, int F, int H, int W, int C, bf16* out, cudaStream_t s);

// LayerNorm(eps 1e-5) statistics per token: stats[token] = (mean, rstd)
// tokens of frames with (frame % T) >= tmax are skipped (hw = tokens per frame)
int ln_stats_launch(const float* x, long tokens, int C, float2* stats, int hw, int T, int tmax, cudaStream_t s);
// out = bf16(LN(x) * gamma + beta); tokens of frames with (frame % T) >= tmax are skipped (hw = tokens per frame)
// f16 = 1: the output holds fp16 (not bf16) values -- the operand of a GEMM that runs with GemmParams::ab_f16
int ln_apply_launch(const float* x, long tokens, int C, const float* gamma, const float* beta, bf16* out, int hw,
                    int T, int tmax, cudaStream_t s, int f16 = 0);

// q = LN_q( depthwise3x3( LN_norm(x) ) )  -> bf16 [tokens][C]      (attention.py:36-48,92 ; transformer.py:151)
// tb (optional): tap tables with the LayerNorm affine folded in (q_dw_prep_launch); with them the narrow stages
// (C = 96, 192) take the second-generation tiled kernel.
int q_dw_prep_launch(const float* w9, const float* ng, const float* nb, int C, float* wg, float* wb, float* wbs, cudaStream_t s);
int q_dwln_launch(const float* x, const float2* stats, int F, int H, int W, int C, const float* ng, const float* nb,
                  const float* wq /*[9][C]*/, const QdwTables* tb, const float* qg, const float* qb, bf16* out, int T, int tmax,
                  cudaStream_t s);

// taps of a depthwise SxS pooling with the pre-attention LayerNorm affine folded in: wg[p][c] = w[p][c] * g[c],
// wbs[c] = b[c] * sum_p w[p][c]
int dw_affine_prep_launch(const float* w, const float* g, const float* b, int taps, int C, float* wg, float* wbs, cudaStream_t s);
// q and v of a narrow stage (C = 96 @ 56x96, C = 192 @ 28x48) in one pass over the stage input, LayerNorm statistics
// computed in the kernel (replaces ln_stats + q_dwln + pool_ln for the audio-visual configuration); -37 if the geometry
// is not one of the two
int qv_tile_launch(const float* x, int F, int H, int W, int C, int s_, const QdwTables& tb, const float* qg, const float* qb,
                   const float* wvg, const float* wvbs, const float* vg, const float* vb, bf16* q_out, bf16* v_out, int T,
                   int tmax, cudaStream_t s);

// v (or visual-only k) = LN( depthwise sxs stride s ( LN_norm(x) ) ) -> bf16 [F*18][C]   (attention.py:53-76,93)
// optional by-product of the K pooling kernels: the folded score bias of every key ("projections folded into K / V"),
// sb[frame][R] at row head*18 + key = k_ln . mb[head] + cb[head]; sb == null: not produced
struct ScoreBias {
    const float* mb;   // [2][C]
    const float* cb;   // [2]
    float* sb;         // [frames][R]
    int R;
};
int pool_ln_launch(const float* x, const float2* stats, int F, int H, int W, int C, int s_, const float* ng,
                   const float* nb, const float* wv /*[s*s][C]*/, const float* vg, const float* vb, bf16* out, int T,
                   int tmax, cudaStream_t s, ScoreBias sbv = ScoreBias{nullptr, nullptr, nullptr, 0});

// audio gate: g[b][c][y][x] = softmax_x( mean_t( a[b,t,y/r,x/r,c] * x[b,t,y,x,c] ) )   (transformer.py:140-144)
int av_gate_launch(const float* x, const float* a_low, int B, int T, int H, int W, int C, float* g, cudaStream_t s);
// channel-major copy of the aligned audio map: a_low[(b*T+t)*84 + p][c] -> a_cm[b][c][t*84 + p]
int audio_cmajor_launch(const float* a_low, int B, int T, int C, float* a_cm, cudaStream_t s);
// k = LN( depthwise sxs stride s ( scrambled (a*g) ) ) -> bf16 [B*T*18][C]   (transformer.py:145-146, attention.py:89-91)
// a_cm: the CHANNEL-MAJOR audio map (audio_cmajor_launch)
int kpool_av_launch(const float* g, const float* a_cm, int B, int T, int H, int W, int C, int s_,
                    const float* wk /*[s*s][C]*/, const float* kg, const float* kb, bf16* out, int tmax, cudaStream_t s, ScoreBias sbv = ScoreBias{nullptr, nullptr, nullptr, 0});

// per-frame tensor-core operands of the 18-key, 2-head attention:
//   KB[f][h*18+j][c] = scale * K[f,j,c] if c in head h else 0        ([F][48][C], rows 36..47 zero)
//   VB[f][c][h*18+j] = V[f,j,c]         if c in head h else 0        ([F][C][64], cols 36..63 zero)
int attn_operands_launch(const float* kp, const float* vp, int F, int C, float scale, bf16* KB, bf16* VB,
                         cudaStream_t s);

// query / output projections folded into the key / value projection WEIGHTS (once per weight set; see kernels.cu):
// MK, MV fp32 [2][C][C] (row h*C + c, K-major), cK, cV, mb fp32 [2*C], cb fp32 [2]
int fold_weights_launch(const float* wq, const float* bq, const float* wk, const float* bk, const float* wp, const float* wv,
                        const float* bv, int C, float scale, float* MK, float* MV, float* cK, float* cV, float* mb, float* cb,
                        cudaStream_t s);
// folded GEMM outputs Kf / Vf [F*18][2C] -> K1 bf16 [F][R][C], sb fp32 [F][R], V2 bf16 [F][C][64] (R = 48 or 64 key rows)
int kv_pack_launch(const float* Kf, const float* Vf, const bf16* k_ln, const float* mb, const float* cb, int F, int C, int R,
                   int T, int tmax, bf16* K1, float* sb, bf16* V2, cudaStream_t s);

// folded per-frame attention operands for the fused attention chain (Wq folded into K, Wp into V; see kernels.cu)
int attn_fold_launch(const float* kp, const float* vp, const float* wq, const float* bq, const float* wpT, int F, int C,
                     float scale, int T, int tmax, bf16* K1, float* sb, bf16* V2, cudaStream_t s);

// S[b][y][x][c] = sum_i bilinear(r_i -> 112x192)[b,y,x,c]  (fp16 bits: operand of the ab_f16 mt_proj GEMM)   (sal_unet.py:482-487)
int ms_sum_launch(const float* const r[4], int B, bf16* S, cudaStream_t s);
// bilinear x2 on a single-channel map: p[B][112][192] -> out[B][224][384]   (sal_unet.py:325-327)
int final_up_launch(const float* p, int B, float* out, cudaStream_t s);

// out = c[0]*in[0] + c[1]*in[1] + ... (nin <= 4) + cn*noise       (sampler updates, n elements)
int axpy_launch(int nin, const float* const in[4], const float c[4], const float* noise, float cn, float* out, long n,
                cudaStream_t s);

// clamp(x, 0, 1) (inverse_data_transform, datasets/__init__.py:26-35) and per-map min-max -> uint8 (normalize_data,
// util/utils.py:11-16); either output may be null.  n = pixels per map.
int postprocess_launch(const float* x, int B, int n, float* clamped, uint8_t* u8, cudaStream_t s);

// per-sample error norm of DPM-Solver's adaptive step-size control (sampler.py:996-999): out[B]
int adaptive_error_launch(const float* xl, const float* xh, const float* xp, int B, int n, float atol, float rtol, float* out,
                          cudaStream_t s);

// *out += 64-bit content fingerprint of a 16-byte-aligned device buffer (out must be zeroed by the caller)
int content_hash_launch(const void* p, size_t bytes, unsigned long long seed, unsigned long long* out, cudaStream_t s);

// vis[B][C][Tv][HW] fp32 -> frames[(b*T + t)][HW][C] for t < Tv  (T = frames per clip in dst)
int nct_to_frames_launch(const float* vis, int B, int C, int Tv, int HW, int T, float* dst, cudaStream_t s);
// audio[B][512][T][84] fp32 -> tokens[(b*T+t)*84 + p][512] bf16
int audio_tokens_launch(const float* audio, int B, int T, bf16* out, cudaStream_t s);
// fp32 -> bf16 elementwise
int to_bf16_launch(const float* x, long n, bf16* out, cudaStream_t s);

// audio transformer glue (models/audio_attention.py:55-90): head split of the bias-free qkv projection, row softmax
// over 756 of 768 columns, final LayerNorm written channels-first
int qkv_split_launch(const bf16* qkv, int B, int n, int npad, bf16* Qh, bf16* Kh, bf16* Vt, cudaStream_t s);
int softmax_rows_launch(const float* S, long rows, int valid, int ld, bf16* P, cudaStream_t s);
int ln_nct_launch(const float* x, int B, int n, int C, const float* gamma, const float* beta, float* out, cudaStream_t s);

// sampler correctors: in-place clamp (util/denoising.py:54) and DPM_Solver.dynamic_thresholding_fn (sampler.py:417-426)
int clamp_launch(float* x, long n, float lo, float hi, cudaStream_t s);
int dyn_threshold_launch(float* x, int B, int n, int k, float w, float max_val, cudaStream_t s);

}  // namespace dsb
