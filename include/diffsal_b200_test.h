/* libdiffsal_b200_test -- TEST-ONLY entry points (tests/test_kernels_gpu.py, tests/test_membound_kernels_gpu.py).
 *
 * One C entry per hand-written kernel so that every kernel can be compared in isolation with a plain fp32 torch op.
 * They live in their own shared library (diff_sal_b200/libdiffsal_b200_test.so, built from csrc/test_api.cu + csrc/probe.cu
 * and linked against the product library) so that the product library exports only the boundary of include/diffsal_b200.h.
 */
#ifndef DIFFSAL_B200_TEST_H
#define DIFFSAL_B200_TEST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- single-kernel test entry (tests/test_kernels_gpu.py) ------------------------------------------------- */
int dsb_test_conv(int kind, int F, int H, int W, int Cin, int N, int dilation, int T, int kt, const void* A,
                  const void* Wt, const float* scale, const float* shift, const float* rowbias, const float* residual,
                  int act, float* out_f32, void* out_bf16, int out_fmul, int out_fadd, const float* head_w,
                  float head_b, float* out_head, void* stream);

/* memory-bound kernels, one entry each (tests/test_kernels_gpu.py); all buffers are caller-owned device memory */
int dsb_test_groupnorm_swish(const float* x, int F, int HW, int C, const float* gamma, const float* beta,
                             double* scratch, void* out_act, void* out_raw, void* stream);
int dsb_test_layernorm(const float* x, long tokens, int C, const float* gamma, const float* beta, void* out, int hw,
                       int T, int tmax, void* stream);
int dsb_test_q_dwln(const float* x, int F, int H, int W, int C, const float* ng, const float* nb, const float* wq9,
                    const float* qg, const float* qb, void* stats_scratch, void* out, int T, int tmax, void* stream);
int dsb_test_qv_tile(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb, const float* wq9,
                     const float* qg, const float* qb, const float* wv, const float* vg, const float* vb, void* q_out,
                     void* v_out, int T, int tmax, void* stream);
int dsb_test_pool_ln(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb,
                     const float* w, const float* g, const float* b, void* stats_scratch, void* out, int T, int tmax,
                     void* stream);
int dsb_test_av_key(const float* x, const float* a_low, int B, int T, int H, int W, int C, int sk, const float* wk,
                    const float* kg, const float* kb, float* gate, void* out_k, int tmax, void* stream);
int dsb_test_upsample2x(const float* x, int F, int H, int W, int C, void* out, void* stream);
int dsb_test_ms_sum(const float* r0, const float* r1, const float* r2, const float* r3, int B, void* out, void* stream);
int dsb_test_final_up(const float* p, int B, float* out, void* stream);
int dsb_test_stem(const float* x, int B, const float* w_in, const float* b_in, const float* w_d, const float* b_d,
                  float* w5_scratch, float* b5_scratch, float* h0, void* stream);
int dsb_test_temb(const float* t, int B, const float* w0t, const float* b0, const float* w1t, const float* b1,
                  const float* wp0t, const float* bp0, const float* wp1t, const float* bp1, const float* wp2t,
                  const float* bp2, float* tp0, float* tp1, float* tp2, void* stream);

/* fused fc1 -> GELU -> fc2 -> +residual (C = 96 or 192): A bf16 [frames][HW][C], W1 bf16 [2C][C], W2 bf16 [C][2C] */
int dsb_test_mlp_fused(int C, int HW, int F, int f_group, int f_used, const void* A, const void* W1, const void* W2,
                       const float* b1, const float* b2, const float* residual, float* out, void* stream);

/* CTA-pair (cta_group::2) policy of dsb_test_conv: -1 never, 0 automatic, 1 always */
void dsb_test_set_two_cta(int mode);
/* split-K scratch of dsb_test_conv (NULL: never split) and the slice count its last call used (0: not split) */
void dsb_test_set_split_ws(float* ws, long elems);
int dsb_test_last_ksplit(void);
/* 1: the A / Wt buffers of dsb_test_conv hold fp16 (not bf16) values -- the operand type of the output-head GEMMs */
void dsb_test_set_ab_f16(int on);
/* per-role clock64 trace of the GEMM kernel for the next dsb_test_conv calls (device buffer of 784 uint64; layout in
 * csrc/gemm_tc.cuh GemmParams::trace; tools/gemm_trace.py); NULL switches it off */
void dsb_test_set_gemm_trace(void* buf);
/* request the halo-tile 3x3 path in dsb_test_conv, and whether its last call took it */
void dsb_test_set_halo(int on);
int dsb_test_last_halo(void);

/* hardware-semantics probe: UMMA SWIZZLE_128B operand descriptor with a row-shifted start and a non-atom SBO
 * (DESIGN.md section 8 item 1).  A bf16 [512][64], B bf16 [32][64], out fp32 [128][32] (device). */
int dsb_test_umma_shift(const void* A, const void* B, float* out, int shift_rows, int sbo_bytes, int base_offset_mode,
                        void* stream);

/* Projections folded into the K / V projection weights (csrc/kernels.cu "projections folded into K / V"): folds the fp32
 * weights, runs the two projection GEMMs with the attention-operand epilogues.  k_ln / v_ln bf16 [F*18][C]; K1 bf16
 * [F][R][C], V2 bf16 [F][C][64] (zero-filled by the caller: padding is never written); mb fp32 [2][C], cb fp32 [2] out. */
int dsb_test_fold_kv(const float* wq, const float* bq, const float* wk, const float* bk, const float* wp, const float* wv,
                     const float* bv, int C, int F, int R, const void* k_ln, const void* v_ln, void* K1, void* V2,
                     float* mb, float* cb, void* stream);
/* dsb_test_pool_ln + the folded score bias sb fp32 [F][R] (row head*18 + key) = pooled_token . mb[head] + cb[head] */
int dsb_test_pool_ln_sb(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb, const float* w,
                        const float* g, const float* b, void* stats_scratch, void* out, int T, int tmax, const float* mb,
                        const float* cb, float* sb, int R, void* stream);

#ifdef __cplusplus
}
#endif
#endif
