// C-ABI entry points that expose single kernels for the per-kernel parity tests (tests/test_kernels_gpu.py).
#include "conv_plan.cuh"

using namespace dsb;

static int num_sms_cached() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

static int g_test_halo = 0;
// 1: request the halo-tile path in dsb_test_conv; dsb_test_last_halo() tells whether the last lowered op took it
static int g_test_last_halo = 0;
extern "C" void dsb_test_set_halo(int on) { g_test_halo = on; }
extern "C" int dsb_test_last_halo(void) { return g_test_last_halo; }
static int g_test_two_cta = 0;
// -1: never use CTA pairs, 0: automatic, 1: always (for the per-kernel tests)
extern "C" void dsb_test_set_two_cta(int mode) { g_test_two_cta = mode; }

static int g_test_ab_f16 = 0;
// 1: A / Wt of dsb_test_conv hold fp16 values (GemmParams::ab_f16)
extern "C" void dsb_test_set_ab_f16(int on) { g_test_ab_f16 = on; }

static unsigned long long* g_test_trace = nullptr;
// per-role clock trace buffer (GemmParams::trace; 784 x u64) for the next dsb_test_conv calls; NULL: off
extern "C" void dsb_test_set_gemm_trace(void* buf) { g_test_trace = (unsigned long long*)buf; }

static float* g_test_split_ws = nullptr;
static long g_test_split_elems = 0;
// split-K scratch for dsb_test_conv (NULL: never split); returns the slice count the LAST lowered op used
static int g_test_last_ksplit = 0;
extern "C" void dsb_test_set_split_ws(float* ws, long elems) { g_test_split_ws = ws; g_test_split_elems = elems; }
extern "C" int dsb_test_last_ksplit(void) { return g_test_last_ksplit; }

extern "C" int dsb_test_conv(int kind, int F, int H, int W, int Cin, int N, int dilation, int T, int kt,
                             const void* A, const void* Wt, const float* scale, const float* shift,
                             const float* rowbias, const float* residual, int act, float* out_f32, void* out_bf16,
                             int out_fmul, int out_fadd, const float* head_w, float head_b, float* out_head,
                             void* stream) {
    ConvOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.F = F; op.H = H; op.W = W; op.Cin = Cin; op.N = N; op.dilation = dilation; op.T = T; op.kt = kt;
    op.A = (const bf16*)A; op.Wt = (const bf16*)Wt;
    op.scale = scale; op.shift = shift; op.rowbias = rowbias; op.residual = residual; op.act = act;
    op.out_f32 = out_f32; op.out_bf16 = (bf16*)out_bf16; op.out_fmul = out_fmul; op.out_fadd = out_fadd;
    op.head_w = head_w; op.head_b = head_b; op.out_head = out_head;
    op.two_cta = g_test_two_cta;
    op.split_ws = g_test_split_ws; op.split_ws_elems = g_test_split_elems;
    op.halo = g_test_halo;
    op.ab_f16 = g_test_ab_f16;
    op.trace = g_test_trace;
    ConvLaunch l;
    int r = conv_lower(op, &l);
    if (r) return r;
    g_test_last_ksplit = l.split.S;
    g_test_last_halo = l.p.halo;
    return conv_run(l, num_sms_cached(), (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------------
// single-kernel entries for the memory-bound kernels (all buffers are caller-owned device memory)
#include "kernels.cuh"
#include "weights.cuh"

extern "C" int dsb_test_groupnorm_swish(const float* x, int F, int HW, int C, const float* gamma, const float* beta,
                                        double* scratch /*[F*64*64]*/, void* out_act, void* out_raw, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!scratch)                                       // NULL scratch: the one-launch cluster kernel the plan uses
        return gn_fused_launch(x, F, HW, C, gamma, beta, (bf16*)out_act, (bf16*)out_raw, s);
    if (int r = gn_stats_launch(x, F, HW, C, scratch, s)) return r;
    return gn_apply_launch(x, F, HW, C, scratch, gamma, beta, (bf16*)out_act, (bf16*)out_raw, s);
}

extern "C" int dsb_test_layernorm(const float* x, long tokens, int C, const float* gamma, const float* beta, void* out,
                                  int hw, int T, int tmax, void* stream) {
    return ln_apply_launch(x, tokens, C, gamma, beta, (bf16*)out, hw, T, tmax, (cudaStream_t)stream);
}

extern "C" int dsb_test_q_dwln(const float* x, int F, int H, int W, int C, const float* ng, const float* nb,
                               const float* wq9, const float* qg, const float* qb, void* stats_scratch, void* out,
                               int T, int tmax, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = ln_stats_launch(x, (long)F * H * W, C, (float2*)stats_scratch, H * W, T, tmax, s)) return r;
    // folded tap tables of the second-generation tiled kernel (prepared per weight set in the product)
    static float* tb = nullptr;
    if (!tb && cudaMalloc(&tb, (size_t)(19 * 768) * sizeof(float)) != cudaSuccess) return -1;
    if (int r = q_dw_prep_launch(wq9, ng, nb, C, tb, tb + 9 * 768, tb + 18 * 768, s)) return r;
    const QdwTables qt = {tb, tb + 9 * 768, tb + 18 * 768};
    return q_dwln_launch(x, (const float2*)stats_scratch, F, H, W, C, ng, nb, wq9, &qt, qg, qb, (bf16*)out, T, tmax, s);
}

extern "C" int dsb_test_qv_tile(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb,
                                const float* wq9, const float* qg, const float* qb, const float* wv, const float* vg,
                                const float* vb, void* q_out, void* v_out, int T, int tmax, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    static float* tb = nullptr;                        // folded tap tables (prepared per weight set in the product)
    const size_t cap = (size_t)(19 + 257) * 768;
    if (!tb && cudaMalloc(&tb, cap * sizeof(float)) != cudaSuccess) return -1;
    float *wg = tb, *wb = tb + 9 * 768, *wbs = tb + 18 * 768, *wvg = tb + 19 * 768, *wvbs = tb + (19 + 256) * 768;
    if (int r = q_dw_prep_launch(wq9, ng, nb, C, wg, wb, wbs, s)) return r;
    if (int r = dw_affine_prep_launch(wv, ng, nb, sk * sk, C, wvg, wvbs, s)) return r;
    const QdwTables qt = {wg, wb, wbs};
    return qv_tile_launch(x, F, H, W, C, sk, qt, qg, qb, wvg, wvbs, vg, vb, (bf16*)q_out, (bf16*)v_out, T, tmax, s);
}

extern "C" int dsb_test_pool_ln(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb,
                                const float* w, const float* g, const float* b, void* stats_scratch, void* out, int T,
                                int tmax, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = ln_stats_launch(x, (long)F * H * W, C, (float2*)stats_scratch, H * W, T, tmax, s)) return r;
    return pool_ln_launch(x, (const float2*)stats_scratch, F, H, W, C, sk, ng, nb, w, g, b, (bf16*)out, T, tmax, s);
}

extern "C" int dsb_test_av_key(const float* x, const float* a_low, int B, int T, int H, int W, int C, int sk,
                               const float* wk, const float* kg, const float* kb, float* gate, void* out_k, int tmax,
                               void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = av_gate_launch(x, a_low, B, T, H, W, C, gate, s)) return r;
    static float* acm = nullptr;                      // channel-major audio map (prepared at conditioning time in the product)
    static size_t acm_cap = 0;
    const size_t need = (size_t)B * T * 84 * C;
    if (need > acm_cap) {
        if (acm) cudaFree(acm);
        if (cudaMalloc(&acm, need * sizeof(float)) != cudaSuccess) return -1;
        acm_cap = need;
    }
    if (int r = audio_cmajor_launch(a_low, B, T, C, acm, s)) return r;
    return kpool_av_launch(gate, acm, B, T, H, W, C, sk, wk, kg, kb, (bf16*)out_k, tmax, s);
}

extern "C" int dsb_test_upsample2x(const float* x, int F, int H, int W, int C, void* out, void* stream) {
    return upsample2x_launch(x, F, H, W, C, (bf16*)out, (cudaStream_t)stream);
}

extern "C" int dsb_test_ms_sum(const float* r0, const float* r1, const float* r2, const float* r3, int B, void* out,
                               void* stream) {
    const float* r[4] = {r0, r1, r2, r3};
    return ms_sum_launch(r, B, (bf16*)out, (cudaStream_t)stream);
}

extern "C" int dsb_test_final_up(const float* p, int B, float* out, void* stream) {
    return final_up_launch(p, B, out, (cudaStream_t)stream);
}

extern "C" int dsb_test_stem(const float* x, int B, const float* w_in, const float* b_in, const float* w_d,
                             const float* b_d, float* w5_scratch /*[2400]*/, float* b5_scratch /*[96]*/, float* h0,
                             void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = stem_compose_launch(w_in, b_in, w_d, b_d, w5_scratch, b5_scratch, s)) return r;
    return stem_launch(x, B, w5_scratch, b5_scratch, h0, s);
}

// weights are given TRANSPOSED ([in][out]) like the handle prepares them
extern "C" int dsb_test_temb(const float* t, int B, const float* w0t, const float* b0, const float* w1t, const float* b1,
                             const float* wp0t, const float* bp0, const float* wp1t, const float* bp1, const float* wp2t,
                             const float* bp2, float* tp0, float* tp1, float* tp2, void* stream) {
    TembWeights w;
    w.w0 = w0t; w.b0 = b0; w.w1 = w1t; w.b1 = b1;
    w.wp[0] = wp0t; w.wp[1] = wp1t; w.wp[2] = wp2t;
    w.bp[0] = bp0; w.bp[1] = bp1; w.bp[2] = bp2;
    w.cout[0] = 192; w.cout[1] = 384; w.cout[2] = 768;
    float* tp[3] = {tp0, tp1, tp2};
    return temb_launch(t, B, w, tp, (cudaStream_t)stream);
}

#include "mlp_fused.cuh"

extern "C" int dsb_test_mlp_fused(int C, int HW, int F, int f_group, int f_used, const void* A, const void* W1,
                                  const void* W2, const float* b1, const float* b2, const float* residual, float* out,
                                  void* stream) {
    MlpOp op;
    memset(&op, 0, sizeof(op));
    op.C = C; op.HW = HW; op.F = F; op.f_group = f_group; op.f_used = f_used;
    op.A = (const bf16*)A; op.W1 = (const bf16*)W1; op.W2 = (const bf16*)W2;
    op.b1 = b1; op.b2 = b2; op.residual = residual; op.out = out;
    MlpLaunch l;
    if (int r = mlp_fused_lower(op, &l)) return r;
    return mlp_fused_run(l, num_sms_cached(), (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------------
// query / output projections folded into the K / V projection weights (kernels.cu, plan.cu): weight folding, the two
// projection GEMMs with the operand-layout epilogues, and the score bias produced by the K pooling kernel.
// k_ln / v_ln: bf16 [F*18][C]; K1: bf16 [F][R][C], V2: bf16 [F][C][64] (caller zero-fills both); mb fp32 [2][C], cb fp32 [2]
extern "C" int dsb_test_fold_kv(const float* wq, const float* bq, const float* wk, const float* bk, const float* wp,
                                const float* wv, const float* bv, int C, int F, int R, const void* k_ln, const void* v_ln,
                                void* K1, void* V2, float* mb, float* cb, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    float *MK = nullptr, *MV = nullptr, *cK = nullptr, *cV = nullptr;
    bf16 *pk = nullptr, *pv = nullptr;
    const size_t cc = (size_t)2 * C * C;
    if (cudaMalloc(&MK, cc * 4) || cudaMalloc(&MV, cc * 4) || cudaMalloc(&cK, 2 * C * 4) || cudaMalloc(&cV, 2 * C * 4) ||
        cudaMalloc(&pk, cc * 2) || cudaMalloc(&pv, cc * 2))
        return -1;
    int r = fold_weights_launch(wq, bq, wk, bk, wp, wv, bv, C, 1.0f / sqrtf((float)C), MK, MV, cK, cV, mb, cb, s);
    if (!r) r = pack_weight_launch(MK, 2 * C, C, 1, pk, s);
    if (!r) r = pack_weight_launch(MV, 2 * C, C, 1, pv, s);
    for (int which = 0; which < 2 && !r; ++which) {
        ConvOp op;
        memset(&op, 0, sizeof(op));
        op.kind = CONV_1X1; op.F = 1; op.H = 1; op.W = F * 18; op.Cin = C; op.N = 2 * C; op.dilation = 1; op.T = 1; op.kt = 1;
        op.A = (const bf16*)(which ? v_ln : k_ln); op.Wt = which ? pv : pk;
        op.shift = which ? cV : cK;
        op.out_bf16 = (bf16*)(which ? V2 : K1);
        op.kv_mode = which ? 2 : 1; op.kv_R = R; op.kv_C = C;
        ConvLaunch l;
        r = conv_lower(op, &l);
        if (!r) r = conv_run(l, num_sms_cached(), s);
    }
    cudaStreamSynchronize(s);
    cudaFree(MK); cudaFree(MV); cudaFree(cK); cudaFree(cV); cudaFree(pk); cudaFree(pv);
    return r;
}

// dsb_test_pool_ln with the folded score bias as a by-product: sb fp32 [F][R] rows head*18 + key (caller zero-fills)
extern "C" int dsb_test_pool_ln_sb(const float* x, int F, int H, int W, int C, int sk, const float* ng, const float* nb,
                                   const float* w, const float* g, const float* b, void* stats_scratch, void* out, int T,
                                   int tmax, const float* mb, const float* cb, float* sb, int R, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = ln_stats_launch(x, (long)F * H * W, C, (float2*)stats_scratch, H * W, T, tmax, s)) return r;
    return pool_ln_launch(x, (const float2*)stats_scratch, F, H, W, C, sk, ng, nb, w, g, b, (bf16*)out, T, tmax, s,
                          ScoreBias{mb, cb, sb, R});
}
