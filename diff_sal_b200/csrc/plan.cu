// Handle, weight store, workspace plan and the C ABI of libdiffsal_b200 (include/diffsal_b200.h).
//
// One denoiser evaluation (reference: SalUNet.forward, sal_unet.py:302-328) is lowered at dsb_set_condition time
// into a flat list of kernel launches over a static workspace; dsb_denoise replays the list on the caller's stream
// and dsb_sample replays it inside the sampler program (optionally captured into a CUDA graph).
#include <diffsal_b200.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "conv_plan.cuh"
#include "kernels.cuh"
#include "mlp_fused.cuh"
#include "weights.cuh"

using namespace dsb;

namespace {

constexpr int kStageC[4] = {768, 384, 192, 96};
constexpr int kStageH[4] = {7, 14, 28, 56};
constexpr int kStageW[4] = {12, 24, 48, 96};
constexpr int kStageS[4] = {2, 4, 8, 16};
constexpr int kT = 9;            // frames per clip inside the decoder (8 MViT slices + the noise slice)
constexpr int kTv = 8;
constexpr int kReduce = 5;       // ReduceTemp kernel == stride (cfgs/audio_visual.py:62)
constexpr size_t frame_elems(int i) { return (size_t)64512 << i; }   // C_i * h_i * w_i of decoder stage i
constexpr size_t kMaxFrame = (size_t)64512 << 3;
constexpr long kMapElems = 224L * 384L;
constexpr long kSplitWsPerClip = 336L * 768L * 8L;      // largest rows x N x slices of a split-K layer, per clip

struct Weight {
    float* p = nullptr;
    std::vector<int64_t> shape;
    long numel = 0;
};

typedef std::function<int(cudaStream_t)> Launch;

struct SamplerGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t last_use = 0;
};
constexpr size_t kMaxGraphs = 8;              // cached (program, batch) graphs per handle, least recently used evicted

}  // namespace

struct dsb_handle {
    dsb_config cfg;
    int device = 0;
    int num_sms = 148;
    std::string err;
    std::map<std::string, Weight> w;
    bool finalized = false;
    std::vector<void*> allocs;
    size_t alloc_bytes = 0;
    size_t weight_bytes = 0;

    // ---- derived weights
    std::map<std::string, bf16*> wpack;          // K-major bf16 GEMM weights by reference key
    std::map<std::string, float*> wf;            // derived fp32 tables (bn scale/shift, depthwise taps, stem)

    // ---- workspace (sized for cfg.max_batch)
    bool ws_ready = false;
    float* tp[3] = {nullptr, nullptr, nullptr};  // temb projections
    float* h0 = nullptr;
    double* gn_acc = nullptr;                    // [6][B][64 splits][32][2]
    bf16 *enc_act = nullptr, *enc_raw = nullptr, *enc_res = nullptr;
    float *enc_c1 = nullptr, *enc_sc = nullptr;
    float* enc_d[3] = {nullptr, nullptr, nullptr};
    float* back[3] = {nullptr, nullptr, nullptr};
    float* a_low[4] = {nullptr, nullptr, nullptr, nullptr};
    float* a_cm[4] = {nullptr, nullptr, nullptr, nullptr};     // channel-major copies for the K-source gather
    bf16* audio_tok = nullptr;
    float *X[4] = {}, *X1[4] = {}, *X2[4] = {};
    float* r[4] = {};
    bf16 *up = nullptr, *mid = nullptr, *q_ln = nullptr, *Qp = nullptr, *k_ln = nullptr, *v_ln = nullptr;
    float *Kp = nullptr, *Vp = nullptr, *gate = nullptr;
    bf16 *K1 = nullptr, *V2 = nullptr;            // folded attention operands (narrow stages)
    bf16 *Ks[4] = {}, *Vs[4] = {};                // per-stage attention operands written by the K / V projection epilogues
    float* sbs[4] = {};                           //   (padding rows / columns stay zero from allocation) + folded score bias
    float* sbias = nullptr;
    bf16 *KB = nullptr, *VB = nullptr, *P = nullptr, *o = nullptr, *ln2 = nullptr, *hid = nullptr, *lnm = nullptr;
    float2* lnstats = nullptr;
    bf16* S = nullptr;
    float* p = nullptr;
    float* sbuf[8] = {};                          // sampler buffers
    float* splitws = nullptr;                     // split-K partial sums of the main-stream GEMMs
    float* splitws_head = nullptr;                // ... and of the ReduceTemp GEMMs, which run on the head stream
    float* t_all = nullptr;                       // [max evals][B]
    int t_all_cap = 0;

    // ---- per-condition state
    int B = 0;
    bool has_audio = false;
    std::vector<Launch> prog;                     // one denoiser evaluation
    std::vector<std::string> prog_name;           // per launch: label
    std::vector<double> prog_flops;               // per launch: algorithmic FLOPs (0 for memory-bound kernels)
    std::vector<double> prog_bytes;               // per launch: algorithmic bytes (memory-bound kernels; 0 if not stated)
    // Independent branches of an evaluation (the q / k / v token producers of a block) run on side streams:
    // kind 0 = launch on stream prog_stream[i]; 1 = record event prog_ev[i] on that stream; 2 = that stream waits for it
    std::vector<int> prog_kind, prog_stream, prog_ev;
    std::vector<int> prog_pdl;                    // per launch: programmatic-dependent-launch mode of its zone (0 off, 2 all kernels)
    int prog_launches = 0;
    std::vector<int> profile_idx;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // V branch, K branch, output-head branch
    cudaEvent_t ev[8] = {};
    int64_t cond_launches = 0;
    const float* cur_x = nullptr;
    const float* cur_t = nullptr;
    float* cur_out = nullptr;
    int64_t last_launches = 0;
    std::map<std::string, SamplerGraph> graphs;
    uint64_t graph_clock = 0;
    uint64_t cond_epoch = 0;
    // handle-owned staging so that a captured graph never bakes in a caller pointer
    float* noise_stage = nullptr;                 // [noise_slabs][max_batch][224*384]
    size_t noise_slabs = 0;
    uint8_t* u8_stage = nullptr;                  // [max_batch][224*384] normalised maps of DSB_OP_POSTPROCESS
    unsigned long long* hash_dev = nullptr;
    std::vector<float> ts_uploaded;               // model times currently in t_all (skip the upload when unchanged)
};

namespace {

int fail(dsb_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(h, DSB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));       \
    } while (0)

template <class T>
int dev_alloc(dsb_handle* h, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = ((count * sizeof(T) + 255) / 256) * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(h, DSB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    h->alloc_bytes += bytes;
    *out = reinterpret_cast<T*>(p);
    return 0;
}

const Weight* find_w(dsb_handle* h, const std::string& key) {
    auto it = h->w.find(key);
    return it == h->w.end() ? nullptr : &it->second;
}

int need_w(dsb_handle* h, const std::string& key, std::initializer_list<int64_t> shape, const float** out) {
    const Weight* w = find_w(h, key);
    if (!w) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s'", key.c_str());
    if (w->shape != std::vector<int64_t>(shape)) return fail(h, DSB_ERR_WEIGHT, "weight '%s' has the wrong shape", key.c_str());
    *out = w->p;
    return 0;
}

const float* W(dsb_handle* h, const std::string& key) {
    const Weight* w = find_w(h, key);
    return w ? w->p : nullptr;
}

// ------------------------------------------------------------------------------------------ finalize helpers
// query / output projections folded into the K / V projection weights (DSB_FOLD_PROJ=0: the separate-GEMM form)
static const bool kFoldProj = [] { const char* e = getenv("DSB_FOLD_PROJ"); return !(e && e[0] == '0'); }();
// ... and their GEMM epilogues write the score / P.V operand layouts directly (DSB_KV_EPI=0: fp32 outputs + kv_pack kernel)
static const bool kKvEpi = kFoldProj && [] { const char* e = getenv("DSB_KV_EPI"); return !(e && e[0] == '0'); }();

int pack_gemm_weight(dsb_handle* h, const std::string& key, int N, int Cin, int taps, int f16 = 0) {
    const Weight* w = find_w(h, key);
    if (!w) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s'", key.c_str());
    if (w->numel != (long)N * Cin * taps) return fail(h, DSB_ERR_WEIGHT, "weight '%s' has the wrong size", key.c_str());
    bf16* dst = nullptr;
    if (int r = dev_alloc(h, &dst, (size_t)N * Cin * taps)) return r;
    if (int r = pack_weight_launch(w->p, N, Cin, taps, dst, 0, f16)) return fail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
    h->wpack[key] = dst;
    return 0;
}

int fold_bn(dsb_handle* h, const std::string& bn, const char* conv_bias_key, int C) {
    const float *w = W(h, bn + ".weight"), *b = W(h, bn + ".bias"), *m = W(h, bn + ".running_mean"),
                *v = W(h, bn + ".running_var");
    if (!w || !b || !m || !v) return fail(h, DSB_ERR_WEIGHT, "missing BatchNorm tensors of '%s'", bn.c_str());
    const float* cb = conv_bias_key ? W(h, conv_bias_key) : nullptr;
    if (conv_bias_key && !cb) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s'", conv_bias_key);
    float *sc = nullptr, *sh = nullptr;
    if (int r = dev_alloc(h, &sc, C)) return r;
    if (int r = dev_alloc(h, &sh, C)) return r;
    if (int r = bn_fold_launch(w, b, m, v, cb, C, 1e-5f, sc, sh, 0)) return fail(h, DSB_ERR_CUDA, "bn_fold launch %d", r);
    h->wf[bn + ".scale"] = sc;
    h->wf[bn + ".shift"] = sh;
    return 0;
}

int pack_dw(dsb_handle* h, const std::string& key, int C, int taps, int src_stride, int src_off) {
    const Weight* w = find_w(h, key);
    if (!w) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s'", key.c_str());
    if (w->numel != (long)C * src_stride) return fail(h, DSB_ERR_WEIGHT, "weight '%s' has the wrong size", key.c_str());
    float* dst = nullptr;
    if (int r = dev_alloc(h, &dst, (size_t)C * taps)) return r;
    if (int r = pack_dw_launch(w->p, C, taps, src_stride, src_off, dst, 0)) return fail(h, DSB_ERR_CUDA, "pack_dw launch %d", r);
    h->wf[key] = dst;
    return 0;
}

std::string stage_key(int i) { return "invpt_decoder.mid_stages." + std::to_string(i) + "."; }

// ------------------------------------------------------------------------------------------ workspace
int alloc_workspace(dsb_handle* h) {
    const size_t B = (size_t)h->cfg.max_batch;
    const size_t F = B * kT;
    const int enc_cout[3] = {192, 384, 768};
    for (int i = 0; i < 3; ++i)
        if (int r = dev_alloc(h, &h->tp[i], B * enc_cout[i])) return r;
    if (int r = dev_alloc(h, &h->h0, B * 5376 * 96)) return r;
    if (int r = dev_alloc(h, &h->gn_acc, 6 * B * 64 * 64)) return r;
    // encoder scratch: largest tensors are [B][5376][192]
    const size_t enc_max = B * 5376 * 192;
    if (int r = dev_alloc(h, &h->enc_act, enc_max)) return r;
    if (int r = dev_alloc(h, &h->enc_raw, enc_max)) return r;
    if (int r = dev_alloc(h, &h->enc_res, enc_max)) return r;
    if (int r = dev_alloc(h, &h->enc_c1, enc_max)) return r;
    if (int r = dev_alloc(h, &h->enc_sc, enc_max)) return r;
    for (int i = 0; i < 3; ++i)
        if (int r = dev_alloc(h, &h->enc_d[i], B * frame_elems(2 - i))) return r;
    for (int i = 0; i < 3; ++i)
        if (int r = dev_alloc(h, &h->back[i], F * frame_elems(i))) return r;
    if (h->cfg.audio_visual) {
        for (int i = 0; i < 4; ++i)
            if (int r = dev_alloc(h, &h->a_low[i], F * 84 * kStageC[i])) return r;
        for (int i = 0; i < 4; ++i)
            if (int r = dev_alloc(h, &h->a_cm[i], F * 84 * kStageC[i])) return r;
        if (int r = dev_alloc(h, &h->audio_tok, F * 84 * 512)) return r;
        if (int r = dev_alloc(h, &h->gate, B * kMaxFrame)) return r;
    }
    for (int i = 0; i < 4; ++i) {
        if (i > 0)
            if (int r = dev_alloc(h, &h->X[i], F * frame_elems(i))) return r;
        if (int r = dev_alloc(h, &h->X1[i], F * frame_elems(i))) return r;
        if (int r = dev_alloc(h, &h->X2[i], F * frame_elems(i))) return r;
        if (int r = dev_alloc(h, &h->r[i], B * kStageH[i] * kStageW[i] * 768)) return r;
    }
    if (int r = dev_alloc(h, &h->up, F * kMaxFrame * 2)) return r;
    if (int r = dev_alloc(h, &h->mid, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->q_ln, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->Qp, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->k_ln, F * 18 * 768)) return r;
    if (int r = dev_alloc(h, &h->v_ln, F * 18 * 768)) return r;
    if (int r = dev_alloc(h, &h->Kp, F * 18 * 1536)) return r;      // folded projections: [F*18][2C]
    if (int r = dev_alloc(h, &h->Vp, F * 18 * 1536)) return r;
    if (int r = dev_alloc(h, &h->KB, F * 48 * 768)) return r;
    if (int r = dev_alloc(h, &h->VB, F * 768 * 64)) return r;
    if (int r = dev_alloc(h, &h->P, F * 5376 * 64)) return r;
    if (int r = dev_alloc(h, &h->K1, F * 64 * 192)) return r;
    if (int r = dev_alloc(h, &h->V2, F * 192 * 64)) return r;
    if (int r = dev_alloc(h, &h->sbias, F * 64)) return r;
    for (int i = 0; i < 4; ++i) {
        const size_t C = kStageC[i], R = C <= 192 ? 64 : 48;
        if (int r = dev_alloc(h, &h->Ks[i], F * R * C)) return r;
        if (int r = dev_alloc(h, &h->Vs[i], F * C * 64)) return r;
        if (int r = dev_alloc(h, &h->sbs[i], F * R)) return r;
    }
    if (int r = dev_alloc(h, &h->o, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->ln2, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->hid, F * kMaxFrame * 2)) return r;
    if (int r = dev_alloc(h, &h->lnm, F * kMaxFrame)) return r;
    if (int r = dev_alloc(h, &h->lnstats, F * 5376)) return r;
    if (int r = dev_alloc(h, &h->S, B * 112 * 192 * 768)) return r;
    if (int r = dev_alloc(h, &h->p, B * 112 * 192)) return r;
    for (int i = 0; i < 8; ++i)
        if (int r = dev_alloc(h, &h->sbuf[i], B * kMapElems)) return r;
    if (int r = dev_alloc(h, &h->splitws, (size_t)kSplitWsPerClip * B)) return r;
    if (int r = dev_alloc(h, &h->splitws_head, (size_t)kSplitWsPerClip * B)) return r;
    h->t_all_cap = 1024;
    if (int r = dev_alloc(h, &h->t_all, (size_t)h->t_all_cap * B)) return r;
    if (int r = dev_alloc(h, &h->u8_stage, B * kMapElems)) return r;
    if (int r = dev_alloc(h, &h->hash_dev, 1)) return r;
    h->ws_ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ program building
struct Builder {
    dsb_handle* h;
    std::vector<Launch>* out;
    int err = 0;

    int cur = 0;                                  // stream index subsequent launches go to (0 = caller's stream)
    // Programmatic dependent launch is applied by ZONE: only in single-stream stretches of the program (nothing runs on the
    // side streams), where a kernel that is scheduled early and waits in pdl_wait() cannot take SMs away from another
    // stream.  Measured whole-program PDL was a loss; DSB_PDL_ZONES is a bit mask for experiments:
    // 1 encoder, 2 up-embedding, 4 token path after the K / V join, 8 multi-scale head.
    int zones = [] { const char* e = getenv("DSB_PDL_ZONES"); return e ? atoi(e) : 0; }();
    int zone = 0;                                 // PDL mode of the launches added from now on (0 or 2)
    void set_zone(int bit) { zone = (zones & bit) ? 2 : 0; }

    void meta(int kind, int ev) {
        h->prog_pdl.push_back(cur == 0 ? zone : 0);
        h->prog_kind.push_back(kind);
        h->prog_stream.push_back(cur);
        h->prog_ev.push_back(ev);
        if (kind == 0) h->prog_launches++;
    }
    void add(Launch l, const char* name = "misc", double bytes = 0.0) {
        out->push_back(std::move(l));
        h->prog_name.push_back(name);
        h->prog_flops.push_back(0.0);
        h->prog_bytes.push_back(bytes);
        meta(0, -1);
    }
    void sync_op(int kind, int ev, int stream) {
        const int keep = cur;
        cur = stream;
        out->push_back(Launch());
        h->prog_name.push_back(kind == 1 ? "record" : "wait");
        h->prog_flops.push_back(0.0);
        h->prog_bytes.push_back(0.0);
        meta(kind, ev);
        cur = keep;
    }
    // event `ev` marks "everything enqueued so far on `from`"; stream `to` will not run past it
    void depend(int ev, int from, int to) { sync_op(1, ev, from); sync_op(2, ev, to); }

    // algo_flops < 0: 2*M*N*K of the lowered GEMM (M = valid output pixels)
    void conv(ConvOp op, const char* name, double algo_flops = -1.0) {
        if (err) return;
        ConvLaunch cl;
        int r = conv_lower(op, &cl);
        if (r) { err = fail(h, DSB_ERR_CUDA, "conv_lower failed (%d) for %s (%dx%d, %d->%d)", r, name, op.H, op.W, op.Cin, op.N); return; }
        const int sms = h->num_sms;
        out->push_back([cl, sms](cudaStream_t s) { return conv_run(cl, sms, s); });
        if (cl.split.S > 1) h->prog_launches++;              // split-K ops launch the GEMM and the slice reduction
        h->prog_name.push_back(std::string("gemm:") + name);
        const double m = (double)op.F * op.H * op.W;
        h->prog_flops.push_back(algo_flops >= 0 ? algo_flops : 2.0 * m * op.N * (double)cl.p.taps * op.Cin);
        h->prog_bytes.push_back(0.0);
        meta(0, -1);
    }
};

ConvOp make_op(int kind, int F, int H, int W, int Cin, int N, const bf16* A, const bf16* Wt) {
    ConvOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.F = F; op.H = H; op.W = W; op.Cin = Cin; op.N = N; op.A = A; op.Wt = Wt;
    op.dilation = 1; op.T = 1; op.kt = 1;
    return op;
}

int build_program(dsb_handle* h) {
    h->prog.clear();
    h->prog_name.clear();
    h->prog_flops.clear();
    h->prog_bytes.clear();
    h->prog_kind.clear();
    h->prog_pdl.clear();
    h->prog_stream.clear();
    h->prog_ev.clear();
    h->prog_launches = 0;
    Builder b{h, &h->prog};
    const int B = h->B, F = B * kT;
    auto WP = [&](const std::string& k) -> const bf16* {
        auto it = h->wpack.find(k);
        return it == h->wpack.end() ? nullptr : it->second;
    };
    auto WF = [&](const std::string& k) -> const float* {
        auto it = h->wf.find(k);
        return it == h->wf.end() ? nullptr : it->second;
    };

    // ------------------------------------------------------------ noise encoder (sal_unet.py:279-300)
    static const bool gn_fused = [] { const char* e = getenv("DSB_GN_FUSED"); return !(e && e[0] == '0'); }();
    {
        TembWeights tw;
        tw.w0 = WF("temb.dense.0.weight.T"); tw.b0 = W(h, "temb.dense.0.bias");      // transposed [in][out]
        tw.w1 = WF("temb.dense.1.weight.T"); tw.b1 = W(h, "temb.dense.1.bias");
        const int cout[3] = {192, 384, 768};
        for (int i = 0; i < 3; ++i) {
            const std::string r = "res_encoder." + std::to_string(i) + ".0.temb_proj.";
            tw.wp[i] = WF(r + "weight.T"); tw.bp[i] = W(h, r + "bias"); tw.cout[i] = cout[i];
        }
        float* tp[3] = {h->tp[0], h->tp[1], h->tp[2]};
        // the timestep MLP only feeds the conv1 epilogues: it runs beside the stem / first GroupNorm on a side stream
        b.set_zone(1);
        b.depend(4, 0, 1);
        b.cur = 1;
        b.add([h, tw, tp, B](cudaStream_t s) { return temb_launch(h->cur_t, B, tw, tp, s); }, "temb");
        b.cur = 0;
        const float *w5 = WF("stem.w5"), *b5 = WF("stem.b5");
        float* h0 = h->h0;
        b.add([h, w5, b5, h0, B](cudaStream_t s) { return stem_launch(h->cur_x, B, w5, b5, h0, s); }, "stem5x5", (double)B * (344064.0 + 2064384.0));

        const float* cur = h->h0;
        int Cin = 96, H = 56, Wd = 96;
        for (int i = 0; i < 3; ++i) {
            const int Cout = cout[i], HW = H * Wd;
            const std::string rk = "res_encoder." + std::to_string(i) + ".0.";
            double* acc1 = h->gn_acc + (size_t)(2 * i) * h->cfg.max_batch * 64 * 64;
            double* acc2 = h->gn_acc + (size_t)(2 * i + 1) * h->cfg.max_batch * 64 * 64;
            const float *g1 = W(h, rk + "norm1.weight"), *b1 = W(h, rk + "norm1.bias");
            const float *g2 = W(h, rk + "norm2.weight"), *b2 = W(h, rk + "norm2.bias");
            bf16 *act = h->enc_act, *raw = h->enc_raw, *res = h->enc_res;
            float *c1 = h->enc_c1, *sc = h->enc_sc;
            if (gn_fused) {
                b.add([=](cudaStream_t s) { return gn_fused_launch(cur, B, HW, Cin, g1, b1, act, raw, s); }, "gn_fused", (double)B * HW * Cin * 8.0);
            } else {
                b.add([=](cudaStream_t s) { return gn_stats_launch(cur, B, HW, Cin, acc1, s); }, "gn_stats", (double)B * HW * Cin * 4.0);
                b.add([=](cudaStream_t s) { return gn_apply_launch(cur, B, HW, Cin, acc1, g1, b1, act, raw, s); }, "gn_apply", (double)B * HW * Cin * 8.0);
            }
            if (i == 0) b.depend(5, 1, 0);                  // timestep projections ready
            {   // 1x1 shortcut on the raw block input: only conv2's epilogue needs it, so it runs beside
                // conv1 / GroupNorm 2 on the side stream
                b.depend(6, 0, 1);
                b.cur = 1;
                ConvOp op = make_op(CONV_1X1, B, H, Wd, Cin, Cout, raw, WP(rk + "nin_shortcut.weight"));
                op.shift = W(h, rk + "nin_shortcut.bias"); op.out_f32 = sc;
                b.conv(op, "res.shortcut");
                b.cur = 0;
            }
            {   // conv1 + bias + temb projection
                ConvOp op = make_op(CONV_3X3, B, H, Wd, Cin, Cout, act, WP(rk + "conv1.weight"));
                op.shift = W(h, rk + "conv1.bias"); op.rowbias = h->tp[i]; op.out_f32 = c1;
                op.split_ws = h->splitws; op.split_ws_elems = kSplitWsPerClip * h->cfg.max_batch; op.split_frames_nominal = 8;
                op.halo = 1;
                b.conv(op, "res.conv1");
            }
            if (gn_fused) {
                b.add([=](cudaStream_t s) { return gn_fused_launch(c1, B, HW, Cout, g2, b2, act, nullptr, s); }, "gn_fused", (double)B * HW * Cout * 6.0);
            } else {
                b.add([=](cudaStream_t s) { return gn_stats_launch(c1, B, HW, Cout, acc2, s); }, "gn_stats", (double)B * HW * Cout * 4.0);
                b.add([=](cudaStream_t s) { return gn_apply_launch(c1, B, HW, Cout, acc2, g2, b2, act, nullptr, s); }, "gn_apply", (double)B * HW * Cout * 6.0);
            }
            b.depend(7, 1, 0);                              // shortcut ready
            {   // conv2 + bias + shortcut -> block output (only ever a GEMM operand: bf16)
                ConvOp op = make_op(CONV_3X3, B, H, Wd, Cout, Cout, act, WP(rk + "conv2.weight"));
                op.shift = W(h, rk + "conv2.bias"); op.residual = sc; op.out_bf16 = res;
                op.split_ws = h->splitws; op.split_ws_elems = kSplitWsPerClip * h->cfg.max_batch; op.split_frames_nominal = 8;
                op.halo = 1;
                b.conv(op, "res.conv2");
            }
            const std::string dk = "res_encoder." + std::to_string(i) + ".1.conv.";
            float* d = h->enc_d[i];
            {   // Downsample: pad (0,1,0,1) + 3x3 stride 2.  The result is the next block's input (compact copy) AND the
                // noise slice = frame 8 of the stage input / skip tensor (sal_unet.py:311-317), written by the same epilogue
                ConvOp op = make_op(CONV_3X3_S2, B, H / 2, Wd / 2, Cout, Cout, res, WP(dk + "weight"));
                op.shift = W(h, dk + "bias"); op.out_f32 = d;
                op.out2_f32 = h->back[2 - i]; op.out2_fmul = kT; op.out2_fadd = kTv;
                op.split_ws = h->splitws; op.split_ws_elems = kSplitWsPerClip * h->cfg.max_batch; op.split_frames_nominal = 8;
                b.conv(op, "res.down");
            }
            cur = d;
            Cin = Cout; H /= 2; Wd /= 2;
        }
    }

    // norm_mts on frames 0..4 only (the only ones ReduceTemp reads), then the (5,1,1) reduction + ReLU of stage j.
    // Nothing before the multi-scale sum reads r_j, so for stages 0..2 this pair can leave the critical path and run on the
    // head stream (3).  DSB_HEAD_STREAM: 0 = in line; 1 = forked right after the stage's MLP (beside the next up-embedding);
    // 2 = forked after the next stage's up-embedding (beside its memory-bound Q / K / V producers).
    static const int head_mode = [] { const char* e = getenv("DSB_HEAD_STREAM"); return e ? atoi(e) : 1; }();   // measured: 1 best (24.76 / 24.97 / 25.2 ms per step for 1 / 0 / 2)
    auto emit_head = [&](int j) {
        const int C = kStageC[j], H = kStageH[j], Wd = kStageW[j], HW = H * Wd;
        const long tokens = (long)F * HW;
        const std::string nk = "invpt_decoder.norm_mts." + std::to_string(j) + ".";
        const float *gm = W(h, nk + "weight"), *bm = W(h, nk + "bias");
        const float* x2 = h->X2[j];
        bf16* lnm = h->lnm;
        b.add([=](cudaStream_t s) { return ln_apply_launch(x2, tokens, C, gm, bm, lnm, HW, kT, kReduce, s, 1); }, "ln_apply_mts", (double)tokens * C * 6.0 * 5.0 / 9.0);
        ConvOp op = make_op(CONV_TEMPORAL, B, H, Wd, C, 768, h->lnm,
                            WP("invpt_decoder.redu_chan_up." + std::to_string(j) + ".proj.0.weight"));
        op.T = kT; op.kt = kReduce; op.act = ACT_RELU; op.out_f32 = h->r[j]; op.ab_f16 = 1;
        op.split_ws = head_mode ? h->splitws_head : h->splitws;
        op.split_ws_elems = kSplitWsPerClip * h->cfg.max_batch; op.split_frames_nominal = 8;
        b.conv(op, "reduce_temp");
    };

    // ------------------------------------------------------------ decoder stages (sal_unet.py:457-491)
    for (int i = 0; i < 4; ++i) {
        const int C = kStageC[i], H = kStageH[i], Wd = kStageW[i], HW = H * Wd, sk = kStageS[i];
        const std::string st = stage_key(i), bk = st + "blocks.0.";
        const float* Xi;
        b.set_zone(2);
        if (i == 0) {
            Xi = h->back[0];
        } else {
            const int Cp = kStageC[i - 1];
            const float* prev = h->X2[i - 1];
            bf16* up = h->up;
            b.add([=](cudaStream_t s) { return upsample2x_launch(prev, F, H / 2, Wd / 2, Cp, up, s); }, "upsample2x", (double)F * HW * Cp * (1.0 + 2.0));
            const std::string pe = st + "patch_embed.0.proj.";
            {
                ConvOp op = make_op(CONV_3X3, F, H, Wd, Cp, C, up, WP(pe + "1.weight"));
                op.dilation = 2; op.scale = WF(pe + "2.scale"); op.shift = WF(pe + "2.shift"); op.act = ACT_RELU;
                op.out_bf16 = h->mid;
                op.halo = 1;                                   // taken where the layer qualifies (N <= 128: the last stage)
                b.conv(op, "upembed.conv1");
            }
            {
                ConvOp op = make_op(CONV_3X3, F, H, Wd, C, C, h->mid, WP(pe + "4.weight"));
                op.dilation = 2; op.scale = WF(pe + "5.scale"); op.shift = WF(pe + "5.shift"); op.act = ACT_RELU;
                op.residual = (i == 1 || i == 2) ? h->back[i] : nullptr;   // stage 3 has no skip (transformer.py:265-270)
                op.out_f32 = h->X[i];
                op.halo = (i != 2);                            // stage 2 (two N tiles + fp32 residual epilogue) measured 8 % slower with it
                b.conv(op, "upembed.conv2");
            }
            Xi = h->X[i];
            if (head_mode == 2) {                               // the previous stage's ReduceTemp pair starts here
                b.depend(4, 0, 3);
                b.cur = 3;
                emit_head(i - 1);
                b.cur = 0;
            }
        }
        b.set_zone(0);                                      // Q / K / V producers: three streams
        const long tokens = (long)F * HW;
        // Stage 3 is the last stage and ReduceTemp reads frames 0..4 only (sal_unet.py:449-454,469-481): its
        // attention / MLP for frames 5..8 of every clip is dead work and is skipped (the gate still sees all 9).
        const int tmax = (i == 3) ? kReduce : kT;
        const bool remap = tmax < kT;
        const int Fu = B * tmax;
        const double live = (double)tmax / kT;
        auto token_op = [&](int Cin_, int N_, const bf16* A_, const bf16* Wt_) {
            // token-linear GEMM over the live frames (all tokens as one long row when nothing is skipped)
            ConvOp op = remap ? make_op(CONV_1X1, Fu, 1, HW, Cin_, N_, A_, Wt_) : make_op(CONV_1X1, 1, 1, (int)tokens, Cin_, N_, A_, Wt_);
            if (remap) { op.f_group = kT; op.f_used = tmax; op.out_remap = 1; }
            return op;
        };
        float2* stats = h->lnstats;
        const int kvR = C <= 192 ? 64 : 48;                 // key rows per frame of the score operand
        ScoreBias sbv = {nullptr, nullptr, nullptr, 0};
        if (kKvEpi) sbv = ScoreBias{WF(bk + "attn.fold_k.mb"), WF(bk + "attn.fold_k.cb"), h->sbs[i], kvR};
        // K / V projections; with the query / output projections folded into their weights (kFoldProj) N = 2C and the
        // output is already "keys seen through Wq" / "values seen through Wp" per head
        auto kv_proj = [&](const char* which, const bf16* A_, float* out_) {
            const std::string w = bk + (kFoldProj ? std::string("attn.fold_") + which : std::string("attn.proj_") + which);
            ConvOp op = make_op(CONV_1X1, 1, 1, F * 18, C, kFoldProj ? 2 * C : C, A_, WP(w + ".weight"));
            op.shift = kFoldProj ? WF(w + ".shift") : W(h, w + ".bias");
            op.out_f32 = out_;
            if (kKvEpi) {
                op.out_f32 = nullptr;
                op.out_bf16 = which[0] == 'k' ? h->Ks[i] : h->Vs[i];
                op.kv_mode = which[0] == 'k' ? 1 : 2; op.kv_R = kvR; op.kv_C = C;
            }
            b.conv(op, which[0] == 'k' ? "attn.proj_k" : "attn.proj_v", 2.0 * (double)F * 18 * C * C);
        };
        const float *ng = W(h, bk + "norm.weight"), *nb = W(h, bk + "norm.bias");
        bf16 *q_ln = h->q_ln, *k_ln = h->k_ln, *v_ln = h->v_ln;
        const float *wq = WF(bk + "attn.conv_proj_q.conv.weight"), *wk = WF(bk + "attn.conv_proj_k.conv.weight"),
                    *wv = WF(bk + "attn.conv_proj_v.conv.weight");
        const float *qg = W(h, bk + "attn.conv_proj_q.bn.weight"), *qb = W(h, bk + "attn.conv_proj_q.bn.bias");
        const float *kg = W(h, bk + "attn.conv_proj_k.bn.weight"), *kb = W(h, bk + "attn.conv_proj_k.bn.bias");
        const float *vg = W(h, bk + "attn.conv_proj_v.bn.weight"), *vb = W(h, bk + "attn.conv_proj_v.bn.bias");
        const QdwTables qtb = {WF(bk + "attn.conv_proj_q.wg"), WF(bk + "attn.conv_proj_q.wb"), WF(bk + "attn.conv_proj_q.wbs")};
        static const bool qv_fused_on = [] { const char* e = getenv("DSB_QV_FUSED"); return e && e[0] == '1'; }();   // measured: no faster than the three-kernel form (DESIGN.md negative results)
        const bool qv_fused = qv_fused_on && h->has_audio && C <= 192;
        if (qv_fused) {
            // narrow stages, audio-visual: Q and V come out of ONE pass over the stage input (LayerNorm statistics computed
            // in the kernel) on the caller's stream; the K branch (audio gate -> scramble -> pool) runs beside it on stream 2
            b.depend(0, 0, 2);                              // Xi is complete -> K branch may start
            b.cur = 2;
            {
                const float* al = h->a_low[i];
                const float* acm = h->a_cm[i];
                float* gate = h->gate;
                b.add([=](cudaStream_t s) { return av_gate_launch(Xi, al, B, kT, H, Wd, C, gate, s); }, "av_gate", (double)tokens * C * 4.0 + (double)B * HW * C * 4.0);
                b.add([=](cudaStream_t s) { return kpool_av_launch(gate, acm, B, kT, H, Wd, C, sk, wk, kg, kb, k_ln, tmax, s, sbv); }, "kpool_av", (double)B * HW * C * 4.0 * live);
                kv_proj("k", k_ln, h->Kp);
            }
            b.cur = 0;
            const float *wvg = WF(bk + "attn.conv_proj_v.wvg"), *wvbs = WF(bk + "attn.conv_proj_v.wvbs");
            b.add([=](cudaStream_t s) { return qv_tile_launch(Xi, F, H, Wd, C, sk, qtb, qg, qb, wvg, wvbs, vg, vb, q_ln, v_ln, kT, tmax, s); },
                  "qv_tile", (double)tokens * C * 6.0 * live);
            b.depend(1, 0, 1);                              // V tokens ready -> their projection runs on stream 1
            b.cur = 1;
            kv_proj("v", v_ln, h->Vp);
            b.cur = 0;
        } else {
        // three independent producers read the stage input: K (audio gate -> scramble -> pool, side stream 2),
        // V (side stream 1) and Q (caller's stream); they rejoin before the attention operands are built
        b.depend(0, 0, 2);                                  // Xi is complete -> K branch may start
        b.add([=](cudaStream_t s) { return ln_stats_launch(Xi, tokens, C, stats, HW, kT, tmax, s); }, "ln_stats", (double)tokens * C * 4.0 * live);
        b.depend(1, 0, 1);                                  // LayerNorm statistics ready -> V branch may start
        b.cur = 2;
        if (!h->has_audio) b.depend(1, 0, 2);               // visual-only K also needs the statistics
        if (h->has_audio) {
            const float* al = h->a_low[i];
            float* gate = h->gate;
            b.add([=](cudaStream_t s) { return av_gate_launch(Xi, al, B, kT, H, Wd, C, gate, s); }, "av_gate", (double)tokens * C * 4.0 + (double)B * HW * C * 4.0);
            const float* acm = h->a_cm[i];
            b.add([=](cudaStream_t s) { return kpool_av_launch(gate, acm, B, kT, H, Wd, C, sk, wk, kg, kb, k_ln, tmax, s, sbv); }, "kpool_av", (double)B * HW * C * 4.0 * live);
        } else {
            b.add([=](cudaStream_t s) { return pool_ln_launch(Xi, stats, F, H, Wd, C, sk, ng, nb, wk, kg, kb, k_ln, kT, tmax, s, sbv); }, "pool_ln_k", (double)tokens * C * 4.0 * live);
        }
        kv_proj("k", k_ln, h->Kp);
        b.cur = 1;
        b.add([=](cudaStream_t s) { return pool_ln_launch(Xi, stats, F, H, Wd, C, sk, ng, nb, wv, vg, vb, v_ln, kT, tmax, s); }, "pool_ln_v", (double)tokens * C * 4.0 * live);
        kv_proj("v", v_ln, h->Vp);
        b.cur = 0;
        b.add([=](cudaStream_t s) { return q_dwln_launch(Xi, stats, F, H, Wd, C, ng, nb, wq, &qtb, qg, qb, q_ln, kT, tmax, s); }, "q_dwln", (double)tokens * C * 6.0 * live);
        }
        // algorithmic FLOPs are always the reference's (all 9 frames), also where dead frames are skipped
        const bool fused_attn = C <= 192;
        if (fused_attn) {
            // narrow stages: Wq is folded into the 18 keys and Wp into the 18 values of every frame, then one chained
            // kernel does  q_ln . K'^T -> +bias -> per-head softmax -> P . V'' -> + proj bias + residual
            b.depend(2, 1, 0);                                  // join V
            b.depend(3, 2, 0);                                  // join K
            b.set_zone(4);                                      // from the K / V join to ReduceTemp: one stream
            const float *Kp = h->Kp, *Vp = h->Vp;
            const float *wqf = W(h, bk + "attn.proj_q.weight"), *bqf = W(h, bk + "attn.proj_q.bias");
            const float* wpT = WF(bk + "attn.proj.weight.T");
            bf16 *K1 = h->K1, *V2 = h->V2;
            float* sb = h->sbias;
            const float scale = 1.0f / sqrtf((float)C);
            if (kKvEpi) {
                // nothing to do: the projection epilogues wrote h->Ks / h->Vs, the K pooling kernel the score bias
            } else if (kFoldProj) {
                const float *mb = WF(bk + "attn.fold_k.mb"), *cb = WF(bk + "attn.fold_k.cb");
                b.add([=](cudaStream_t s) { return kv_pack_launch(Kp, Vp, k_ln, mb, cb, F, C, 64, kT, tmax, K1, sb, V2, s); }, "kv_pack");
            } else {
                b.add([=](cudaStream_t s) { return attn_fold_launch(Kp, Vp, wqf, bqf, wpT, F, C, scale, kT, tmax, K1, sb, V2, s); }, "attn_fold");
            }
            MlpOp mo;
            memset(&mo, 0, sizeof(mo));
            mo.mode = 1;
            mo.C = C; mo.HW = HW; mo.F = remap ? Fu : F;
            mo.f_group = remap ? kT : 0; mo.f_used = remap ? tmax : 0;
            mo.A = q_ln; mo.W1 = kKvEpi ? h->Ks[i] : h->K1; mo.W2 = kKvEpi ? h->Vs[i] : h->V2;
            mo.b1 = kKvEpi ? h->sbs[i] : h->sbias; mo.b2 = W(h, bk + "attn.proj.bias");
            mo.residual = Xi; mo.out = h->X1[i];
            MlpLaunch ml;
            if (int r = mlp_fused_lower(mo, &ml)) return fail(h, DSB_ERR_CUDA, "attention chain lower failed (%d)", r);
            const int sms = h->num_sms;
            b.add([ml, sms](cudaStream_t s) { return mlp_fused_run(ml, sms, s); }, "gemm:attn.fused");
            h->prog_flops.back() = 4.0 * (double)tokens * C * C + 4.0 * (double)tokens * 18 * C;   // q, proj, QK^T, PV
        } else if (kFoldProj) {
            // wide stages: the same folding, three launches -- pack, scores + softmax straight off the query tokens,
            // P . V'' with the projection bias and the residual in its epilogue
            b.depend(2, 1, 0);                                  // join V
            b.depend(3, 2, 0);                                  // join K
            b.set_zone(4);
            if (!kKvEpi) {
                const float *Kp = h->Kp, *Vp = h->Vp;
                const float *mb = WF(bk + "attn.fold_k.mb"), *cb = WF(bk + "attn.fold_k.cb");
                bf16 *KB = h->KB, *VB = h->VB;
                float* sb = h->sbias;
                b.add([=](cudaStream_t s) { return kv_pack_launch(Kp, Vp, k_ln, mb, cb, F, C, 48, kT, tmax, KB, sb, VB, s); }, "kv_pack");
            }
            {   // scores (query projection inside the keys) + per-head softmax over the 18 keys
                ConvOp op = make_op(CONV_1X1, remap ? Fu : F, 1, HW, C, 48, q_ln, kKvEpi ? h->Ks[i] : h->KB);
                op.b_rows_per_frame = 48; op.rowbias = kKvEpi ? h->sbs[i] : h->sbias; op.out_softmax = h->P;
                if (remap) { op.f_group = kT; op.f_used = tmax; op.out_remap = 1; }
                b.conv(op, "attn.qk_softmax", 2.0 * (double)tokens * C * C + 2.0 * (double)tokens * 18 * C);   // proj_q + bmm
            }
            {   // P . V'' (output projection inside the values) + bias + residual
                ConvOp op = make_op(CONV_1X1, remap ? Fu : F, 1, HW, 64, C, h->P, kKvEpi ? h->Vs[i] : h->VB);
                op.b_rows_per_frame = C; op.shift = W(h, bk + "attn.proj.bias"); op.residual = Xi; op.out_f32 = h->X1[i];
                if (remap) { op.f_group = kT; op.f_used = tmax; op.out_remap = 1; }
                b.conv(op, "attn.pv_proj", 2.0 * (double)tokens * 18 * C + 2.0 * (double)tokens * C * C);      // bmm + proj
            }
        } else {
        {
            ConvOp op = token_op(C, C, q_ln, WP(bk + "attn.proj_q.weight"));
            op.shift = W(h, bk + "attn.proj_q.bias"); op.out_bf16 = h->Qp;
            b.conv(op, "attn.proj_q", 2.0 * (double)tokens * C * C);
        }
        b.depend(2, 1, 0);                                  // join V
        b.depend(3, 2, 0);                                  // join K
        b.set_zone(4);                                      // from the K / V join to ReduceTemp: one stream
        {
            const float *Kp = h->Kp, *Vp = h->Vp;
            bf16 *KB = h->KB, *VB = h->VB;
            const float scale = 1.0f / sqrtf((float)C);          // dim_out ** -0.5 (attention.py:33)
            b.add([=](cudaStream_t s) { return attn_operands_launch(Kp, Vp, F, C, scale, KB, VB, s); }, "attn_operands");
        }
        {   // scores + per-head softmax over the 18 keys
            ConvOp op = make_op(CONV_1X1, remap ? Fu : F, 1, HW, C, 48, h->Qp, h->KB);
            op.b_rows_per_frame = 48; op.out_softmax = h->P;
            if (remap) { op.f_group = kT; op.f_used = tmax; op.out_remap = 1; }
            b.conv(op, "attn.qk_softmax", 2.0 * (double)tokens * 18 * C);   // reference bmm count (2 heads x C/2)
        }
        {   // P . V
            ConvOp op = make_op(CONV_1X1, remap ? Fu : F, 1, HW, 64, C, h->P, h->VB);
            op.b_rows_per_frame = C; op.out_bf16 = h->o;
            if (remap) { op.f_group = kT; op.f_used = tmax; op.out_remap = 1; }
            b.conv(op, "attn.pv", 2.0 * (double)tokens * 18 * C);
        }
        {   // output projection + residual
            ConvOp op = token_op(C, C, h->o, WP(bk + "attn.proj.weight"));
            op.shift = W(h, bk + "attn.proj.bias"); op.residual = Xi; op.out_f32 = h->X1[i];
            b.conv(op, "attn.proj", 2.0 * (double)tokens * C * C);
        }
        }
        {
            const float *g2 = W(h, bk + "norm2.weight"), *b2 = W(h, bk + "norm2.bias");
            const float* x1 = h->X1[i];
            bf16* ln2 = h->ln2;
            b.add([=](cudaStream_t s) { return ln_apply_launch(x1, tokens, C, g2, b2, ln2, HW, kT, tmax, s); }, "ln_apply", (double)tokens * C * 6.0 * live);
        }
        if (C <= 192) {
            // narrow stages: fc1 -> GELU -> fc2 -> +residual in one kernel, hidden activation kept on chip
            MlpOp mo;
            memset(&mo, 0, sizeof(mo));
            mo.C = C; mo.HW = remap ? HW : (int)tokens; mo.F = remap ? Fu : 1;
            mo.f_group = remap ? kT : 0; mo.f_used = remap ? tmax : 0;
            mo.A = h->ln2; mo.W1 = WP(bk + "mlp.fc1.weight"); mo.W2 = WP(bk + "mlp.fc2.weight");
            mo.b1 = W(h, bk + "mlp.fc1.bias"); mo.b2 = W(h, bk + "mlp.fc2.bias");
            mo.residual = h->X1[i]; mo.out = h->X2[i];
            MlpLaunch ml;
            if (int r = mlp_fused_lower(mo, &ml)) return fail(h, DSB_ERR_CUDA, "mlp_fused_lower failed (%d)", r);
            const int sms = h->num_sms;
            b.add([ml, sms](cudaStream_t s) { return mlp_fused_run(ml, sms, s); }, "gemm:mlp.fused");
            h->prog_flops.back() = 8.0 * (double)tokens * C * C;      // fc1 + fc2, reference count over all 9 frames
        } else {
            {
                ConvOp op = token_op(C, 2 * C, h->ln2, WP(bk + "mlp.fc1.weight"));
                op.shift = W(h, bk + "mlp.fc1.bias"); op.act = ACT_GELU; op.out_bf16 = h->hid;
                b.conv(op, "mlp.fc1", 4.0 * (double)tokens * C * C);
            }
            {
                ConvOp op = token_op(2 * C, C, h->hid, WP(bk + "mlp.fc2.weight"));
                op.shift = W(h, bk + "mlp.fc2.bias"); op.residual = h->X1[i]; op.out_f32 = h->X2[i];
                b.conv(op, "mlp.fc2", 4.0 * (double)tokens * C * C);
            }
        }
        if (head_mode == 0) emit_head(i);                       // in line on the caller's stream
        else if (head_mode == 1 || i == 3) {                    // forked as soon as X2_i exists (stage 3: nothing to hide under)
            if (i < 3) { b.depend(4, 0, 3); b.cur = 3; }
            emit_head(i);
            b.cur = 0;
        }
        if (head_mode && i == 3) b.depend(5, 3, 0);             // r_0..r_2 complete before the multi-scale sum
    }

    // ------------------------------------------------------------ multi-scale head (sal_unet.py:482-489,320-327)
    b.set_zone(8);
    {
        const float* rr[4] = {h->r[0], h->r[1], h->r[2], h->r[3]};
        bf16* S = h->S;
        b.add([=](cudaStream_t s) {
            const float* r4[4] = {rr[0], rr[1], rr[2], rr[3]};
            return ms_sum_launch(r4, B, S, s);
        }, "ms_sum", (double)B * (7140.0 * 768 * 4 + 21504.0 * 768 * 2));
        ConvOp op = make_op(CONV_3X3, B, 112, 192, 768, 96, h->S, WP("invpt_decoder.mt_proj.0.weight"));
        op.scale = WF("invpt_decoder.mt_proj.1.scale"); op.shift = WF("invpt_decoder.mt_proj.1.shift");
        op.act = ACT_RELU; op.head_w = W(h, "logits.linear_pred.weight"); op.ab_f16 = 1;
        float hb = 0.0f;
        cudaMemcpy(&hb, W(h, "logits.linear_pred.bias"), sizeof(float), cudaMemcpyDeviceToHost);
        op.head_b = hb; op.out_head = h->p;
        op.halo = 1;
        b.conv(op, "mt_proj_head");
        const float* p = h->p;
        b.add([h, p, B](cudaStream_t s) { return final_up_launch(p, B, h->cur_out, s); }, "final_up", (double)B * (86016.0 + 344064.0));
    }
    return b.err;
}

int run_program(dsb_handle* h, cudaStream_t s, bool serial = false) {
    // programmatic dependent launch (common.cuh): a launch that directly follows a cross-stream wait keeps a full
    // dependency (no PDL attribute); every other launch may be scheduled under the tail of its stream predecessor
    bool after_wait[4] = {false, false, false, false};
    for (size_t i = 0; i < h->prog.size(); ++i) {
        const int kind = h->prog_kind[i], si = h->prog_stream[i];
        cudaStream_t st = (serial || si == 0) ? s : h->side[si - 1];
        int r = 0;
        if (kind == 0) {
            pdl_allow_next() = !after_wait[si] || serial;
            after_wait[si] = false;
            const int keep = pdl_mode();
            if (h->prog_pdl[i] > keep) pdl_mode() = h->prog_pdl[i];
            r = h->prog[i](st);
            pdl_mode() = keep;
            pdl_allow_next() = true;
        }
        else if (serial) continue;
        else if (kind == 1) r = (int)cudaEventRecord(h->ev[h->prog_ev[i]], st);
        else { r = (int)cudaStreamWaitEvent(st, h->ev[h->prog_ev[i]], 0); after_wait[si] = true; }
        if (r) return fail(h, DSB_ERR_CUDA, "denoiser step %zu (%s) failed: %s (%d)", i, h->prog_name[i].c_str(),
                           r > 0 ? cudaGetErrorString((cudaError_t)r) : "launcher error", r);
    }
    h->last_launches += (int64_t)h->prog_launches;
    return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int dsb_create(const dsb_config* cfg, dsb_handle** out) {
    if (!cfg || !out) return DSB_ERR_ARG;
    if (cfg->max_batch < 1 || cfg->max_batch > 4096) return DSB_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DSB_ERR_CUDA;   // no CPU fallback
    dsb_handle* h = new dsb_handle();
    h->cfg = *cfg;
    cudaGetDevice(&h->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { delete h; return DSB_ERR_CUDA; }
    if (prop.major != 10) {
        delete h;
        return DSB_ERR_UNSUPPORTED;      // sm_100a only: tcgen05 / TMEM kernels
    }
    h->num_sms = prop.multiProcessorCount;
    if (gemm_init()) { delete h; return DSB_ERR_CUDA; }
    {
        // Block-scheduling priorities (DSB_PRIO=1, an experiment kept for the record): with the K branch -- the longest of the
        // three Q / K / V producers -- at the highest priority it finishes in half the time, the Q producers on the caller's
        // stream take correspondingly longer and the step time does not move (24.5 / 24.5 / 24.7 / 24.8 ms, on / off / on /
        // off): the phase is bound by the SM time of its kernels, not by their order.  Default: one priority for all.
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);         // lo = least (numerically greatest), hi = greatest priority
        const char* e = getenv("DSB_PRIO");
        const bool on = e && e[0] == '1';
        const int mid = (lo + hi) / 2;
        const int prio[3] = {on ? mid : lo, on ? hi : lo, lo};
        for (int i = 0; i < 3; ++i)
            if (cudaStreamCreateWithPriority(&h->side[i], cudaStreamNonBlocking, prio[i]) != cudaSuccess) { delete h; return DSB_ERR_CUDA; }
    }
    for (int i = 0; i < 8; ++i)
        if (cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return DSB_ERR_CUDA; }
    *out = h;
    return DSB_OK;
}

extern "C" void dsb_destroy(dsb_handle* h) {
    if (!h) return;
    for (auto& g : h->graphs)
        if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    for (void* p : h->allocs) cudaFree(p);
    if (h->noise_stage) cudaFree(h->noise_stage);
    for (int i = 0; i < 3; ++i)
        if (h->side[i]) cudaStreamDestroy(h->side[i]);
    for (int i = 0; i < 8; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
}

extern "C" const char* dsb_last_error(const dsb_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" size_t dsb_workspace_bytes(const dsb_handle* h, int B) {
    if (!h) return 0;
    if (h->ws_ready && B == h->cfg.max_batch) return h->alloc_bytes;
    // same formula as alloc_workspace, in bytes per clip (dominant terms) + weights
    const size_t F = (size_t)B * kT;
    size_t per = 0;
    per += F * (size_t)64512 * 4 * (7 + 14 + 15 + 15);     // back, X, X1, X2 over the stages
    per += F * kMaxFrame * 2 * (2 + 1 + 1 + 1 + 1 + 1 + 2 + 1);   // up, mid, q_ln, Qp, o, ln2, hid, lnm
    per += F * 5376 * 64 * 2 + F * 5376 * 8;
    per += (size_t)B * 112 * 192 * 768 * 2 + (size_t)B * 7140 * 768 * 4;
    per += (size_t)B * 5376 * 192 * (2 * 3 + 4 * 2) + (size_t)B * kMapElems * 4 * 8;
    return per + h->weight_bytes;
}

extern "C" int dsb_load_weight(dsb_handle* h, const char* ref_key, const void* data, const int64_t* shape, int ndim) {
    if (!h || !ref_key || !data || ndim < 0 || ndim > 8) return fail(h, DSB_ERR_ARG, "dsb_load_weight: bad argument");
    if (h->finalized) return fail(h, DSB_ERR_ARG, "dsb_load_weight after dsb_finalize_weights");
    Weight w;
    w.numel = 1;
    for (int i = 0; i < ndim; ++i) { w.shape.push_back(shape[i]); w.numel *= shape[i]; }
    if (w.numel < 1) return fail(h, DSB_ERR_ARG, "weight '%s' is empty", ref_key);
    if (int r = dev_alloc(h, &w.p, (size_t)w.numel)) return r;
    CUDA_TRY(h, cudaMemcpy(w.p, data, (size_t)w.numel * sizeof(float), cudaMemcpyDefault));
    h->w[ref_key] = w;
    return DSB_OK;
}

extern "C" int dsb_finalize_weights(dsb_handle* h) {
    if (!h) return DSB_ERR_ARG;
    if (h->finalized) return fail(h, DSB_ERR_ARG, "weights already finalized");
    const float* dummy;
    // ---- noise encoder
    if (int r = need_w(h, "conv_in.weight", {96, 1, 3, 3}, &dummy)) return r;
    if (int r = need_w(h, "down1.conv.weight", {96, 96, 3, 3}, &dummy)) return r;
    for (const char* k : {"conv_in.bias", "down1.conv.bias", "temb.dense.0.weight", "temb.dense.0.bias",
                          "temb.dense.1.weight", "temb.dense.1.bias", "logits.linear_pred.weight",
                          "logits.linear_pred.bias"})
        if (!W(h, k)) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s'", k);
    {
        float *w5 = nullptr, *b5 = nullptr;
        if (int r = dev_alloc(h, &w5, 25 * 96)) return r;
        if (int r = dev_alloc(h, &b5, 96)) return r;
        if (int r = stem_compose_launch(W(h, "conv_in.weight"), W(h, "conv_in.bias"), W(h, "down1.conv.weight"),
                                        W(h, "down1.conv.bias"), w5, b5, 0))
            return fail(h, DSB_ERR_CUDA, "stem_compose launch %d", r);
        h->wf["stem.w5"] = w5;
        h->wf["stem.b5"] = b5;
    }
    auto transposed = [&](const std::string& key, int R, int Cc) -> int {
        const Weight* w = find_w(h, key);
        if (!w || w->numel != (long)R * Cc) return fail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s'", key.c_str());
        float* dst = nullptr;
        if (int r = dev_alloc(h, &dst, (size_t)R * Cc)) return r;
        if (int r = transpose_launch(w->p, R, Cc, dst, 0)) return fail(h, DSB_ERR_CUDA, "transpose launch %d", r);
        h->wf[key + ".T"] = dst;
        return 0;
    };
    if (int r = transposed("temb.dense.0.weight", 384, 96)) return r;
    if (int r = transposed("temb.dense.1.weight", 384, 384)) return r;
    int cin = 96;
    const int cout[3] = {192, 384, 768};
    for (int i = 0; i < 3; ++i) {
        const std::string rk = "res_encoder." + std::to_string(i) + ".0.";
        if (int r = transposed(rk + "temb_proj.weight", cout[i], 384)) return r;
        for (const char* k : {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "conv1.bias", "conv2.bias",
                              "nin_shortcut.bias", "temb_proj.weight", "temb_proj.bias"})
            if (!W(h, rk + k)) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s%s'", rk.c_str(), k);
        if (int r = pack_gemm_weight(h, rk + "conv1.weight", cout[i], cin, 9)) return r;
        if (int r = pack_gemm_weight(h, rk + "conv2.weight", cout[i], cout[i], 9)) return r;
        if (int r = pack_gemm_weight(h, rk + "nin_shortcut.weight", cout[i], cin, 1)) return r;
        const std::string dk = "res_encoder." + std::to_string(i) + ".1.conv.";
        if (!W(h, dk + "bias")) return fail(h, DSB_ERR_WEIGHT, "missing weight '%sbias'", dk.c_str());
        if (int r = pack_gemm_weight(h, dk + "weight", cout[i], cout[i], 9)) return r;
        cin = cout[i];
    }
    // ---- decoder stages
    for (int i = 0; i < 4; ++i) {
        const int C = kStageC[i], sk = kStageS[i];
        const std::string st = stage_key(i), bk = st + "blocks.0.";
        if (i > 0) {
            const std::string pe = st + "patch_embed.0.proj.";
            if (int r = pack_gemm_weight(h, pe + "1.weight", C, kStageC[i - 1], 9)) return r;
            if (int r = pack_gemm_weight(h, pe + "4.weight", C, C, 9)) return r;
            if (int r = fold_bn(h, pe + "2", nullptr, C)) return r;
            if (int r = fold_bn(h, pe + "5", nullptr, C)) return r;
        }
        for (const char* k : {"norm.weight", "norm.bias", "norm2.weight", "norm2.bias", "attn.conv_proj_q.bn.weight",
                              "attn.conv_proj_q.bn.bias", "attn.conv_proj_k.bn.weight", "attn.conv_proj_k.bn.bias",
                              "attn.conv_proj_v.bn.weight", "attn.conv_proj_v.bn.bias", "attn.proj_q.bias",
                              "attn.proj_k.bias", "attn.proj_v.bias", "attn.proj.bias", "mlp.fc1.bias", "mlp.fc2.bias"})
            if (!W(h, bk + k)) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s%s'", bk.c_str(), k);
        for (const char* k : {"attn.proj_q.weight", "attn.proj_k.weight", "attn.proj_v.weight", "attn.proj.weight"})
            if (int r = pack_gemm_weight(h, bk + k, C, C, 1)) return r;
        if (C <= 192)
            if (int r = transposed(bk + "attn.proj.weight", C, C)) return r;
        if (kFoldProj) {
            // proj_q folded into proj_k, proj into proj_v (kernels.cu "projections folded into K / V"): fp32 products,
            // rounded to bf16 once
            float *MK = nullptr, *MV = nullptr, *cK = nullptr, *cV = nullptr, *mb = nullptr, *cb = nullptr;
            if (int r = dev_alloc(h, &MK, (size_t)2 * C * C)) return r;
            if (int r = dev_alloc(h, &MV, (size_t)2 * C * C)) return r;
            if (int r = dev_alloc(h, &cK, (size_t)2 * C)) return r;
            if (int r = dev_alloc(h, &cV, (size_t)2 * C)) return r;
            if (int r = dev_alloc(h, &mb, (size_t)2 * C)) return r;
            if (int r = dev_alloc(h, &cb, (size_t)2)) return r;
            if (int r = fold_weights_launch(W(h, bk + "attn.proj_q.weight"), W(h, bk + "attn.proj_q.bias"), W(h, bk + "attn.proj_k.weight"),
                                            W(h, bk + "attn.proj_k.bias"), W(h, bk + "attn.proj.weight"), W(h, bk + "attn.proj_v.weight"),
                                            W(h, bk + "attn.proj_v.bias"), C, 1.0f / sqrtf((float)C), MK, MV, cK, cV, mb, cb, 0))
                return fail(h, DSB_ERR_CUDA, "fold_weights launch %d", r);
            bf16 *pk = nullptr, *pv = nullptr;
            if (int r = dev_alloc(h, &pk, (size_t)2 * C * C)) return r;
            if (int r = dev_alloc(h, &pv, (size_t)2 * C * C)) return r;
            if (int r = pack_weight_launch(MK, 2 * C, C, 1, pk, 0)) return fail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
            if (int r = pack_weight_launch(MV, 2 * C, C, 1, pv, 0)) return fail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
            h->wpack[bk + "attn.fold_k.weight"] = pk;
            h->wpack[bk + "attn.fold_v.weight"] = pv;
            h->wf[bk + "attn.fold_k.shift"] = cK;
            h->wf[bk + "attn.fold_v.shift"] = cV;
            h->wf[bk + "attn.fold_k.mb"] = mb;
            h->wf[bk + "attn.fold_k.cb"] = cb;
        }
        if (int r = pack_gemm_weight(h, bk + "mlp.fc1.weight", 2 * C, C, 1)) return r;
        if (int r = pack_gemm_weight(h, bk + "mlp.fc2.weight", C, 2 * C, 1)) return r;
        // depthwise Conv3d(3,3,3) on a depth-1 volume: only the middle temporal tap touches data (attention.py:36-44)
        if (int r = pack_dw(h, bk + "attn.conv_proj_q.conv.weight", C, 9, 27, 9)) return r;
        {   // depthwise taps with the pre-attention LayerNorm affine folded in (q_dwln_tile2_kernel)
            float *wg = nullptr, *wb = nullptr, *wbs = nullptr;
            if (int r = dev_alloc(h, &wg, (size_t)9 * C)) return r;
            if (int r = dev_alloc(h, &wb, (size_t)9 * C)) return r;
            if (int r = dev_alloc(h, &wbs, (size_t)C)) return r;
            if (int r = q_dw_prep_launch(h->wf[bk + "attn.conv_proj_q.conv.weight"], W(h, bk + "norm.weight"), W(h, bk + "norm.bias"),
                                         C, wg, wb, wbs, 0))
                return fail(h, DSB_ERR_CUDA, "q_dw_prep launch %d", r);
            h->wf[bk + "attn.conv_proj_q.wg"] = wg;
            h->wf[bk + "attn.conv_proj_q.wb"] = wb;
            h->wf[bk + "attn.conv_proj_q.wbs"] = wbs;
        }
        if (int r = pack_dw(h, bk + "attn.conv_proj_k.conv.weight", C, sk * sk, sk * sk, 0)) return r;
        if (int r = pack_dw(h, bk + "attn.conv_proj_v.conv.weight", C, sk * sk, sk * sk, 0)) return r;
        {   // pooling taps of V with the pre-attention LayerNorm affine folded in (qv_tile_kernel)
            float *wvg = nullptr, *wvbs = nullptr;
            if (int r = dev_alloc(h, &wvg, (size_t)sk * sk * C)) return r;
            if (int r = dev_alloc(h, &wvbs, (size_t)C)) return r;
            if (int r = dw_affine_prep_launch(h->wf[bk + "attn.conv_proj_v.conv.weight"], W(h, bk + "norm.weight"),
                                              W(h, bk + "norm.bias"), sk * sk, C, wvg, wvbs, 0))
                return fail(h, DSB_ERR_CUDA, "dw_affine_prep launch %d", r);
            h->wf[bk + "attn.conv_proj_v.wvg"] = wvg;
            h->wf[bk + "attn.conv_proj_v.wvbs"] = wvbs;
        }
        if (h->cfg.audio_visual) {
            if (!W(h, bk + "align_conv.bias")) return fail(h, DSB_ERR_WEIGHT, "missing weight '%salign_conv.bias'", bk.c_str());
            if (int r = pack_gemm_weight(h, bk + "align_conv.weight", C, 512, 1)) return r;
        }
        const std::string nk = "invpt_decoder.norm_mts." + std::to_string(i) + ".";
        if (!W(h, nk + "weight") || !W(h, nk + "bias")) return fail(h, DSB_ERR_WEIGHT, "missing weight '%s*'", nk.c_str());
        // output-head chain (ReduceTemp, mt_proj): fp16 operands (precision policy)
        if (int r = pack_gemm_weight(h, "invpt_decoder.redu_chan_up." + std::to_string(i) + ".proj.0.weight", 768, C, kReduce, 1)) return r;
    }
    if (int r = pack_gemm_weight(h, "invpt_decoder.mt_proj.0.weight", 96, 768, 9, 1)) return r;
    if (int r = fold_bn(h, "invpt_decoder.mt_proj.1", "invpt_decoder.mt_proj.0.bias", 96)) return r;
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->weight_bytes = h->alloc_bytes;
    if (int r = alloc_workspace(h)) return r;
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->finalized = true;
    return DSB_OK;
}

extern "C" int dsb_set_condition(dsb_handle* h, const void* const feat[4], const void* audio, int B, void* stream) {
    if (!h || !feat) return DSB_ERR_ARG;
    if (!h->finalized) return fail(h, DSB_ERR_ARG, "dsb_set_condition before dsb_finalize_weights");
    if (B < 1 || B > h->cfg.max_batch) return fail(h, DSB_ERR_ARG, "batch %d outside [1, %d]", B, h->cfg.max_batch);
    if (audio && !h->cfg.audio_visual) return fail(h, DSB_ERR_UNSUPPORTED, "audio features given to a visual-only handle");
    for (int i = 0; i < 3; ++i)
        if (!feat[i]) return fail(h, DSB_ERR_ARG, "feat[%d] is null", i);
    cudaStream_t s = (cudaStream_t)stream;
    const bool rebuild = (B != h->B) || ((audio != nullptr) != h->has_audio) || h->prog.empty();
    h->B = B;
    h->has_audio = audio != nullptr;
    // frames 0..7 of the stage-0 input and of the two skip tensors (feat[3] is never read: sal_unet.py:311-317)
    for (int i = 0; i < 3; ++i) {
        int r = nct_to_frames_launch((const float*)feat[i], B, kStageC[i], kTv, kStageH[i] * kStageW[i], kT, h->back[i], s);
        if (r) return fail(h, DSB_ERR_CUDA, "nct_to_frames launch failed (%d)", r);
    }
    if (audio) {
        // loop-invariant align_conv of every stage (transformer.py:128-131), at 7x12
        int r = audio_tokens_launch((const float*)audio, B, kT, h->audio_tok, s);
        if (r) return fail(h, DSB_ERR_CUDA, "audio_tokens launch failed (%d)", r);
        for (int i = 0; i < 4; ++i) {
            const std::string bk = stage_key(i) + "blocks.0.";
            ConvOp op = make_op(CONV_1X1, 1, 1, B * kT * 84, 512, kStageC[i], h->audio_tok, h->wpack[bk + "align_conv.weight"]);
            op.shift = W(h, bk + "align_conv.bias"); op.out_f32 = h->a_low[i];
            ConvLaunch cl;
            if ((r = conv_lower(op, &cl))) return fail(h, DSB_ERR_CUDA, "conv_lower(align_conv) failed (%d)", r);
            if ((r = conv_run(cl, h->num_sms, s))) return fail(h, DSB_ERR_CUDA, "align_conv launch failed (%d)", r);
            if ((r = audio_cmajor_launch(h->a_low[i], B, kT, kStageC[i], h->a_cm[i], s)))
                return fail(h, DSB_ERR_CUDA, "audio_cmajor launch failed (%d)", r);
        }
    }
    if (rebuild) {
        for (auto& g : h->graphs)
            if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        h->graphs.clear();
        if (int r = build_program(h)) return r;
    }
    h->cond_epoch++;
    h->cond_launches = 3 + (audio ? 9 : 0);
    return DSB_OK;
}

extern "C" int dsb_denoise(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream) {
    if (!h || !x || !t || !out) return DSB_ERR_ARG;
    if (h->B == 0 || h->prog.empty()) return fail(h, DSB_ERR_ARG, "dsb_denoise before dsb_set_condition");
    if (B != h->B) return fail(h, DSB_ERR_ARG, "batch %d differs from the conditioned batch %d", B, h->B);
    h->cur_x = x; h->cur_t = t; h->cur_out = out;
    h->last_launches = 0;
    return run_program(h, (cudaStream_t)stream);
}

extern "C" int dsb_sampler_update(dsb_handle* h, const float* coef, const float* const* in, int nin,
                                  const float* noise, float noise_coef, float* out, int64_t n, void* stream) {
    if (!coef || !in || !out) return DSB_ERR_ARG;     // h may be NULL: the update needs no handle state
    if (nin < 1 || nin > 4 || n < 4 || (n & 3)) return fail(h, DSB_ERR_ARG, "dsb_sampler_update: nin in [1,4], n %% 4 == 0");
    const float* ins[4] = {nullptr, nullptr, nullptr, nullptr};
    float cs[4] = {0, 0, 0, 0};
    for (int k = 0; k < nin; ++k) { ins[k] = in[k]; cs[k] = coef[k]; }
    int r = axpy_launch(nin, ins, cs, noise, noise_coef, out, (long)n, (cudaStream_t)stream);
    if (r) return fail(h, DSB_ERR_CUDA, "axpy launch failed (%d)", r);
    return DSB_OK;
}

extern "C" int dsb_sampler_clamp(float* x, int64_t n, float lo, float hi, void* stream) {
    if (!x || n < 4 || (n & 3) || !(lo <= hi)) return DSB_ERR_ARG;
    return clamp_launch(x, (long)n, lo, hi, (cudaStream_t)stream) ? DSB_ERR_CUDA : DSB_OK;
}

extern "C" int dsb_sampler_adaptive_error(const float* x_lower, const float* x_higher, const float* x_prev, int B, int64_t n,
                                          float atol, float rtol, float* out, void* stream) {
    if (!x_lower || !x_higher || !x_prev || !out || B < 1 || n < 1 || n > (1 << 30)) return DSB_ERR_ARG;
    return adaptive_error_launch(x_lower, x_higher, x_prev, B, (int)n, atol, rtol, out, (cudaStream_t)stream) ? DSB_ERR_CUDA : DSB_OK;
}

extern "C" int dsb_sampler_dynamic_threshold(float* x, int B, int64_t n, int k, float w, float max_val, void* stream) {
    if (!x || B < 1 || n < 2 || n > (1 << 30)) return DSB_ERR_ARG;
    return dyn_threshold_launch(x, B, (int)n, k, w, max_val, (cudaStream_t)stream) ? DSB_ERR_ARG : DSB_OK;
}

static int enqueue_sampler(dsb_handle* h, const dsb_sampler_desc* d, int B, cudaStream_t s) {
    const long n = (long)B * kMapElems;
    int eval_idx = 0;
    for (int i = 0; i < d->n_ops; ++i) {
        const dsb_sampler_op& op = d->ops[i];
        if (op.kind == DSB_OP_EVAL) {
            if (op.src[0] < 0 || op.src[0] > 7 || op.src[0] == 1) return fail(h, DSB_ERR_ARG, "sampler op %d: bad EVAL source buffer", i);
            h->cur_x = h->sbuf[op.src[0]];
            h->cur_t = h->t_all + (size_t)eval_idx * B;
            h->cur_out = h->sbuf[1];
            ++eval_idx;
            if (int r = run_program(h, s)) return r;
        } else if (op.kind == DSB_OP_AXPY) {
            if (op.nin < 1 || op.nin > 4 || op.dst < 0 || op.dst > 7) return fail(h, DSB_ERR_ARG, "sampler op %d malformed", i);
            const float* ins[4] = {nullptr, nullptr, nullptr, nullptr};
            float cs[4] = {0, 0, 0, 0};
            for (int k = 0; k < op.nin; ++k) {
                if (op.src[k] < 0 || op.src[k] > 7) return fail(h, DSB_ERR_ARG, "sampler op %d: bad source buffer", i);
                ins[k] = h->sbuf[op.src[k]];
                cs[k] = op.coef[k];
            }
            const float* nz = (op.noise_index >= 0 && d->noise) ? h->noise_stage + (size_t)op.noise_index * n : nullptr;
            int r = axpy_launch(op.nin, ins, cs, nz, op.noise_coef, h->sbuf[op.dst], n, s);
            if (r) return fail(h, DSB_ERR_CUDA, "axpy launch failed (%d)", r);
            h->last_launches += 1;
        } else if (op.kind == DSB_OP_CLAMP) {
            if (op.dst < 0 || op.dst > 7) return fail(h, DSB_ERR_ARG, "sampler op %d malformed", i);
            if (int r = clamp_launch(h->sbuf[op.dst], n, op.coef[0], op.coef[1], s)) return fail(h, DSB_ERR_CUDA, "clamp launch failed (%d)", r);
            h->last_launches += 1;
        } else if (op.kind == DSB_OP_DYNTHRESH) {
            if (op.dst < 0 || op.dst > 7) return fail(h, DSB_ERR_ARG, "sampler op %d malformed", i);
            if (int r = dyn_threshold_launch(h->sbuf[op.dst], B, kMapElems, op.noise_index, op.coef[0], op.coef[1], s))
                return fail(h, DSB_ERR_ARG, "dynamic thresholding launch failed (%d)", r);
            h->last_launches += 1;
        } else if (op.kind == DSB_OP_POSTPROCESS) {
            if (op.dst < 0 || op.dst > 7) return fail(h, DSB_ERR_ARG, "sampler op %d malformed", i);
            if (!d->out_u8) return fail(h, DSB_ERR_ARG, "sampler op %d: DSB_OP_POSTPROCESS without desc.out_u8", i);
            if (int r = postprocess_launch(h->sbuf[op.dst], B, (int)kMapElems, h->sbuf[op.dst], h->u8_stage, s))
                return fail(h, DSB_ERR_CUDA, "postprocess launch failed (%d)", r);
            h->last_launches += 1;
        } else {
            return fail(h, DSB_ERR_ARG, "sampler op %d: unknown kind %d", i, op.kind);
        }
    }
    return DSB_OK;
}

extern "C" int dsb_sample(dsb_handle* h, const dsb_sampler_desc* d, float* x_inout, int B, void* stream) {
    if (!h || !d || !x_inout || !d->ops || d->n_ops < 1) return DSB_ERR_ARG;
    if (h->B == 0 || h->prog.empty()) return fail(h, DSB_ERR_ARG, "dsb_sample before dsb_set_condition");
    if (B != h->B) return fail(h, DSB_ERR_ARG, "batch %d differs from the conditioned batch %d", B, h->B);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)B * kMapElems;
    const size_t bytes = n * sizeof(float);
    // model times of all EVAL ops, replicated per clip; noise slabs / post-processing the program refers to
    std::vector<float> ts;
    int slabs = 0, evals = 0, others = 0;
    bool post = false;
    for (int i = 0; i < d->n_ops; ++i) {
        const dsb_sampler_op& op = d->ops[i];
        if (op.kind == DSB_OP_EVAL) {
            ++evals;
            for (int b = 0; b < B; ++b) ts.push_back(op.t);
        } else {
            ++others;
            if (op.kind == DSB_OP_AXPY && op.noise_index >= slabs) slabs = op.noise_index + 1;
            if (op.kind == DSB_OP_POSTPROCESS) post = true;
        }
    }
    if (evals > h->t_all_cap) return fail(h, DSB_ERR_UNSUPPORTED, "more than %d evaluations", h->t_all_cap);
    if (slabs > 0 && !d->noise) slabs = 0;
    h->last_launches = 0;
    if (ts != h->ts_uploaded) {
        // (rare: a new program) synchronous upload, so that the host vector may go away
        if (!ts.empty()) CUDA_TRY(h, cudaMemcpyAsync(h->t_all, ts.data(), ts.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
        h->ts_uploaded = ts;
    }
    if (slabs > 0) {
        // the caller's noise goes through a handle-owned buffer with a fixed address: a captured graph stays valid for
        // fresh randn tensors (eta > 0 DDIM, ancestral sampling allocate new noise on every call)
        const size_t slab_cap = (size_t)h->cfg.max_batch * kMapElems;
        if ((size_t)slabs > h->noise_slabs) {
            CUDA_TRY(h, cudaStreamSynchronize(s));
            for (auto& g : h->graphs)
                if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
            h->graphs.clear();
            if (h->noise_stage) cudaFree(h->noise_stage);
            h->noise_stage = nullptr;
            h->noise_slabs = 0;
            CUDA_TRY(h, cudaMalloc(&h->noise_stage, (size_t)slabs * slab_cap * sizeof(float)));
            h->noise_slabs = (size_t)slabs;
        }
        CUDA_TRY(h, cudaMemcpyAsync(h->noise_stage, d->noise, (size_t)slabs * bytes, cudaMemcpyDeviceToDevice, s));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->sbuf[0], x_inout, bytes, cudaMemcpyDeviceToDevice, s));
    int rc = DSB_OK;
    if (d->use_graph) {
        std::string key((const char*)d->ops, sizeof(dsb_sampler_op) * (size_t)d->n_ops);
        const int flags = (slabs > 0 ? 1 : 0) | (post ? 2 : 0);
        key.append((const char*)&flags, sizeof(flags));
        key.append((const char*)&B, sizeof(B));
        auto it = h->graphs.find(key);
        if (it == h->graphs.end()) {
            if (h->graphs.size() >= kMaxGraphs) {                     // evict the least recently used graph
                auto old = h->graphs.begin();
                for (auto g = h->graphs.begin(); g != h->graphs.end(); ++g)
                    if (g->second.last_use < old->second.last_use) old = g;
                CUDA_TRY(h, cudaStreamSynchronize(s));
                if (old->second.exec) cudaGraphExecDestroy(old->second.exec);
                h->graphs.erase(old);
            }
            cudaGraph_t graph = nullptr;
            cudaStream_t cs = nullptr;
            CUDA_TRY(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) { cudaStreamDestroy(cs); return fail(h, DSB_ERR_CUDA, "begin capture: %s", cudaGetErrorString(e)); }
            rc = enqueue_sampler(h, d, B, cs);
            e = cudaStreamEndCapture(cs, &graph);
            cudaStreamDestroy(cs);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(h, DSB_ERR_CUDA, "end capture: %s", cudaGetErrorString(e));
            SamplerGraph sg;
            e = cudaGraphInstantiate(&sg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(h, DSB_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
            it = h->graphs.emplace(key, sg).first;
        } else {
            // launches counted as if enqueued individually
            h->last_launches = (int64_t)evals * (int64_t)h->prog_launches + others;
        }
        it->second.last_use = ++h->graph_clock;
        CUDA_TRY(h, cudaGraphLaunch(it->second.exec, s));
    } else {
        rc = enqueue_sampler(h, d, B, s);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(x_inout, h->sbuf[0], bytes, cudaMemcpyDeviceToDevice, s));
    if (post) CUDA_TRY(h, cudaMemcpyAsync(d->out_u8, h->u8_stage, n, cudaMemcpyDeviceToDevice, s));
    return DSB_OK;
}

// 64-bit fingerprint of the conditioning tensors (same shapes as dsb_set_condition); synchronises `stream`.
extern "C" int dsb_condition_hash(dsb_handle* h, const void* const feat[4], const void* audio, int B, uint64_t* out,
                                  void* stream) {
    if (!h || !feat || !out) return DSB_ERR_ARG;
    if (!h->finalized) return fail(h, DSB_ERR_ARG, "dsb_condition_hash before dsb_finalize_weights");
    if (B < 1 || B > h->cfg.max_batch) return fail(h, DSB_ERR_ARG, "batch %d outside [1, %d]", B, h->cfg.max_batch);
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(h, cudaMemsetAsync(h->hash_dev, 0, sizeof(unsigned long long), s));
    for (int i = 0; i < 3; ++i) {
        if (!feat[i]) return fail(h, DSB_ERR_ARG, "feat[%d] is null", i);
        const size_t bytes = (size_t)B * kTv * frame_elems(i) * sizeof(float);
        if (int r = content_hash_launch(feat[i], bytes, 0x1000ull * (i + 1), h->hash_dev, s))
            return fail(h, DSB_ERR_ARG, "feat[%d] must be a 16-byte aligned contiguous fp32 tensor (%d)", i, r);
    }
    if (audio)
        if (int r = content_hash_launch(audio, (size_t)B * 512 * kT * 84 * sizeof(float), 0x5000ull, h->hash_dev, s))
            return fail(h, DSB_ERR_ARG, "audio must be a 16-byte aligned contiguous fp32 tensor (%d)", r);
    unsigned long long v = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&v, h->hash_dev, sizeof(v), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    *out = (uint64_t)v ^ ((uint64_t)B << 56) ^ (audio ? 0x8000000000ull : 0ull);
    return DSB_OK;
}

extern "C" int dsb_postprocess(const float* x, int B, int64_t pixels_per_map, float* clamped_or_null,
                               uint8_t* u8_or_null, void* stream) {
    if (!x || B < 1 || pixels_per_map < 1 || (!clamped_or_null && !u8_or_null)) return DSB_ERR_ARG;
    return postprocess_launch(x, B, (int)pixels_per_map, clamped_or_null, u8_or_null, (cudaStream_t)stream) ? DSB_ERR_CUDA : DSB_OK;
}

extern "C" int64_t dsb_last_launch_count(const dsb_handle* h) { return h ? h->last_launches : 0; }

extern "C" void dsb_set_pdl(int mode) { pdl_mode() = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
extern "C" int dsb_get_pdl(void) { return pdl_mode(); }

extern "C" int64_t dsb_debug_read(dsb_handle* h, const char* name, float* dst, int64_t max_elems, void* stream) {
    if (!h || !name || !dst || h->B == 0) return DSB_ERR_ARG;
    const std::string n(name);
    const float* src = nullptr;
    int64_t count = 0;
    const int B = h->B;
    if (n.size() == 2 && n[0] == 'x' && n[1] >= '0' && n[1] <= '3') { src = h->X2[n[1] - '0']; count = (int64_t)B * kT * (int64_t)frame_elems(n[1] - '0'); }
    else if (n.size() == 2 && n[0] == 'r' && n[1] >= '0' && n[1] <= '3') { const int i = n[1] - '0'; src = h->r[i]; count = (int64_t)B * kStageH[i] * kStageW[i] * 768; }
    else if (n.size() == 6 && n.compare(0, 5, "noise") == 0 && n[5] >= '0' && n[5] <= '2') { src = h->enc_d[2 - (n[5] - '0')]; count = (int64_t)B * (int64_t)frame_elems(n[5] - '0'); }
    else if (n == "p") { src = h->p; count = (int64_t)B * 112 * 192; }
    else if (n == "h0") { src = h->h0; count = (int64_t)B * 5376 * 96; }
    else if (n.size() == 3 && n.compare(0, 2, "tp") == 0 && n[2] >= '0' && n[2] <= '2') { const int c[3] = {192, 384, 768}; src = h->tp[n[2] - '0']; count = (int64_t)B * c[n[2] - '0']; }
    else if (n.size() == 2 && n[0] == 'a' && n[1] >= '0' && n[1] <= '3' && h->has_audio) { const int i = n[1] - '0'; src = h->a_low[i]; count = (int64_t)B * kT * 84 * kStageC[i]; }
    else return fail(h, DSB_ERR_ARG, "unknown debug buffer '%s'", name);
    if (count > max_elems) return fail(h, DSB_ERR_ARG, "debug buffer '%s' needs %lld elements", name, (long long)count);
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)count * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(h, DSB_ERR_CUDA, "debug copy failed: %s", cudaGetErrorString(e));
    return count;
}

// Times every launch of one denoiser evaluation with CUDA events on `stream` (the stream the kernels run on).
// ms / flops / bytes receive up to `cap` entries; returns the number of launches.
extern "C" int dsb_profile_denoise(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream,
                                   float* ms, double* flops, double* bytes, int cap) {
    if (!h || !x || !t || !out || !ms || !flops || !bytes) return DSB_ERR_ARG;
    if (h->B == 0 || h->prog.empty()) return fail(h, DSB_ERR_ARG, "dsb_profile_denoise before dsb_set_condition");
    if (B != h->B) return fail(h, DSB_ERR_ARG, "batch %d differs from the conditioned batch %d", B, h->B);
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int> idx;
    for (size_t i = 0; i < h->prog.size(); ++i)
        if (h->prog_kind[i] == 0) idx.push_back((int)i);
    const int n = (int)idx.size();
    if (cap < n) return fail(h, DSB_ERR_ARG, "profile buffers too small (%d < %d)", cap, n);
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CUDA_TRY(h, cudaEventCreate(&e));
    h->cur_x = x; h->cur_t = t; h->cur_out = out;
    int rc = 0;
    for (int k = 0; k < n && !rc; ++k) {          // everything serialised on `s`: per-launch times are exclusive
        cudaEventRecord(ev[k], s);
        int r = h->prog[idx[k]](s);
        if (r) rc = fail(h, DSB_ERR_CUDA, "launch %d (%s) failed: %d", k, h->prog_name[idx[k]].c_str(), r);
    }
    cudaEventRecord(ev[n], s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (!rc && e != cudaSuccess) rc = fail(h, DSB_ERR_CUDA, "profile sync: %s", cudaGetErrorString(e));
    for (int k = 0; k < n && !rc; ++k) {
        cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
        flops[k] = h->prog_flops[idx[k]];
        bytes[k] = h->prog_bytes[idx[k]];
    }
    for (auto& e2 : ev) cudaEventDestroy(e2);
    h->profile_idx = idx;
    return rc ? rc : n;
}

// Timeline of one evaluation as it really runs (three streams, event joins): an event pair around every launch on the
// stream it is enqueued on; start / end in ms relative to the first launch.  Shows the critical path and the idle gaps
// that the serialised profile above cannot.  Returns the number of launches.
extern "C" int dsb_profile_timeline(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream,
                                    float* start_ms, float* end_ms, int* stream_idx, int cap) {
    if (!h || !x || !t || !out || !start_ms || !end_ms || !stream_idx) return DSB_ERR_ARG;
    if (h->B == 0 || h->prog.empty()) return fail(h, DSB_ERR_ARG, "dsb_profile_timeline before dsb_set_condition");
    if (B != h->B) return fail(h, DSB_ERR_ARG, "batch %d differs from the conditioned batch %d", B, h->B);
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int> idx;
    for (size_t i = 0; i < h->prog.size(); ++i)
        if (h->prog_kind[i] == 0) idx.push_back((int)i);
    const int n = (int)idx.size();
    if (cap < n) return fail(h, DSB_ERR_ARG, "timeline buffers too small (%d < %d)", cap, n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (int k = 0; k < n; ++k) { CUDA_TRY(h, cudaEventCreate(&e0[k])); CUDA_TRY(h, cudaEventCreate(&e1[k])); }
    cudaEvent_t origin;
    CUDA_TRY(h, cudaEventCreate(&origin));
    h->cur_x = x; h->cur_t = t; h->cur_out = out;
    CUDA_TRY(h, cudaEventRecord(origin, s));
    int rc = 0, k = 0;
    const int keep = pdl_mode();
    pdl_mode() = 0;                                   // event records between launches would defeat it anyway
    for (size_t i = 0; i < h->prog.size() && !rc; ++i) {
        const int kind = h->prog_kind[i], si = h->prog_stream[i];
        cudaStream_t st = si == 0 ? s : h->side[si - 1];
        int r = 0;
        if (kind == 0) {
            cudaEventRecord(e0[k], st);
            r = h->prog[i](st);
            cudaEventRecord(e1[k], st);
            stream_idx[k] = si;
            ++k;
        } else if (kind == 1) r = (int)cudaEventRecord(h->ev[h->prog_ev[i]], st);
        else r = (int)cudaStreamWaitEvent(st, h->ev[h->prog_ev[i]], 0);
        if (r) rc = fail(h, DSB_ERR_CUDA, "timeline step %zu (%s) failed: %d", i, h->prog_name[i].c_str(), r);
    }
    pdl_mode() = keep;
    cudaError_t e = cudaDeviceSynchronize();
    if (!rc && e != cudaSuccess) rc = fail(h, DSB_ERR_CUDA, "timeline sync: %s", cudaGetErrorString(e));
    for (int j = 0; j < n && !rc; ++j) {
        cudaEventElapsedTime(&start_ms[j], origin, e0[j]);
        cudaEventElapsedTime(&end_ms[j], origin, e1[j]);
    }
    for (int j = 0; j < n; ++j) { cudaEventDestroy(e0[j]); cudaEventDestroy(e1[j]); }
    cudaEventDestroy(origin);
    h->profile_idx = idx;
    return rc ? rc : n;
}

extern "C" const char* dsb_profile_name(const dsb_handle* h, int i) {
    if (!h || i < 0 || i >= (int)h->profile_idx.size()) return "";
    return h->prog_name[h->profile_idx[i]].c_str();
}

extern "C" int64_t dsb_condition_launch_count(const dsb_handle* h) { return h ? h->cond_launches : 0; }
