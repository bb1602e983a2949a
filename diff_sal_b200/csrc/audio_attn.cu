// Once-per-clip audio transformer: B200 path for the reference's AudioAttnNet
// (models/audio_attention.py:93-143, configured by cfgs/audio_visual.py:34-48: depth 1, 2 heads x 64, dim 512, mlp 256).
//
//   tokens x [B*756][512] fp32 (channels-first input transposed once)
//   per layer:  LN -> qkv GEMM (512 -> 384, no bias) -> head split (q pre-scaled by 64^-0.5)
//               -> scores GEMM per (head, clip) [756 x 768] -> row softmax over 756 keys -> P.V GEMM per head
//               -> to_out GEMM (128 -> 512) + bias + residual
//               -> LN -> fc1 GEMM + exact-erf GELU -> fc2 GEMM + bias + residual
//   final LN, written back channels-first [B][512][9][7][12].
//
// Every contraction runs on the tcgen05 implicit-GEMM kernel of gemm_tc.cu (bf16 operands, fp32 accumulation and
// epilogue); the residual stream, LayerNorm statistics and the softmax are fp32.  The patch embedding / position
// embedding of the reference do not influence its output (audio_attention.py:134-141 rebinds x) and are not computed.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "conv_plan.cuh"
#include "diffsal_b200.h"
#include "kernels.cuh"
#include "weights.cuh"

using namespace dsb;

namespace {
constexpr int kTok = 9 * 7 * 12;     // 756 tokens per clip
constexpr int kTokPad = 768;         // key dimension padded to the GEMM's K block
constexpr int kDim = 512, kInner = 128, kMlp = 256, kHeads = 2, kHeadDim = 64;
}  // namespace

struct dsb_audio {
    int max_batch = 0;
    int depth = 0;
    int num_sms = 148;
    bool finalized = false;
    int launches = 0;
    std::string err;
    struct Wt { float* p; long numel; };
    std::map<std::string, Wt> w;
    std::map<std::string, bf16*> wp;
    std::vector<void*> allocs;
    float *xa = nullptr, *xb = nullptr, *x1 = nullptr, *sc = nullptr;
    bf16 *ln = nullptr, *qkv = nullptr, *Qh = nullptr, *Kh = nullptr, *Vt = nullptr, *P = nullptr, *O = nullptr, *hid = nullptr;
};

static int afail(dsb_audio* h, int code, const char* fmt, ...) {
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

template <class T>
static int aalloc(dsb_audio* h, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = ((count * sizeof(T) + 255) / 256) * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return afail(h, DSB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

extern "C" int dsb_audio_create(int max_batch, dsb_audio** out) {
    if (!out || max_batch < 1 || max_batch > 64) return DSB_ERR_ARG;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return DSB_ERR_CUDA;
    if (prop.major != 10) return DSB_ERR_UNSUPPORTED;             // sm_100a only, no fallback
    dsb_audio* h = new dsb_audio();
    h->max_batch = max_batch;
    h->num_sms = prop.multiProcessorCount;
    if (gemm_init()) { delete h; return DSB_ERR_CUDA; }
    *out = h;
    return DSB_OK;
}

extern "C" void dsb_audio_destroy(dsb_audio* h) {
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" const char* dsb_audio_last_error(const dsb_audio* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int dsb_audio_last_launch_count(const dsb_audio* h) { return h ? h->launches : 0; }

extern "C" int dsb_audio_load_weight(dsb_audio* h, const char* ref_key, const void* data, const int64_t* shape, int ndim) {
    if (!h || !ref_key || !data || ndim < 0 || ndim > 8) return afail(h, DSB_ERR_ARG, "dsb_audio_load_weight: bad argument");
    if (h->finalized) return afail(h, DSB_ERR_ARG, "dsb_audio_load_weight after dsb_audio_finalize");
    long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    if (numel < 1) return afail(h, DSB_ERR_ARG, "weight '%s' is empty", ref_key);
    float* p = nullptr;
    if (int r = aalloc(h, &p, (size_t)numel)) return r;
    if (cudaMemcpy(p, data, (size_t)numel * sizeof(float), cudaMemcpyDefault) != cudaSuccess)
        return afail(h, DSB_ERR_CUDA, "copy of weight '%s' failed", ref_key);
    h->w[ref_key] = {p, numel};
    return DSB_OK;
}

static const float* AW(dsb_audio* h, const std::string& k, long numel) {
    auto it = h->w.find(k);
    return (it == h->w.end() || it->second.numel != numel) ? nullptr : it->second.p;
}

static int apack(dsb_audio* h, const std::string& key, int N, int K) {
    const float* src = AW(h, key, (long)N * K);
    if (!src) return afail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s'", key.c_str());
    bf16* dst = nullptr;
    if (int r = aalloc(h, &dst, (size_t)N * K)) return r;
    if (int r = pack_weight_launch(src, N, K, 1, dst, 0)) return afail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
    h->wp[key] = dst;
    return 0;
}

extern "C" int dsb_audio_finalize(dsb_audio* h) {
    if (!h) return DSB_ERR_ARG;
    if (h->finalized) return afail(h, DSB_ERR_ARG, "weights already finalized");
    int depth = 0;
    while (h->w.count("transformer.layers." + std::to_string(depth) + ".0.to_qkv.weight")) ++depth;
    if (depth < 1) return afail(h, DSB_ERR_WEIGHT, "missing weight 'transformer.layers.0.0.to_qkv.weight'");
    h->depth = depth;
    if (!AW(h, "transformer.norm.weight", kDim) || !AW(h, "transformer.norm.bias", kDim))
        return afail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight 'transformer.norm.*'");
    for (int i = 0; i < depth; ++i) {
        const std::string a = "transformer.layers." + std::to_string(i) + ".0.", f = "transformer.layers." + std::to_string(i) + ".1.";
        const std::pair<const char*, long> vec[] = {{"norm.weight", kDim}, {"norm.bias", kDim}, {"to_out.0.bias", kDim}};
        for (auto& kv : vec)
            if (!AW(h, a + kv.first, kv.second)) return afail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s%s'", a.c_str(), kv.first);
        const std::pair<const char*, long> fvec[] = {{"net.0.weight", kDim}, {"net.0.bias", kDim}, {"net.1.bias", kMlp}, {"net.4.bias", kDim}};
        for (auto& kv : fvec)
            if (!AW(h, f + kv.first, kv.second)) return afail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s%s'", f.c_str(), kv.first);
        if (int r = apack(h, a + "to_qkv.weight", 3 * kInner, kDim)) return r;
        if (int r = apack(h, a + "to_out.0.weight", kDim, kInner)) return r;
        if (int r = apack(h, f + "net.1.weight", kMlp, kDim)) return r;
        if (int r = apack(h, f + "net.4.weight", kDim, kMlp)) return r;
    }
    const size_t M = (size_t)h->max_batch * kTok, FH = (size_t)kHeads * h->max_batch;
    if (int r = aalloc(h, &h->xa, M * kDim)) return r;
    if (int r = aalloc(h, &h->xb, M * kDim)) return r;
    if (int r = aalloc(h, &h->x1, M * kDim)) return r;
    if (int r = aalloc(h, &h->ln, M * kDim)) return r;
    if (int r = aalloc(h, &h->qkv, M * 3 * kInner)) return r;
    if (int r = aalloc(h, &h->Qh, FH * kTok * kHeadDim)) return r;
    if (int r = aalloc(h, &h->Kh, FH * kTokPad * kHeadDim)) return r;      // rows 756..767 stay zero
    if (int r = aalloc(h, &h->Vt, FH * kHeadDim * kTokPad)) return r;      // columns 756..767 stay zero
    if (int r = aalloc(h, &h->sc, FH * kTok * kTokPad)) return r;
    if (int r = aalloc(h, &h->P, FH * kTok * kTokPad)) return r;
    if (int r = aalloc(h, &h->O, M * kInner)) return r;
    if (int r = aalloc(h, &h->hid, M * kMlp)) return r;
    if (cudaDeviceSynchronize() != cudaSuccess) return afail(h, DSB_ERR_CUDA, "weight repack failed");
    h->finalized = true;
    return DSB_OK;
}

static ConvOp linear_op(int F, int rows, int K, int N, const bf16* A, const bf16* Wt) {
    ConvOp op;
    memset(&op, 0, sizeof(op));
    op.kind = CONV_1X1;
    op.F = F; op.H = 1; op.W = rows; op.Cin = K; op.N = N;
    op.dilation = 1;
    op.A = A; op.Wt = Wt;
    return op;
}

static int run_op(dsb_audio* h, const ConvOp& op, const char* what, cudaStream_t s) {
    ConvLaunch cl;
    if (int r = conv_lower(op, &cl)) return afail(h, DSB_ERR_CUDA, "%s: lowering failed (%d)", what, r);
    if (int r = conv_run(cl, h->num_sms, s)) return afail(h, DSB_ERR_CUDA, "%s: launch failed (%d)", what, r);
    ++h->launches;
    return 0;
}

#define AUDIO_TRY(expr, what)                                                                    \
    do {                                                                                         \
        if (int r_ = (expr)) return afail(h, DSB_ERR_CUDA, "%s launch failed (%d)", what, r_);   \
        ++h->launches;                                                                           \
    } while (0)

extern "C" int dsb_audio_forward(dsb_audio* h, const float* audio, float* out, int B, void* stream) {
    if (!h || !audio || !out) return DSB_ERR_ARG;
    if (!h->finalized) return afail(h, DSB_ERR_ARG, "dsb_audio_forward before dsb_audio_finalize");
    if (B < 1 || B > h->max_batch) return afail(h, DSB_ERR_ARG, "batch %d outside [1, %d]", B, h->max_batch);
    cudaStream_t s = (cudaStream_t)stream;
    h->launches = 0;
    const int M = B * kTok;
    // 'b c t h w -> b (t h w) c' (audio_attention.py:140)
    AUDIO_TRY(nct_to_frames_launch(audio, B, kDim, 9, 84, 9, h->xa, s), "token transpose");
    float *x = h->xa, *y = h->xb;
    for (int i = 0; i < h->depth; ++i) {
        const std::string a = "transformer.layers." + std::to_string(i) + ".0.", f = "transformer.layers." + std::to_string(i) + ".1.";
        // ---- Attention.forward (audio_attention.py:55-69)
        AUDIO_TRY(ln_apply_launch(x, M, kDim, AW(h, a + "norm.weight", kDim), AW(h, a + "norm.bias", kDim), h->ln, kTok, 1, 1, s), "attention LayerNorm");
        {
            ConvOp op = linear_op(1, M, kDim, 3 * kInner, h->ln, h->wp[a + "to_qkv.weight"]);
            op.out_bf16 = h->qkv;
            if (int r = run_op(h, op, "to_qkv", s)) return r;
        }
        AUDIO_TRY(qkv_split_launch(h->qkv, B, kTok, kTokPad, h->Qh, h->Kh, h->Vt, s), "qkv head split");
        {   // dots = (q * d^-0.5) k^T for every (head, clip): frames are head-major
            ConvOp op = linear_op(kHeads * B, kTok, kHeadDim, kTokPad, h->Qh, h->Kh);
            op.b_rows_per_frame = kTokPad;
            op.out_f32 = h->sc;
            if (int r = run_op(h, op, "attention scores", s)) return r;
        }
        AUDIO_TRY(softmax_rows_launch(h->sc, (long)kHeads * M, kTok, kTokPad, h->P, s), "attention softmax");
        for (int hd = 0; hd < kHeads; ++hd) {   // out[:, head*64 : head*64+64] = P_head . V_head
            ConvOp op = linear_op(B, kTok, kTokPad, kHeadDim, h->P + (size_t)hd * M * kTokPad, h->Vt + (size_t)hd * B * kHeadDim * kTokPad);
            op.b_rows_per_frame = kHeadDim;
            op.out_bf16 = h->O + hd * kHeadDim;
            op.ldo = kInner;
            if (int r = run_op(h, op, "attention P.V", s)) return r;
        }
        {
            ConvOp op = linear_op(1, M, kInner, kDim, h->O, h->wp[a + "to_out.0.weight"]);
            op.shift = AW(h, a + "to_out.0.bias", kDim);
            op.residual = x;
            op.out_f32 = h->x1;
            if (int r = run_op(h, op, "to_out", s)) return r;
        }
        // ---- FeedForward.forward (audio_attention.py:15-27)
        AUDIO_TRY(ln_apply_launch(h->x1, M, kDim, AW(h, f + "net.0.weight", kDim), AW(h, f + "net.0.bias", kDim), h->ln, kTok, 1, 1, s), "feed-forward LayerNorm");
        {
            ConvOp op = linear_op(1, M, kDim, kMlp, h->ln, h->wp[f + "net.1.weight"]);
            op.shift = AW(h, f + "net.1.bias", kMlp);
            op.act = ACT_GELU;
            op.out_bf16 = h->hid;
            if (int r = run_op(h, op, "feed-forward fc1", s)) return r;
        }
        {
            ConvOp op = linear_op(1, M, kMlp, kDim, h->hid, h->wp[f + "net.4.weight"]);
            op.shift = AW(h, f + "net.4.bias", kDim);
            op.residual = h->x1;
            op.out_f32 = y;
            if (int r = run_op(h, op, "feed-forward fc2", s)) return r;
        }
        float* t = x; x = y; y = t;
    }
    // Transformer.norm, then 'b (t h w) c -> b c t h w' (audio_attention.py:90,141)
    AUDIO_TRY(ln_nct_launch(x, B, kTok, kDim, AW(h, "transformer.norm.weight", kDim), AW(h, "transformer.norm.bias", kDim), out, s), "final LayerNorm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return afail(h, DSB_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
    return DSB_OK;
}
