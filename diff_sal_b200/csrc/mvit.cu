// Once-per-clip video encoder (video half): B200 path for the reference's MViTv2-S
// (models/mvit.py:796-1152, built by cfgs/audio_visual.py:27-32 as MViT(arch="small", out_scales=[0,1,2,3])).
//
//   video [B][3][16][224][384] -> Conv3d (3,7,7)/(2,4,4)/(1,3,3) patch embedding -> 1 + 8*56*96 tokens of 96 channels
//   16 MultiScaleBlocks (mvit.py:763-793), dims 96 / 192 / 384 / 768, heads 1 / 2 / 4 / 8 (head dim 96 throughout):
//       LN -> qkv GEMM -> per-head depthwise 3x3x3 pooling of q (stride 1 or 2) and k, v (stride 8 / 4 / 2 / 1) + LN(96)
//       -> scores GEMM per (head, clip) -> + decomposed relative-position bias, softmax -> P.V GEMM + residual pooling (+q)
//       -> proj GEMM + bias + skip (identity, or Linear(LN(x)) followed by a (1,3,3)/(1,2,2) max pool when the block
//          down-samples) -> LN -> fc1 GEMM + GELU -> fc2 GEMM + bias + residual
//   after blocks 0 / 2 / 13 / 15: x = LN_s(x) (the normalised tokens CONTINUE into the next block, mvit.py:1133) and the
//   patch tokens leave as [B][C][8][h][w]; the list is returned coarsest first (mvit.py:1152).
//
// Every contraction runs on the tcgen05 implicit-GEMM kernel of gemm_tc.cu with FP16 operands (GemmParams::ab_f16: every
// operand is a LayerNorm output, a GELU output, a probability or a weight, all far inside fp16's range; three more
// mantissa bits than bf16 at the same tensor-core rate keep the 16-block encoder within ~1e-3 of the fp32 reference), fp32
// accumulation and epilogue; the token stream, LayerNorm statistics, pooling, relative-position terms and the softmax
// are fp32.  Buffers typed `bf16*` below hold fp16 bits.
// Only the reference's geometry (16 x 224 x 384 clips) is accepted; anything else fails loudly.
#include <cuda_runtime.h>
#include <math.h>
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

constexpr int kHd = 96;                       // head dimension of every block
constexpr int kT = 8, kH0 = 56, kW0 = 96;     // token grid after the patch embedding
constexpr int kL0 = 1 + kT * kH0 * kW0;       // 43009
struct BlockCfg { int cin, cout, heads, sq, skv, rel; };
const BlockCfg kBlocks[16] = {{96, 96, 1, 1, 8, 111},  {96, 192, 2, 2, 4, 55},  {192, 192, 2, 1, 4, 55}, {192, 384, 4, 2, 2, 27},
                              {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27},
                              {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27},
                              {384, 384, 4, 1, 2, 27}, {384, 384, 4, 1, 2, 27}, {384, 768, 8, 2, 1, 27}, {768, 768, 8, 1, 1, 13}};
inline int stage_after(int i) { return i == 0 ? 0 : (i == 2 ? 1 : (i == 13 ? 2 : (i == 15 ? 3 : -1))); }
inline int pad64(int n) { return (n + 63) / 64 * 64; }
// fp16 bits in a 2-byte slot (the operand buffers are typed bf16* because the tensor maps only move 2-byte elements)
__device__ __forceinline__ bf16 h16(float v) {
    const __half t = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
    return __ushort_as_bfloat16(__half_as_ushort(t));
}

// ------------------------------------------------------------------------------------------ patch embedding
// Conv3d 3 -> 96, kernel (3,7,7), stride (2,4,4), padding (1,3,3) (mvit.py:975-981).  Block = (32 output columns, output
// row, clip x output frame); thread = output channel.  The 3 x 3 x 7 x 131 input patch is staged in shared memory.
__global__ void __launch_bounds__(96) mvit_patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ wT /*[441][96]*/,
                                                             const float* __restrict__ bias, float* __restrict__ tok) {
    __shared__ float patch[3 * 3 * 7][132];
    const int xo0 = blockIdx.x * 32, yo = blockIdx.y, bt = blockIdx.z;
    const int b = bt / kT, to = bt % kT, c = threadIdx.x;
    for (int i = threadIdx.x; i < 63 * 131; i += 96) {
        const int r = i / 131, col = i % 131;
        const int ci = r / 21, kt = (r / 7) % 3, ky = r % 7;
        const int ti = 2 * to + kt - 1, yi = 4 * yo + ky - 3, xi = 4 * xo0 + col - 3;
        float v = 0.0f;
        if (ti >= 0 && ti < 16 && yi >= 0 && yi < 224 && xi >= 0 && xi < 384)
            v = x[((((size_t)b * 3 + ci) * 16 + ti) * 224 + yi) * 384 + xi];
        patch[r][col] = v;
    }
    __syncthreads();
    float acc[32];
    const float bv = bias[c];
#pragma unroll
    for (int p = 0; p < 32; ++p) acc[p] = bv;
    for (int r = 0; r < 63; ++r) {
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            const float w = __ldg(wT + (r * 7 + kx) * 96 + c);
#pragma unroll
            for (int p = 0; p < 32; ++p) acc[p] = fmaf(w, patch[r][4 * p + kx], acc[p]);
        }
    }
    float* o = tok + ((size_t)b * kL0 + 1 + ((size_t)to * kH0 + yo) * kW0 + xo0) * 96 + c;
#pragma unroll
    for (int p = 0; p < 32; ++p) o[(size_t)p * 96] = acc[p];
}

__global__ void mvit_cls_kernel(const float* __restrict__ cls, float* __restrict__ tok, size_t row_stride) {
    tok[(size_t)blockIdx.x * row_stride + threadIdx.x] = cls[threadIdx.x];
}

// ------------------------------------------------------------------------------------------ q / k / v pooling
// attention_pool (mvit.py:459-510): per head, depthwise Conv3d 3x3x3 (padding 1, stride (1, s, s)) on the patch tokens, the
// cls token passes through, then LayerNorm(96).  One warp per output token (lane owns channels lane, lane+32, lane+64).
//   MODE 0 (q): qh[f][l][96] bf16 = LN(.) * 96^-0.5 (operand of the score GEMM), qres[f][l][96] fp32 = LN(.) (relative
//               position terms and residual pooling; the cls row is stored as 0 because the residual skips it, :639-643)
//   MODE 1 (k): kh[f][l][96] bf16, rows Lout .. Lpad-1 zero
//   MODE 2 (v): vt[f][96][Lpad] bf16 (transposed: B operand of the P.V GEMM), columns Lout .. Lpad-1 zero
// f = head * B + b (head-major, so that one head's clips are contiguous frames of a GEMM).
template <int MODE>
__global__ void __launch_bounds__(256) mvit_pool_kernel(const float* __restrict__ qkv, int B, int L, int ld, int col0, int heads,
                                                       int T, int H, int W, int s, const float* __restrict__ wT /*[27][96]*/,
                                                       const float* __restrict__ g, const float* __restrict__ bta, int Lout,
                                                       int Lpad, bf16* __restrict__ o16, float* __restrict__ o32) {
    const int lane = threadIdx.x & 31;
    const int l = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int f = blockIdx.y, head = f / B, b = f % B;
    if (l >= Lpad) return;
    if (l >= Lout) {                                   // zero padding of the key / value operands
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) o16[((size_t)f * Lpad + l) * 96 + lane + 32 * i] = __ushort_as_bfloat16((unsigned short)0);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 3; ++i) o16[((size_t)f * 96 + lane + 32 * i) * Lpad + l] = __ushort_as_bfloat16((unsigned short)0);
        }
        return;
    }
    const float* base = qkv + (size_t)b * L * ld + col0 + head * 96;
    float v[3];
    if (l == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = base[lane + 32 * i];
    } else {
        const int Ho = H / s, Wo = W / s;
        const int r = l - 1, to = r / (Ho * Wo), yo = (r / Wo) % Ho, xo = r % Wo;
        v[0] = v[1] = v[2] = 0.0f;
        for (int kt = 0; kt < 3; ++kt) {
            const int ti = to + kt - 1;
            if (ti < 0 || ti >= T) continue;
            for (int ky = 0; ky < 3; ++ky) {
                const int yi = yo * s + ky - 1;
                if (yi < 0 || yi >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xi = xo * s + kx - 1;
                    if (xi < 0 || xi >= W) continue;
                    const float* src = base + (size_t)(1 + (ti * H + yi) * W + xi) * ld;
                    const float* w = wT + ((kt * 3 + ky) * 3 + kx) * 96;
#pragma unroll
                    for (int i = 0; i < 3; ++i) v[i] = fmaf(__ldg(w + lane + 32 * i), src[lane + 32 * i], v[i]);
                }
            }
        }
    }
    const float mean = warp_sum(v[0] + v[1] + v[2]) * (1.0f / 96.0f);
    const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean;
    const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.0f / 96.0f) + 1e-5f);
    float y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = (v[i] - mean) * rstd * __ldg(g + lane + 32 * i) + __ldg(bta + lane + 32 * i);
    if (MODE == 0) {
        const float sc = 0.10206207261596577f;             // 96^-0.5 (mvit.py:562)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o16[((size_t)f * Lout + l) * 96 + lane + 32 * i] = h16(y[i] * sc);
            o32[((size_t)f * Lout + l) * 96 + lane + 32 * i] = l == 0 ? 0.0f : y[i];
        }
    } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i) o16[((size_t)f * Lpad + l) * 96 + lane + 32 * i] = h16(y[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) o16[((size_t)f * 96 + lane + 32 * i) * Lpad + l] = h16(y[i]);
    }
}

// ------------------------------------------------------------------------------------------ softmax with relative positions
// add_decomposed_rel_pos + softmax (mvit.py:366-404,633): for a patch query (t,h,w) and a patch key (kt,kh,kw)
//   score += q . Rt[t][kt] + q . Rh[h][kh] + q . Rw[w][kw]      (q = the pooled, normalised, UNscaled query)
// the cls query row and the cls key column get no bias.  One warp per query row; P is written as bf16 with the key
// padding zeroed.  Rt/Rh/Rw: [q_size][k_size][96] fp32, prepared at finalize time (resize_decomposed_rel_pos, :330-363).
__global__ void __launch_bounds__(256) mvit_softmax_kernel(const float* __restrict__ sc, const float* __restrict__ qres, int Lq,
                                                          int Lk, int Lpad, int qh, int qw, int kh, int kw,
                                                          const float* __restrict__ Rt, const float* __restrict__ Rh,
                                                          const float* __restrict__ Rw, bf16* __restrict__ P) {
    extern __shared__ uint32_t kidx[];                     // [Lk]: key k -> (kt | kh << 8 | kw << 16) offsets into rel[]; cls: none
    __shared__ float rel[8][64];                           // per warp: kt (8) | kh (<= 14) | kw (<= 24)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = blockIdx.x * 8 + wid, f = blockIdx.y;
    // the key -> (t, y, x) decode is the same for every query row: done once per block (8 rows), not per element
    for (int k = threadIdx.x; k < Lk; k += 256) {
        uint32_t v = 0xFFFFFFFFu;
        if (k > 0) {
            const int r = k - 1, t = r / (kh * kw), y = (r / kw) % kh, x = r % kw;
            v = (uint32_t)t | ((uint32_t)(kT + y) << 8) | ((uint32_t)(kT + kh + x) << 16);
        }
        kidx[k] = v;
    }
    __syncthreads();
    if (q >= Lq) return;
    const float* row = sc + ((size_t)f * Lq + q) * Lpad;
    bf16* prow = P + ((size_t)f * Lq + q) * Lpad;
    const int nrel = kT + kh + kw;
    const bool biased = q > 0;
    if (biased) {
        const int r = q - 1, t = r / (qh * qw), y = (r / qw) % qh, x = r % qw;
        const float4* qv = reinterpret_cast<const float4*>(qres + ((size_t)f * Lq + q) * 96);
        for (int j = lane; j < nrel; j += 32) {
            const float* R = j < kT ? Rt + ((size_t)t * kT + j) * 96
                                    : (j < kT + kh ? Rh + ((size_t)y * kh + (j - kT)) * 96 : Rw + ((size_t)x * kw + (j - kT - kh)) * 96);
            const float4* R4 = reinterpret_cast<const float4*>(R);
            float a = 0.0f;
#pragma unroll 6
            for (int c = 0; c < 24; ++c) {
                const float4 u = qv[c], w = __ldg(R4 + c);
                a = fmaf(u.x, w.x, a); a = fmaf(u.y, w.y, a); a = fmaf(u.z, w.z, a); a = fmaf(u.w, w.w, a);
            }
            rel[wid][j] = a;
        }
    }
    __syncwarp();
    const float* rl = rel[wid];
    // pass 1: online maximum / sum
    float mx = -INFINITY, sum = 0.0f;
    for (int k = lane; k < Lk; k += 32) {
        float v = row[k];
        const uint32_t id = kidx[k];
        if (biased && id != 0xFFFFFFFFu) v += rl[id & 255u] + rl[(id >> 8) & 255u] + rl[id >> 16];
        if (v > mx) { sum *= __expf(mx - v); mx = v; }
        sum += __expf(v - mx);
    }
    const float gmx = warp_max(mx);
    sum *= __expf(mx - gmx);                                // lanes without any element hold mx = -inf, sum = 0 -> 0 * 0
    const float inv = 1.0f / warp_sum(mx == -INFINITY ? 0.0f : sum);
    // pass 2: probabilities (the key padding is written as zeros)
    for (int k = lane; k < Lpad; k += 32) {
        float pv = 0.0f;
        if (k < Lk) {
            float v = row[k];
            const uint32_t id = kidx[k];
            if (biased && id != 0xFFFFFFFFu) v += rl[id & 255u] + rl[(id >> 8) & 255u] + rl[id >> 16];
            pv = __expf(v - gmx) * inv;
        }
        prow[k] = h16(pv);
    }
}

// ------------------------------------------------------------------------------------------ skip max pool
// pool_skip = MaxPool3d((1,3,3), (1,2,2), (0,1,1)) on the patch tokens, cls passes through (mvit.py:746-752,773-777)
__global__ void __launch_bounds__(256) mvit_maxpool_kernel(const float* __restrict__ in, int L, int C, int T, int H, int W,
                                                          float* __restrict__ out, int Lo) {
    const int cv_n = C >> 2;
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= (long)Lo * cv_n) return;
    const int l = (int)(i / cv_n), cv = (int)(i % cv_n);
    const float4* src = reinterpret_cast<const float4*>(in + (size_t)b * L * C) + cv;
    float4 m;
    if (l == 0) {
        m = src[0];
    } else {
        const int Ho = H / 2, Wo = W / 2;
        const int r = l - 1, t = r / (Ho * Wo), yo = (r / Wo) % Ho, xo = r % Wo;
        m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int ky = 0; ky < 3; ++ky) {
            const int yi = 2 * yo + ky - 1;
            if (yi < 0 || yi >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int xi = 2 * xo + kx - 1;
                if (xi < 0 || xi >= W) continue;
                const float4 v = src[(size_t)(1 + (t * H + yi) * W + xi) * cv_n];
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
    }
    reinterpret_cast<float4*>(out + (size_t)b * Lo * C)[(size_t)l * cv_n + cv] = m;
}

// ------------------------------------------------------------------------------------------ stage norm + feature output
// x <- LayerNorm_s(x) in place (fp32: the normalised tokens continue into the next block), one warp per token
__global__ void __launch_bounds__(256) mvit_ln_f32_kernel(float* __restrict__ x, long tokens, int C, const float* __restrict__ g,
                                                         const float* __restrict__ b) {
    const int lane = threadIdx.x & 31;
    const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tok >= tokens) return;
    float* row = x + tok * C;
    float v[24];
    const int n = C >> 5;                                  // 3, 6, 12 or 24 values per lane
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 24; ++i)
        if (i < n) { v[i] = row[lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 24; ++i)
        if (i < n) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
    for (int i = 0; i < 24; ++i)
        if (i < n) row[lane + 32 * i] = (v[i] - mean) * rstd * __ldg(g + lane + 32 * i) + __ldg(b + lane + 32 * i);
}

// tokens [B][1 + N][C] (cls skipped) -> features [B][C][N]   (32 x 32 shared-memory transpose)
__global__ void __launch_bounds__(256) mvit_tokens_to_nct_kernel(const float* __restrict__ tok, int N, int C, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int p = p0 + k, c = c0 + tx;
        tile[k][tx] = (p < N && c < C) ? tok[((size_t)b * (N + 1) + 1 + p) * C + c] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, p = p0 + tx;
        if (p < N && c < C) out[((size_t)b * C + c) * N + p] = tile[tx][k];
    }
}

// resize_decomposed_rel_pos (mvit.py:330-363) on the host: table [rel_len][96] -> R[q][k][96]
void build_rel_table(const std::vector<float>& rel, int rel_len, int q_size, int k_size, std::vector<float>* out) {
    const int max_rel = 2 * (q_size > k_size ? q_size : k_size) - 1;
    std::vector<float> resized((size_t)max_rel * kHd);
    if (rel_len != max_rel) {
        // F.interpolate(mode="linear", align_corners=False): src = scale * (dst + 0.5) - 0.5, clamped at 0
        const float scale = (float)rel_len / (float)max_rel;
        for (int j = 0; j < max_rel; ++j) {
            float src = scale * ((float)j + 0.5f) - 0.5f;
            if (src < 0.0f) src = 0.0f;
            const int i0 = (int)src;
            const int i1 = i0 + (i0 < rel_len - 1 ? 1 : 0);
            const float w1 = src - (float)i0, w0 = 1.0f - w1;
            for (int c = 0; c < kHd; ++c) resized[(size_t)j * kHd + c] = w0 * rel[(size_t)i0 * kHd + c] + w1 * rel[(size_t)i1 * kHd + c];
        }
    } else {
        resized = rel;
    }
    const double qr = (double)k_size / q_size > 1.0 ? (double)k_size / q_size : 1.0;
    const double kr = (double)q_size / k_size > 1.0 ? (double)q_size / k_size : 1.0;
    out->assign((size_t)q_size * k_size * kHd, 0.0f);
    for (int qi = 0; qi < q_size; ++qi)
        for (int ki = 0; ki < k_size; ++ki) {
            const float rc = ((float)qi * (float)qr - (float)ki * (float)kr) + (float)(k_size - 1) * (float)kr;
            const long idx = (long)rc;
            memcpy(out->data() + ((size_t)qi * k_size + ki) * kHd, resized.data() + (size_t)idx * kHd, kHd * sizeof(float));
        }
}

}  // namespace

struct dsb_mvit {
    int max_batch = 0;
    int num_sms = 148;
    bool finalized = false;
    int launches = 0;
    std::string err;
    struct Wt { float* p; long numel; };
    std::map<std::string, Wt> w;
    std::map<std::string, std::vector<float>> host_rel;     // rel_pos_* tables (host copies for the resize)
    std::map<std::string, bf16*> wp;                        // bf16 K-major GEMM weights
    std::map<std::string, float*> wf;                       // derived fp32 tables
    std::vector<void*> allocs;
    float *xa = nullptr, *xb = nullptr, *x1 = nullptr, *skp = nullptr, *skq = nullptr, *qkv = nullptr, *qres = nullptr, *sc = nullptr;
    bf16 *ln = nullptr, *Qh = nullptr, *Kh = nullptr, *Vt = nullptr, *P = nullptr, *O = nullptr, *hid = nullptr;
};

static int mfail(dsb_mvit* h, int code, const char* fmt, ...) {
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
static int malloc_dev(dsb_mvit* h, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = ((count * sizeof(T) + 255) / 256) * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return mfail(h, DSB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

extern "C" int dsb_mvit_create(int max_batch, dsb_mvit** out) {
    if (!out || max_batch < 1 || max_batch > 16) return DSB_ERR_ARG;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return DSB_ERR_CUDA;
    if (prop.major != 10) return DSB_ERR_UNSUPPORTED;             // sm_100a only, no fallback
    dsb_mvit* h = new dsb_mvit();
    h->max_batch = max_batch;
    h->num_sms = prop.multiProcessorCount;
    if (gemm_init()) { delete h; return DSB_ERR_CUDA; }
    *out = h;
    return DSB_OK;
}

extern "C" void dsb_mvit_destroy(dsb_mvit* h) {
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" const char* dsb_mvit_last_error(const dsb_mvit* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int dsb_mvit_last_launch_count(const dsb_mvit* h) { return h ? h->launches : 0; }

extern "C" int dsb_mvit_load_weight(dsb_mvit* h, const char* ref_key, const void* data, const int64_t* shape, int ndim) {
    if (!h || !ref_key || !data || ndim < 0 || ndim > 8) return mfail(h, DSB_ERR_ARG, "dsb_mvit_load_weight: bad argument");
    if (h->finalized) return mfail(h, DSB_ERR_ARG, "dsb_mvit_load_weight after dsb_mvit_finalize");
    long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    if (numel < 1) return mfail(h, DSB_ERR_ARG, "weight '%s' is empty", ref_key);
    float* p = nullptr;
    if (int r = malloc_dev(h, &p, (size_t)numel)) return r;
    if (cudaMemcpy(p, data, (size_t)numel * sizeof(float), cudaMemcpyDefault) != cudaSuccess)
        return mfail(h, DSB_ERR_CUDA, "copy of weight '%s' failed", ref_key);
    h->w[ref_key] = {p, numel};
    if (strstr(ref_key, "rel_pos_")) {
        std::vector<float> v((size_t)numel);
        if (cudaMemcpy(v.data(), p, (size_t)numel * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
            return mfail(h, DSB_ERR_CUDA, "copy of weight '%s' failed", ref_key);
        h->host_rel[ref_key] = std::move(v);
    }
    return DSB_OK;
}

static const float* MW(dsb_mvit* h, const std::string& k, long numel) {
    auto it = h->w.find(k);
    return (it == h->w.end() || it->second.numel != numel) ? nullptr : it->second.p;
}

static int mpack(dsb_mvit* h, const std::string& key, int N, int K) {
    const float* src = MW(h, key, (long)N * K);
    if (!src) return mfail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s'", key.c_str());
    bf16* dst = nullptr;
    if (int r = malloc_dev(h, &dst, (size_t)N * K)) return r;
    if (int r = pack_weight_launch(src, N, K, 1, dst, 0, 1)) return mfail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
    h->wp[key] = dst;
    return 0;
}

static int mtranspose(dsb_mvit* h, const std::string& key, int R, int Cc) {
    const float* src = MW(h, key, (long)R * Cc);
    if (!src) return mfail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s'", key.c_str());
    float* dst = nullptr;
    if (int r = malloc_dev(h, &dst, (size_t)R * Cc)) return r;
    if (int r = transpose_launch(src, R, Cc, dst, 0)) return mfail(h, DSB_ERR_CUDA, "transpose launch %d", r);
    h->wf[key + ".T"] = dst;
    return 0;
}

static int need_vec(dsb_mvit* h, const std::string& key, long n) {
    return MW(h, key, n) ? 0 : mfail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s'", key.c_str());
}

extern "C" int dsb_mvit_finalize(dsb_mvit* h) {
    if (!h) return DSB_ERR_ARG;
    if (h->finalized) return mfail(h, DSB_ERR_ARG, "weights already finalized");
    if (int r = need_vec(h, "cls_token", 96)) return r;
    if (int r = need_vec(h, "patch_embed.projection.bias", 96)) return r;
    if (int r = mtranspose(h, "patch_embed.projection.weight", 96, 441)) return r;
    int H = kH0, W = kW0;
    for (int i = 0; i < 16; ++i) {
        const BlockCfg& c = kBlocks[i];
        const std::string b = "blocks." + std::to_string(i) + ".", a = b + "attn.";
        for (const char* k : {"norm1.weight", "norm1.bias"})
            if (int r = need_vec(h, b + k, c.cin)) return r;
        for (const char* k : {"norm2.weight", "norm2.bias", "attn.proj.bias", "mlp.fc2.bias"})
            if (int r = need_vec(h, b + k, c.cout)) return r;
        if (int r = need_vec(h, a + "qkv.bias", 3 * c.cout)) return r;
        if (int r = need_vec(h, b + "mlp.fc1.bias", 4 * c.cout)) return r;
        if (int r = mpack(h, a + "qkv.weight", 3 * c.cout, c.cin)) return r;
        if (int r = mpack(h, a + "proj.weight", c.cout, c.cout)) return r;
        if (int r = mpack(h, b + "mlp.fc1.weight", 4 * c.cout, c.cout)) return r;
        if (int r = mpack(h, b + "mlp.fc2.weight", c.cout, 4 * c.cout)) return r;
        if (c.cin != c.cout) {
            if (int r = need_vec(h, b + "proj.bias", c.cout)) return r;
            if (int r = mpack(h, b + "proj.weight", c.cout, c.cin)) return r;
        }
        for (const char* n : {"q", "k", "v"}) {
            if (int r = mtranspose(h, a + "pool_" + n + ".weight", 96, 27)) return r;
            if (int r = need_vec(h, a + "norm_" + n + ".weight", 96)) return r;
            if (int r = need_vec(h, a + "norm_" + n + ".bias", 96)) return r;
        }
        // decomposed relative-position tables for this block's query / key grids
        const int qh = H / c.sq, qw = W / c.sq, kh = H / c.skv, kw = W / c.skv;
        const struct { const char* key; int len, q, k; const char* out; } rp[3] = {
            {"rel_pos_t", 15, kT, kT, ".Rt"}, {"rel_pos_h", c.rel, qh, kh, ".Rh"}, {"rel_pos_w", c.rel, qw, kw, ".Rw"}};
        for (auto& e : rp) {
            auto it = h->host_rel.find(a + e.key);
            if (it == h->host_rel.end() || (long)it->second.size() != (long)e.len * kHd)
                return mfail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s%s'", a.c_str(), e.key);
            std::vector<float> tab;
            build_rel_table(it->second, e.len, e.q, e.k, &tab);
            float* d = nullptr;
            if (int r = malloc_dev(h, &d, tab.size())) return r;
            if (cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
                return mfail(h, DSB_ERR_CUDA, "upload of a relative-position table failed");
            h->wf[a + e.out] = d;
        }
        H = qh; W = qw;
    }
    const int sc_[4] = {96, 192, 384, 768};
    for (int s = 0; s < 4; ++s) {
        if (int r = need_vec(h, "norm" + std::to_string(s) + ".weight", sc_[s])) return r;
        if (int r = need_vec(h, "norm" + std::to_string(s) + ".bias", sc_[s])) return r;
    }
    // ---- workspace (per clip maxima over the 16 blocks, see the size table in DESIGN.md)
    const size_t B = (size_t)h->max_batch;
    const size_t tokC = (size_t)kL0 * 96;                                   // largest [L][C] token tensor (block 0 / 1 input)
    if (int r = malloc_dev(h, &h->xa, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->xb, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->x1, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->skp, B * (size_t)kL0 * 192)) return r;    // Linear skip of block 1 at the input resolution
    if (int r = malloc_dev(h, &h->skq, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->qkv, B * (size_t)kL0 * 576)) return r;    // block 1: 43009 x 3*192
    if (int r = malloc_dev(h, &h->ln, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->Qh, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->qres, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->Kh, B * (size_t)8 * 2752 * 96)) return r; // block 14: 8 heads x 2752 padded keys
    if (int r = malloc_dev(h, &h->Vt, B * (size_t)8 * 2752 * 96)) return r;
    const size_t sc_max = (size_t)2 * 10753 * 2752;                         // block 1: 2 heads x 10753 queries x 2752 keys
    if (int r = malloc_dev(h, &h->sc, B * sc_max)) return r;
    if (int r = malloc_dev(h, &h->P, B * sc_max)) return r;
    if (int r = malloc_dev(h, &h->O, B * tokC)) return r;
    if (int r = malloc_dev(h, &h->hid, B * tokC * 4)) return r;             // block 0: 43009 x 384
    if (cudaDeviceSynchronize() != cudaSuccess) return mfail(h, DSB_ERR_CUDA, "weight preparation failed");
    h->finalized = true;
    return DSB_OK;
}

static ConvOp mlinear(int F, int rows, int K, int N, const bf16* A, const bf16* Wt) {
    ConvOp op;
    memset(&op, 0, sizeof(op));
    op.kind = CONV_1X1;
    op.F = F; op.H = 1; op.W = rows; op.Cin = K; op.N = N;
    op.dilation = 1;
    op.A = A; op.Wt = Wt;
    op.ab_f16 = 1;                                      // fp16 operands throughout the encoder
    op.out_f16 = 1;                                     // ... and every 16-bit output feeds such a GEMM
    return op;
}

static int mrun(dsb_mvit* h, const ConvOp& op, const char* what, int blk, cudaStream_t s) {
    ConvLaunch cl;
    if (int r = conv_lower(op, &cl)) return mfail(h, DSB_ERR_CUDA, "block %d %s: lowering failed (%d)", blk, what, r);
    if (int r = conv_run(cl, h->num_sms, s)) return mfail(h, DSB_ERR_CUDA, "block %d %s: launch failed (%d)", blk, what, r);
    ++h->launches;
    return 0;
}

#define MVIT_TRY(expr, what)                                                                                   \
    do {                                                                                                       \
        if (int r_ = (expr)) return mfail(h, DSB_ERR_CUDA, "%s launch failed (%d)", what, r_);               \
        ++h->launches;                                                                                         \
    } while (0)
#define MVIT_KERNEL(what)                                                                                      \
    do {                                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                                   \
        if (e_ != cudaSuccess) return mfail(h, DSB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_));         \
        ++h->launches;                                                                                         \
    } while (0)

extern "C" int dsb_mvit_forward(dsb_mvit* h, const float* video, float* const out[4], int B, void* stream) {
    if (!h || !video || !out || !out[0] || !out[1] || !out[2] || !out[3]) return DSB_ERR_ARG;
    if (!h->finalized) return mfail(h, DSB_ERR_ARG, "dsb_mvit_forward before dsb_mvit_finalize");
    if (B < 1 || B > h->max_batch) return mfail(h, DSB_ERR_ARG, "batch %d outside [1, %d]", B, h->max_batch);
    cudaStream_t s = (cudaStream_t)stream;
    h->launches = 0;
    // ---- patch embedding + cls token (mvit.py:1113-1116)
    mvit_patch_embed_kernel<<<dim3(kW0 / 32, kH0, B * kT), 96, 0, s>>>(video, h->wf["patch_embed.projection.weight.T"],
                                                                       MW(h, "patch_embed.projection.bias", 96), h->xa);
    MVIT_KERNEL("patch embedding");
    mvit_cls_kernel<<<B, 96, 0, s>>>(MW(h, "cls_token", 96), h->xa, (size_t)kL0 * 96);
    MVIT_KERNEL("cls token");
    float *x = h->xa, *y = h->xb;
    int H = kH0, W = kW0;
    for (int i = 0; i < 16; ++i) {
        const BlockCfg& c = kBlocks[i];
        const std::string b = "blocks." + std::to_string(i) + ".", a = b + "attn.";
        const int L = 1 + kT * H * W;
        const int qh = H / c.sq, qw = W / c.sq, kh = H / c.skv, kw = W / c.skv;
        const int Lq = 1 + kT * qh * qw, Lk = 1 + kT * kh * kw, Lpad = pad64(Lk);
        const int FH = c.heads * B;
        const long M = (long)B * L, Mq = (long)B * Lq;
        // ---- norm1 -> qkv (mvit.py:764, 606-609)
        MVIT_TRY(ln_apply_launch(x, M, c.cin, MW(h, b + "norm1.weight", c.cin), MW(h, b + "norm1.bias", c.cin), h->ln, L, 1, 1, s, 1), "norm1");
        {
            ConvOp op = mlinear(1, (int)M, c.cin, 3 * c.cout, h->ln, h->wp[a + "qkv.weight"]);
            op.shift = MW(h, a + "qkv.bias", 3 * c.cout);
            op.out_f32 = h->qkv;
            if (int r = mrun(h, op, "qkv", i, s)) return r;
        }
        // ---- pooled, normalised q / k / v per head (attention_pool)
        const int ld = 3 * c.cout;
        mvit_pool_kernel<0><<<dim3((Lq + 7) / 8, FH), 256, 0, s>>>(h->qkv, B, L, ld, 0, c.heads, kT, H, W, c.sq,
            h->wf[a + "pool_q.weight.T"], MW(h, a + "norm_q.weight", 96), MW(h, a + "norm_q.bias", 96), Lq, Lq, h->Qh, h->qres);
        MVIT_KERNEL("q pooling");
        mvit_pool_kernel<1><<<dim3((Lpad + 7) / 8, FH), 256, 0, s>>>(h->qkv, B, L, ld, c.cout, c.heads, kT, H, W, c.skv,
            h->wf[a + "pool_k.weight.T"], MW(h, a + "norm_k.weight", 96), MW(h, a + "norm_k.bias", 96), Lk, Lpad, h->Kh, nullptr);
        MVIT_KERNEL("k pooling");
        mvit_pool_kernel<2><<<dim3((Lpad + 7) / 8, FH), 256, 0, s>>>(h->qkv, B, L, ld, 2 * c.cout, c.heads, kT, H, W, c.skv,
            h->wf[a + "pool_v.weight.T"], MW(h, a + "norm_v.weight", 96), MW(h, a + "norm_v.bias", 96), Lk, Lpad, h->Vt, nullptr);
        MVIT_KERNEL("v pooling");
        // ---- scores = (q * d^-0.5) k^T per (head, clip)  (mvit.py:627)
        {
            ConvOp op = mlinear(FH, Lq, kHd, Lpad, h->Qh, h->Kh);
            op.b_rows_per_frame = Lpad;
            op.out_f32 = h->sc;
            if (int r = mrun(h, op, "attention scores", i, s)) return r;
        }
        mvit_softmax_kernel<<<dim3((Lq + 7) / 8, FH), 256, (size_t)Lk * sizeof(uint32_t), s>>>(h->sc, h->qres, Lq, Lk, Lpad, qh, qw, kh, kw, h->wf[a + ".Rt"],
                                                                   h->wf[a + ".Rh"], h->wf[a + ".Rw"], h->P);
        MVIT_KERNEL("relative-position softmax");
        // ---- out[:, head] = P_head . V_head + q_head (residual pooling, cls row excluded)  (mvit.py:634-646)
        for (int hd = 0; hd < c.heads; ++hd) {
            ConvOp op = mlinear(B, Lq, Lpad, kHd, h->P + (size_t)hd * B * Lq * Lpad, h->Vt + (size_t)hd * B * kHd * Lpad);
            op.b_rows_per_frame = kHd;
            op.residual = h->qres + (size_t)hd * B * Lq * kHd;
            op.out_bf16 = h->O + hd * kHd;
            op.ldo = c.cout;
            if (int r = mrun(h, op, "attention P.V", i, s)) return r;
        }
        // ---- skip path (mvit.py:767-777): identity, or Linear(norm1(x)) followed by the (1,3,3)/(1,2,2) max pool
        const float* skip = x;
        if (c.cin != c.cout) {
            ConvOp op = mlinear(1, (int)M, c.cin, c.cout, h->ln, h->wp[b + "proj.weight"]);
            op.shift = MW(h, b + "proj.bias", c.cout);
            op.out_f32 = h->skp;
            if (int r = mrun(h, op, "skip projection", i, s)) return r;
            skip = h->skp;
        }
        if (c.sq > 1) {
            const long n = (long)Lq * (c.cout / 4);
            mvit_maxpool_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, s>>>(skip, L, c.cout, kT, H, W, h->skq, Lq);
            MVIT_KERNEL("skip max pool");
            skip = h->skq;
        }
        // ---- x = skip + proj(attn)  (mvit.py:647, 779)
        {
            ConvOp op = mlinear(1, (int)Mq, c.cout, c.cout, h->O, h->wp[a + "proj.weight"]);
            op.shift = MW(h, a + "proj.bias", c.cout);
            op.residual = skip;
            op.out_f32 = h->x1;
            if (int r = mrun(h, op, "attention projection", i, s)) return r;
        }
        // ---- MLP (mvit.py:780-789)
        MVIT_TRY(ln_apply_launch(h->x1, Mq, c.cout, MW(h, b + "norm2.weight", c.cout), MW(h, b + "norm2.bias", c.cout), h->ln, Lq, 1, 1, s, 1), "norm2");
        {
            ConvOp op = mlinear(1, (int)Mq, c.cout, 4 * c.cout, h->ln, h->wp[b + "mlp.fc1.weight"]);
            op.shift = MW(h, b + "mlp.fc1.bias", 4 * c.cout);
            op.act = ACT_GELU;
            op.out_bf16 = h->hid;
            if (int r = mrun(h, op, "mlp.fc1", i, s)) return r;
        }
        {
            ConvOp op = mlinear(1, (int)Mq, 4 * c.cout, c.cout, h->hid, h->wp[b + "mlp.fc2.weight"]);
            op.shift = MW(h, b + "mlp.fc2.bias", c.cout);
            op.residual = h->x1;
            op.out_f32 = y;
            if (int r = mrun(h, op, "mlp.fc2", i, s)) return r;
        }
        float* t = x; x = y; y = t;
        H = qh; W = qw;
        // ---- output scale: x = norm_s(x) (kept for the next block), patch tokens -> [B][C][8][h][w]  (mvit.py:1129-1147)
        const int st = stage_after(i);
        if (st >= 0) {
            const std::string nk = "norm" + std::to_string(st);
            mvit_ln_f32_kernel<<<(unsigned)((Mq + 7) / 8), 256, 0, s>>>(x, Mq, c.cout, MW(h, nk + ".weight", c.cout), MW(h, nk + ".bias", c.cout));
            MVIT_KERNEL("stage norm");
            const int N = kT * H * W;
            mvit_tokens_to_nct_kernel<<<dim3((N + 31) / 32, (c.cout + 31) / 32, B), 256, 0, s>>>(x, N, c.cout, out[3 - st]);
            MVIT_KERNEL("feature output");
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return mfail(h, DSB_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
    return DSB_OK;
}
