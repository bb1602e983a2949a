// VGGish feature stack on the B200 path (audio half): the reference's VGGish.forward_feat
// (models/vggish.py:87-103) as VideoSaliencyModel.forward_vggish calls it (models/diff_model.py:70-76) on
// audio.view(-1, 1, 112, 192): six 3x3 convolutions + ReLU with four 2x2 max pools, 50 GFLOP per clip, once per clip.
//
//   conv 1->64 + ReLU + pool : one direct kernel (K = 9 is no GEMM), bf16 channels-last [F,56,96,64]
//   conv 64->128 ... 512->512: the tcgen05 implicit-GEMM kernel (bias + ReLU epilogue, bf16 channels-last out)
//   pools                    : bf16 channels-last 2x2 max; the last one writes fp32 [F,512,7,12] (the reference layout)
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "conv_plan.cuh"
#include "diffsal_b200.h"
#include "weights.cuh"

using namespace dsb;

namespace {

constexpr int kH = 112, kW = 192;

// conv3x3(1 -> 64, pad 1) + bias + ReLU + maxpool 2x2 -> bf16 [F][56][96][64].  Block = one pooled row of one frame;
// thread = (pooled x, 8-channel group): the 4x4 input patch is shared by the 2x2 conv outputs under the pool window.
__global__ void __launch_bounds__(256) vgg_conv0_pool_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ out) {
    __shared__ float rows[4][kW + 2];
    __shared__ float sw[9][64];
    __shared__ float sb[64];
    const int py = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const float* xf = x + (size_t)f * kH * kW;
    for (int i = tid; i < 4 * (kW + 2); i += 256) {
        const int r = i / (kW + 2), c = i % (kW + 2);
        const int yy = 2 * py - 1 + r, xx = c - 1;
        rows[r][c] = (yy >= 0 && yy < kH && xx >= 0 && xx < kW) ? xf[yy * kW + xx] : 0.0f;
    }
    for (int i = tid; i < 9 * 64; i += 256) sw[i / 64][i % 64] = w[(i % 64) * 9 + i / 64];      // [tap][cout]
    if (tid < 64) sb[tid] = bias[tid];
    __syncthreads();
    for (int item = tid; item < 96 * 8; item += 256) {
        const int px = item >> 3, c0 = (item & 7) * 8;
        float p[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) p[r][c] = rows[r][2 * px + c];
        uint32_t packed[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float best[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int co = c0 + j + u;
                float m = 0.0f;                                  // ReLU floor: max(relu(a), ...) = max(0, a, ...)
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        float a = sb[co];
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) a = fmaf(sw[ky * 3 + kx][co], p[dy + ky][dx + kx], a);
                        m = fmaxf(m, a);
                    }
                best[u] = m;
            }
            packed[j >> 1] = pack_bf16x2(best[0], best[1]);
        }
        *reinterpret_cast<uint4*>(out + (((size_t)f * 56 + py) * 96 + px) * 64 + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

// 2x2 max pool, bf16 channels-last [F][2H][2W][C] -> [F][H][W][C]; thread = 8 channels of one output pixel
__global__ void __launch_bounds__(256) vgg_pool_kernel(const bf16* __restrict__ in, int H, int W, int C, long total,
                                                      bf16* __restrict__ out) {
    const int cv = C >> 3;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int c8 = (int)(i % cv);
        const long pix = i / cv;
        const int xo = (int)(pix % W), yo = (int)((pix / W) % H);
        const long f = pix / ((long)W * H);
        const uint4* src = reinterpret_cast<const uint4*>(in) + (((f * 2 * H + 2 * yo) * 2 * W) + 2 * xo) * cv + c8;
        const uint4 a = src[0], b = src[cv], c = src[(size_t)2 * W * cv], d = src[(size_t)2 * W * cv + cv];
        uint4 o;
        o.x = bf16x2_max(bf16x2_max(a.x, b.x), bf16x2_max(c.x, d.x));
        o.y = bf16x2_max(bf16x2_max(a.y, b.y), bf16x2_max(c.y, d.y));
        o.z = bf16x2_max(bf16x2_max(a.z, b.z), bf16x2_max(c.z, d.z));
        o.w = bf16x2_max(bf16x2_max(a.w, b.w), bf16x2_max(c.w, d.w));
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// last pool: bf16 channels-last [F][14][24][512] -> fp32 channels-first [F][512][7][12] (what forward_feat returns)
__global__ void __launch_bounds__(256) vgg_pool_out_kernel(const bf16* __restrict__ in, float* __restrict__ out) {
    __shared__ float tile[84][65];
    const int f = blockIdx.y, c0 = blockIdx.x * 64, tid = threadIdx.x;
    for (int i = tid; i < 84 * 64; i += 256) {
        const int c = i & 63, pix = i >> 6, yo = pix / 12, xo = pix % 12;
        const bf16* src = in + (((size_t)f * 14 + 2 * yo) * 24 + 2 * xo) * 512 + c0 + c;
        const float m = fmaxf(fmaxf(__bfloat162float(src[0]), __bfloat162float(src[512])),
                              fmaxf(__bfloat162float(src[24 * 512]), __bfloat162float(src[24 * 512 + 512])));
        tile[pix][c] = m;
    }
    __syncthreads();
    for (int i = tid; i < 64 * 84; i += 256) {
        const int pix = i % 84, c = i / 84;
        out[((size_t)f * 512 + c0 + c) * 84 + pix] = tile[pix][c];
    }
}

}  // namespace

struct dsb_vggish {
    int max_frames = 0;
    int num_sms = 148;
    bool finalized = false;
    int launches = 0;
    std::string err;
    struct Wt { float* p; long numel; };
    std::map<std::string, Wt> w;
    std::map<std::string, bf16*> wp;
    std::vector<void*> allocs;
    bf16 *bufA = nullptr, *bufB = nullptr;
};

static int vfail(dsb_vggish* h, int code, const char* fmt, ...) {
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
static int valloc(dsb_vggish* h, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = ((count * sizeof(T) + 255) / 256) * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return vfail(h, DSB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

extern "C" int dsb_vggish_create(int max_frames, dsb_vggish** out) {
    if (!out || max_frames < 1 || max_frames > 4096) return DSB_ERR_ARG;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return DSB_ERR_CUDA;
    if (prop.major != 10) return DSB_ERR_UNSUPPORTED;             // sm_100a only, no fallback
    dsb_vggish* h = new dsb_vggish();
    h->max_frames = max_frames;
    h->num_sms = prop.multiProcessorCount;
    if (gemm_init()) { delete h; return DSB_ERR_CUDA; }
    *out = h;
    return DSB_OK;
}

extern "C" void dsb_vggish_destroy(dsb_vggish* h) {
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" const char* dsb_vggish_last_error(const dsb_vggish* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int dsb_vggish_last_launch_count(const dsb_vggish* h) { return h ? h->launches : 0; }

extern "C" int dsb_vggish_load_weight(dsb_vggish* h, const char* ref_key, const void* data, const int64_t* shape, int ndim) {
    if (!h || !ref_key || !data || ndim < 0 || ndim > 8) return vfail(h, DSB_ERR_ARG, "dsb_vggish_load_weight: bad argument");
    if (h->finalized) return vfail(h, DSB_ERR_ARG, "dsb_vggish_load_weight after dsb_vggish_finalize");
    if (strncmp(ref_key, "features.", 9) != 0) return DSB_OK;      // the embeddings MLP is not on forward_feat's path
    long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    if (numel < 1) return vfail(h, DSB_ERR_ARG, "weight '%s' is empty", ref_key);
    float* p = nullptr;
    if (int r = valloc(h, &p, (size_t)numel)) return r;
    if (cudaMemcpy(p, data, (size_t)numel * sizeof(float), cudaMemcpyDefault) != cudaSuccess)
        return vfail(h, DSB_ERR_CUDA, "copy of weight '%s' failed", ref_key);
    h->w[ref_key] = {p, numel};
    return DSB_OK;
}

static const float* VW(dsb_vggish* h, const std::string& k, long numel) {
    auto it = h->w.find(k);
    return (it == h->w.end() || it->second.numel != numel) ? nullptr : it->second.p;
}

namespace {
const int kIdx[6] = {0, 3, 6, 8, 11, 13};
const int kCout[6] = {64, 128, 256, 256, 512, 512};
}  // namespace

extern "C" int dsb_vggish_finalize(dsb_vggish* h) {
    if (!h) return DSB_ERR_ARG;
    if (h->finalized) return vfail(h, DSB_ERR_ARG, "weights already finalized");
    int cin = 1;
    for (int l = 0; l < 6; ++l) {
        const std::string k = "features." + std::to_string(kIdx[l]) + ".";
        const float* w = VW(h, k + "weight", (long)kCout[l] * cin * 9);
        if (!w || !VW(h, k + "bias", kCout[l])) return vfail(h, DSB_ERR_WEIGHT, "missing / mis-shaped weight '%s*'", k.c_str());
        if (l > 0) {
            bf16* dst = nullptr;
            if (int r = valloc(h, &dst, (size_t)kCout[l] * cin * 9)) return r;
            if (int r = pack_weight_launch(w, kCout[l], cin, 9, dst, 0)) return vfail(h, DSB_ERR_CUDA, "pack_weight launch %d", r);
            h->wp[k + "weight"] = dst;
        }
        cin = kCout[l];
    }
    // ping-pong activation buffers: the largest tensor is conv 64->128's output [F][56][96][128]
    const size_t big = (size_t)h->max_frames * 56 * 96 * 128;
    if (int r = valloc(h, &h->bufA, big)) return r;
    if (int r = valloc(h, &h->bufB, big)) return r;
    if (cudaDeviceSynchronize() != cudaSuccess) return vfail(h, DSB_ERR_CUDA, "weight repack failed");
    h->finalized = true;
    return DSB_OK;
}

extern "C" int dsb_vggish_forward_feat(dsb_vggish* h, const float* audio, float* out, int frames, void* stream) {
    if (!h || !audio || !out) return DSB_ERR_ARG;
    if (!h->finalized) return vfail(h, DSB_ERR_ARG, "dsb_vggish_forward_feat before dsb_vggish_finalize");
    if (frames < 1 || frames > h->max_frames) return vfail(h, DSB_ERR_ARG, "%d frames outside [1, %d]", frames, h->max_frames);
    cudaStream_t s = (cudaStream_t)stream;
    h->launches = 0;
    const int F = frames;
    vgg_conv0_pool_kernel<<<dim3(56, F), 256, 0, s>>>(audio, VW(h, "features.0.weight", 576), VW(h, "features.0.bias", 64), h->bufA);
    ++h->launches;
    auto conv = [&](int l, int H, int W, const bf16* in, bf16* o) -> int {
        const std::string k = "features." + std::to_string(kIdx[l]) + ".";
        ConvOp op;
        memset(&op, 0, sizeof(op));
        op.kind = CONV_3X3; op.F = F; op.H = H; op.W = W; op.Cin = kCout[l - 1]; op.N = kCout[l]; op.dilation = 1;
        op.A = in; op.Wt = h->wp[k + "weight"];
        op.shift = VW(h, k + "bias", kCout[l]); op.act = ACT_RELU; op.out_bf16 = o;
        ConvLaunch cl;
        if (int r = conv_lower(op, &cl)) return vfail(h, DSB_ERR_CUDA, "features.%d: lowering failed (%d)", kIdx[l], r);
        if (int r = conv_run(cl, h->num_sms, s)) return vfail(h, DSB_ERR_CUDA, "features.%d: launch failed (%d)", kIdx[l], r);
        ++h->launches;
        return 0;
    };
    auto pool = [&](int H, int W, int C, const bf16* in, bf16* o) {
        const long total = (long)F * H * W * (C / 8);
        long g = (total + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        vgg_pool_kernel<<<(int)g, 256, 0, s>>>(in, H, W, C, total, o);
        ++h->launches;
    };
    if (int r = conv(1, 56, 96, h->bufA, h->bufB)) return r;         // 64 -> 128 at 56x96
    pool(28, 48, 128, h->bufB, h->bufA);
    if (int r = conv(2, 28, 48, h->bufA, h->bufB)) return r;         // 128 -> 256
    if (int r = conv(3, 28, 48, h->bufB, h->bufA)) return r;         // 256 -> 256
    pool(14, 24, 256, h->bufA, h->bufB);
    if (int r = conv(4, 14, 24, h->bufB, h->bufA)) return r;         // 256 -> 512
    if (int r = conv(5, 14, 24, h->bufA, h->bufB)) return r;         // 512 -> 512
    vgg_pool_out_kernel<<<dim3(8, F), 256, 0, s>>>(h->bufB, out);
    ++h->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return vfail(h, DSB_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
    return DSB_OK;
}
