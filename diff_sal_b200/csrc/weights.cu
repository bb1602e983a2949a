#include "weights.cuh"

namespace dsb {

__global__ void pack_weight_kernel(const float* __restrict__ src, int N, int Cin, int taps, bf16* __restrict__ dst, int f16) {
    const long total = (long)N * Cin * taps;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const int tap = (int)((i / Cin) % taps);
        const long n = i / ((long)Cin * taps);
        const float v = src[(n * Cin + c) * taps + tap];
        if (f16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
        else dst[i] = __float2bfloat16(v);
    }
}

int pack_weight_launch(const float* src, int N, int Cin, int taps, bf16* dst, cudaStream_t s, int f16) {
    const long total = (long)N * Cin * taps;
    long g = (total + 255) / 256;
    if (g > 4096) g = 4096;
    pack_weight_kernel<<<(int)g, 256, 0, s>>>(src, N, Cin, taps, dst, f16);
    return (int)cudaGetLastError();
}

__global__ void pack_dw_kernel(const float* __restrict__ src, int C, int taps, int src_stride, int src_off,
                               float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * taps) return;
    const int c = i % C, tap = i / C;
    dst[i] = src[(size_t)c * src_stride + src_off + tap];
}

int pack_dw_launch(const float* src, int C, int taps, int src_stride, int src_off, float* dst, cudaStream_t s) {
    pack_dw_kernel<<<(C * taps + 255) / 256, 256, 0, s>>>(src, C, taps, src_stride, src_off, dst);
    return (int)cudaGetLastError();
}

__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var,
                               const float* conv_bias, int C, float eps, float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = w[c] / sqrtf(var[c] + eps);
    const float cb = conv_bias ? conv_bias[c] : 0.0f;
    scale[c] = sc;
    shift[c] = (cb - mean[c]) * sc + b[c];
}

int bn_fold_launch(const float* w, const float* b, const float* mean, const float* var, const float* conv_bias, int C,
                   float eps, float* scale, float* shift, cudaStream_t s) {
    bn_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(w, b, mean, var, conv_bias, C, eps, scale, shift);
    return (int)cudaGetLastError();
}

__global__ void transpose_kernel(const float* __restrict__ src, int R, int Cc, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * Cc) return;
    const int r = i / Cc, c = i % Cc;
    dst[(size_t)c * R + r] = src[i];
}

int transpose_launch(const float* src, int R, int Cc, float* dst, cudaStream_t s) {
    transpose_kernel<<<(R * Cc + 255) / 256, 256, 0, s>>>(src, R, Cc, dst);
    return (int)cudaGetLastError();
}

// w5[(r*5+q)*96 + co] = sum_ci sum_{dy+a=r, dx+b=q} w_d[co][ci][dy][dx] * w_in[ci][a][b]
// b5[co] = b_d[co] + sum_{ci,dy,dx} w_d[co][ci][dy][dx] * b_in[ci]
__global__ void stem_compose_kernel(const float* w_in, const float* b_in, const float* w_d, const float* b_d, float* w5,
                                    float* b5) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 25 * 96) {
        const int co = i % 96, rq = i / 96, r = rq / 5, q = rq % 5;
        double acc = 0.0;
        for (int ci = 0; ci < 96; ++ci)
            for (int dy = 0; dy < 3; ++dy) {
                const int a = r - dy;
                if (a < 0 || a > 2) continue;
                for (int dx = 0; dx < 3; ++dx) {
                    const int b = q - dx;
                    if (b < 0 || b > 2) continue;
                    acc += (double)w_d[((co * 96 + ci) * 3 + dy) * 3 + dx] * (double)w_in[(ci * 3 + a) * 3 + b];
                }
            }
        w5[i] = (float)acc;
    } else if (i < 25 * 96 + 96) {
        const int co = i - 25 * 96;
        double acc = (double)b_d[co];
        for (int ci = 0; ci < 96; ++ci) {
            double ws = 0.0;
            for (int k = 0; k < 9; ++k) ws += (double)w_d[(co * 96 + ci) * 9 + k];
            acc += ws * (double)b_in[ci];
        }
        b5[co] = (float)acc;
    }
}

int stem_compose_launch(const float* w_in, const float* b_in, const float* w_d, const float* b_d, float* w5, float* b5,
                        cudaStream_t s) {
    stem_compose_kernel<<<(25 * 96 + 96 + 127) / 128, 128, 0, s>>>(w_in, b_in, w_d, b_d, w5, b5);
    return (int)cudaGetLastError();
}

}  // namespace dsb
