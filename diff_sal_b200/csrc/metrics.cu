// Saliency metrics on the device: CC / SIM / NSS / AUC-Judd of the reference's offline evaluation
// (metrics/metrics.py:7-64,178-252 with metrics/utils.py:11-52 `normalize`), so that predicted maps need not leave
// the GPU to be scored.  One block per clip, fp64 accumulation, two passes over the three maps (they stay in L2).
//
//   CC   = corrcoef(standardise(p), standardise(g))                     (scale invariant: one pass of moments)
//   NSS  = mean over fixations of (p - mean p) / std p                   (population std, as np.std)
//   SIM  = sum min(a, b), a = range-normalised p divided by its sum, b likewise for g
//   AUCJ = area under (fp, tp) with one threshold per fixation value (descending), tp_k = (k+1)/n_fix,
//          fp_k = (#{p >= thr_k} - (k+1)) / (n_pix - n_fix), end points (0,0) and (1,1)
// The reference adds rand*1e-7 jitter from numpy's global RNG before AUC-J (metrics.py:44-45); pass the same numbers in
// `jitter` (fp64, [B][n]) to reproduce it, or NULL for none.  At most 1024 fixations per clip.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "diffsal_b200.h"

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxFix = 1024;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reduction of NV values with `op` (0 sum, 1 min, 2 max); result broadcast to every thread
template <int NV>
__device__ void block_reduce(double (&v)[NV], const int (&op)[NV], double* scratch /*[32][NV]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = op[i] == 0 ? warp_sum_d(v[i]) : (op[i] == 1 ? warp_min_d(v[i]) : warp_max_d(v[i]));
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[wid * NV + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double a = scratch[i];                              // warps are folded in index order: deterministic
        for (int w = 1; w < kThreads / 32; ++w) {
            const double b = scratch[w * NV + i];
            a = op[i] == 0 ? a + b : (op[i] == 1 ? fmin(a, b) : fmax(a, b));
        }
        v[i] = a;
    }
}

__global__ void __launch_bounds__(kThreads) metrics_kernel(const float* __restrict__ pred, const float* __restrict__ dens,
                                                          const float* __restrict__ fix, const double* __restrict__ jitter,
                                                          int n, double* __restrict__ out) {
    __shared__ double scratch[32 * 11];
    __shared__ double thr[kMaxFix];
    __shared__ unsigned hist[kMaxFix + 1];
    __shared__ unsigned nfix_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* p = pred + (size_t)b * n;
    const float* g = dens + (size_t)b * n;
    const float* fx = fix + (size_t)b * n;
    const double* jt = jitter ? jitter + (size_t)b * n : nullptr;
    if (tid == 0) nfix_s = 0;
    for (int i = tid; i <= kMaxFix; i += kThreads) hist[i] = 0;
    __syncthreads();

    // ---- pass 1: moments, extrema, fixation statistics; fixation values (with jitter) collected for AUC-J
    double v[11] = {0, 0, 0, 0, 0, 1e300, -1e300, 1e300, -1e300, 0, 0};
    // 0 sum p, 1 sum p^2, 2 sum g, 3 sum g^2, 4 sum pg, 5 min p, 6 max p, 7 min g, 8 max g, 9 sum_fix p, 10 n_fix
    for (int i = tid; i < n; i += kThreads) {
        const double a = (double)p[i], c = (double)g[i];
        v[0] += a; v[1] += a * a; v[2] += c; v[3] += c * c; v[4] += a * c;
        v[5] = fmin(v[5], a); v[6] = fmax(v[6], a); v[7] = fmin(v[7], c); v[8] = fmax(v[8], c);
        if (fx[i] > 0.5f) {
            v[9] += a; v[10] += 1.0;
            const unsigned slot = atomicAdd(&nfix_s, 1u);
            if (slot < kMaxFix) thr[slot] = a + (jt ? jt[i] : 0.0);
        }
    }
    const int ops[11] = {0, 0, 0, 0, 0, 1, 2, 1, 2, 0, 0};
    block_reduce<11>(v, ops, scratch);
    const double N = (double)n;
    const double mp = v[0] / N, mg = v[2] / N;
    const double vp = fmax(v[1] / N - mp * mp, 0.0), vg = fmax(v[3] / N - mg * mg, 0.0);
    const double cc = (v[4] / N - mp * mg) / (sqrt(vp) * sqrt(vg));
    const double nfix = v[10];
    const double nss = nfix > 0 ? (v[9] / nfix - mp) / sqrt(vp) : nan("");
    const double rp = v[6] - v[5], rg = v[8] - v[7];
    const double sa = (v[0] - N * v[5]) / rp, sb = (v[2] - N * v[7]) / rg;     // sums of the range-normalised maps

    // ---- pass 2: SIM, and the AUC-J histogram (pixel counted at the first threshold it reaches)
    const int nf = (int)min(nfix_s, (unsigned)kMaxFix);
    // sort the fixation values descending (bitonic network over kMaxFix slots padded with -inf)
    for (int i = nf + tid; i < kMaxFix; i += kThreads) thr[i] = -1e300;
    __syncthreads();
    for (int k = 2; k <= kMaxFix; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int i = tid, ixj = i ^ j;
            if (ixj > i) {
                const double x = thr[i], y = thr[ixj];
                const bool desc = (i & k) == 0;
                if (desc ? (x < y) : (x > y)) { thr[i] = y; thr[ixj] = x; }
            }
            __syncthreads();
        }
    double s2[1] = {0.0};
    for (int i = tid; i < n; i += kThreads) {
        const double a = ((double)p[i] - v[5]) / rp / sa, c = ((double)g[i] - v[7]) / rg / sb;
        s2[0] += fmin(a, c);
        if (nf > 0) {
            const double val = (double)p[i] + (jt ? jt[i] : 0.0);
            // first k with thr[k] <= val (thr descending); nf if none
            int lo = 0, hi = nf;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (thr[mid] <= val) hi = mid; else lo = mid + 1;
            }
            atomicAdd(&hist[lo], 1u);
        }
    }
    const int ops2[1] = {0};
    block_reduce<1>(s2, ops2, scratch);
    __syncthreads();
    if (tid == 0) {
        double auc = nan("");
        if (nf > 0 && nfix_s <= (unsigned)kMaxFix) {
            // trapezoid over (fp, tp): (0,0), (fp_k, tp_k) for k = 0..nf-1, (1,1)
            double area = 0.0, fp_prev = 0.0, tp_prev = 0.0;
            unsigned above = 0;
            for (int k = 0; k < nf; ++k) {
                above += hist[k];
                const double tp = (double)(k + 1) / (double)nf;
                const double fp = ((double)above - (double)(k + 1)) / (N - (double)nf);
                area += (fp - fp_prev) * (tp + tp_prev) * 0.5;
                fp_prev = fp; tp_prev = tp;
            }
            area += (1.0 - fp_prev) * (1.0 + tp_prev) * 0.5;
            auc = area;
        }
        out[b * 4 + 0] = cc;
        out[b * 4 + 1] = s2[0];
        out[b * 4 + 2] = nss;
        out[b * 4 + 3] = auc;
    }
}

// Validation losses of the reference's sampling loop (models/sal_losses.py:14-176, get_kl_cc_sim_loss_wo_weight :207-233,
// called on every validated batch at diffusion_trainer.py:741,797,868): per clip
//   kl  = sum g' log(eps + g' / (s' + eps)),   s' = s / sum s, g' = g / sum g,  eps = 2.2204e-16         (kldiv2)
//   cc  = sum a b / sqrt(sum a^2 sum b^2),     a, b = maps standardised with the UNBIASED std (torch.std) (cc_s2)
//   sim = sum min(s~, g~),                     s~, g~ = min-max normalised maps divided by their sums      (similarity2)
//   nss = sum ((s - mean s) / (std s + eps)) g / sum g                                                     (nss2)
// One block per clip, two passes (moments / extrema, then the four sums), fp64 accumulation; the batch mean of each
// column is what the reference returns.
__global__ void __launch_bounds__(kThreads) val_losses_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                              int n, double* __restrict__ out) {
    __shared__ double scratch[32 * 8];
    const int b = blockIdx.x;
    const float* p = pred + (size_t)b * n;
    const float* g = gt + (size_t)b * n;
    double v[8] = {0.0, 0.0, 0.0, 0.0, 1e300, -1e300, 1e300, -1e300};      // sum s, sum g, (unused), (unused), min/max s, min/max g
    const int op1[8] = {0, 0, 0, 0, 1, 2, 1, 2};
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const double a = p[i], c = g[i];
        v[0] += a; v[1] += c;
        v[4] = fmin(v[4], a); v[5] = fmax(v[5], a); v[6] = fmin(v[6], c); v[7] = fmax(v[7], c);
    }
    block_reduce<8>(v, op1, scratch);
    const double N = (double)n, sum_s = v[0], sum_g = v[1], mean_s = sum_s / N, mean_g = sum_g / N;
    const double min_s = v[4], max_s = v[5], min_g = v[6], max_g = v[7];
    const double rng_s = max_s - min_s, rng_g = max_g - min_g;
    const double nsum_s = (sum_s - N * min_s) / rng_s, nsum_g = (sum_g - N * min_g) / rng_g;   // sums of the min-max maps
    const double eps = 2.2204e-16;
    double w[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // kl, sum ds^2, sum dg^2, sum ds dg, sim, sum ds g, -, -
    const int op2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const double a = p[i], c = g[i];
        const double sp = a / sum_s, gp = c / sum_g;
        w[0] += gp * log(eps + gp / (sp + eps));
        const double ds = a - mean_s, dg = c - mean_g;
        w[1] += ds * ds; w[2] += dg * dg; w[3] += ds * dg;
        w[4] += fmin(((a - min_s) / rng_s) / nsum_s, ((c - min_g) / rng_g) / nsum_g);
        w[5] += ds * c;
    }
    block_reduce<8>(w, op2, scratch);
    if (threadIdx.x == 0) {
        const double std_s = sqrt(w[1] / (N - 1.0));                       // torch.std: unbiased
        out[b * 4 + 0] = w[0];
        out[b * 4 + 1] = w[3] / sqrt(w[1] * w[2]);                         // the standard deviations cancel
        out[b * 4 + 2] = w[4];
        out[b * 4 + 3] = (w[5] / (std_s + eps)) / sum_g;
    }
}

}  // namespace

extern "C" int dsb_val_losses(const float* pred, const float* gt, int B, int64_t elems_per_clip, double* out4, void* stream) {
    if (!pred || !gt || !out4 || B < 1 || elems_per_clip < 2 || elems_per_clip > (1 << 30)) return DSB_ERR_ARG;
    val_losses_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(pred, gt, (int)elems_per_clip, out4);
    return cudaGetLastError() == cudaSuccess ? DSB_OK : DSB_ERR_CUDA;
}

extern "C" int dsb_metrics(const float* pred, const float* density, const float* fixations, const double* jitter_or_null,
                           int B, int64_t pixels_per_map, double* out4, void* stream) {
    if (!pred || !density || !fixations || !out4 || B < 1 || pixels_per_map < 2 || pixels_per_map > (1 << 30)) return DSB_ERR_ARG;
    metrics_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(pred, density, fixations, jitter_or_null, (int)pixels_per_map, out4);
    return cudaGetLastError() == cudaSuccess ? DSB_OK : DSB_ERR_CUDA;
}
