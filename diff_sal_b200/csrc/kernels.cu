// Memory-bound kernels of the SalUNet denoiser (see kernels.cuh).  Vectorised, coalesced over the channel axis,
// warp-shuffle reductions for the LayerNorm / softmax statistics.
#include "kernels.cuh"

#include <stdlib.h>

namespace dsb {

#define DSB_LAUNCH_CHECK() return (int)cudaGetLastError()
// launch with the programmatic-dependent-launch attribute (see common.cuh); errors surface through cudaGetLastError
#define DSB_PDL_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

int& pdl_mode() {
    static int mode = [] { const char* e = getenv("DSB_PDL"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0; }();
    return mode;
}
bool& pdl_allow_next() {
    static thread_local bool allow = true;
    return allow;
}

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? red[lane] : 0.0f;
    return warp_sum(t);
}

// ------------------------------------------------------------------------------------------ timestep embedding
// All weights are pre-transposed to [in][out] so that thread j reads column j with coalesced loads.
// grid (B, 4): every block recomputes the 2-layer MLP (cheap) and produces a quarter of the 1344 projection outputs.
// Three dependent matrix-vector products per clip: latency bound, so every layer is split 4 ways along k (thread =
// 4 consecutive outputs x one k slice, float4 weight loads, 8 loads in flight) and folded through shared memory.
__device__ __forceinline__ float4 temb_slice(const float* __restrict__ wt, int ld, int col, const float* __restrict__ v,
                                             int k0, int k1) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(wt + (size_t)k * ld + col);
        const float x = v[k];
        acc.x = fmaf(w.x, x, acc.x); acc.y = fmaf(w.y, x, acc.y); acc.z = fmaf(w.z, x, acc.z); acc.w = fmaf(w.w, x, acc.w);
    }
    return acc;
}

__global__ void __launch_bounds__(384) temb_kernel(const float* __restrict__ t, TembWeights w, float* tp0, float* tp1,
                                                  float* tp2) {
    pdl_trigger();
    pdl_wait();
    __shared__ float emb[96];
    __shared__ float h1[384];
    __shared__ float h2[384];
    __shared__ float4 part[4][96];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int og = tid % 96, ks = tid / 96;                  // output group (4 outputs), k slice
    const float tv = t[b];
    if (tid < 48) {
        const float fr = expf((float)tid * -0.19596468876545072f);   // exp(-k * ln(1e4) / 47)
        const float a = tv * fr;
        emb[tid] = sinf(a);
        emb[48 + tid] = cosf(a);
    }
    __syncthreads();
    part[ks][og] = temb_slice(w.w0, 384, 4 * og, emb, ks * 24, ks * 24 + 24);
    __syncthreads();
    {
        const float* pf = reinterpret_cast<const float*>(part);
        h1[tid] = swishf(w.b0[tid] + ((pf[tid] + pf[384 + tid]) + (pf[768 + tid] + pf[1152 + tid])));
    }
    __syncthreads();
    part[ks][og] = temb_slice(w.w1, 384, 4 * og, h1, ks * 96, ks * 96 + 96);
    __syncthreads();
    {
        const float* pf = reinterpret_cast<const float*>(part);
        // only swish(temb) is consumed (sal_unet.py:129)
        h2[tid] = swishf(w.b1[tid] + ((pf[tid] + pf[384 + tid]) + (pf[768 + tid] + pf[1152 + tid])));
    }
    __syncthreads();
    // 1344 = 192 + 384 + 768 projection outputs, 336 per block = 84 groups of 4 (group boundaries never straddle a layer)
    const int pg = tid % 84, pks = tid / 84;
    int i = 0, r = 0, co = 0;
    if (pks < 4) {
        const int o = blockIdx.y * 336 + 4 * pg;
        if (o < 192) { i = 0; r = o; } else if (o < 576) { i = 1; r = o - 192; } else { i = 2; r = o - 576; }
        co = w.cout[i];
        part[pks][pg] = temb_slice(w.wp[i], co, r, h2, pks * 96, pks * 96 + 96);
    }
    __syncthreads();
    if (tid < 84) {
        const float4 p0 = part[0][tid], p1 = part[1][tid], p2 = part[2][tid], p3 = part[3][tid];
        const float4 bb = *reinterpret_cast<const float4*>(w.bp[i] + r);
        float* out = i == 0 ? tp0 : (i == 1 ? tp1 : tp2);
        *reinterpret_cast<float4*>(out + (size_t)b * co + r) =
            make_float4(bb.x + ((p0.x + p1.x) + (p2.x + p3.x)), bb.y + ((p0.y + p1.y) + (p2.y + p3.y)),
                        bb.z + ((p0.z + p1.z) + (p2.z + p3.z)), bb.w + ((p0.w + p1.w) + (p2.w + p3.w)));
    }
}

int temb_launch(const float* t, int B, const TembWeights& w, float* const tp[3], cudaStream_t s) {
    DSB_PDL_LAUNCH(temb_kernel, dim3(B, 4), 384, 0, s, t, w, tp[0], tp[1], tp[2]);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ stem 5x5 stride 4
__global__ void __launch_bounds__(96) stem_kernel(const float* __restrict__ x, const float* __restrict__ w5,
                                                 const float* __restrict__ b5, float* __restrict__ h0) {
    pdl_trigger();
    pdl_wait();
    __shared__ float patch[5][132];
    const int xo0 = blockIdx.x * 32, y = blockIdx.y, b = blockIdx.z, c = threadIdx.x;
    const float* xb = x + (size_t)b * 224 * 384;
    for (int i = threadIdx.x; i < 5 * 129; i += 96) {
        const int r = i / 129, col = i % 129;
        const int yy = 4 * y - 1 + r, xx = 4 * xo0 - 1 + col;
        patch[r][col] = (yy >= 0 && yy < 224 && xx >= 0 && xx < 384) ? xb[yy * 384 + xx] : 0.0f;
    }
    float w[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) w[k] = w5[k * 96 + c];
    const float bias = b5[c];
    __syncthreads();
    float* out = h0 + (((size_t)b * 56 + y) * 96 + xo0) * 96 + c;
    for (int px = 0; px < 32; ++px) {
        float acc = bias;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int q = 0; q < 5; ++q) acc = fmaf(w[r * 5 + q], patch[r][4 * px + q], acc);
        out[(size_t)px * 96] = acc;
    }
}

int stem_launch(const float* x, int B, const float* w5, const float* b5, float* h0, cudaStream_t s) {
    DSB_PDL_LAUNCH(stem_kernel, dim3(3, 56, B), 96, 0, s, x, w5, b5, h0);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ GroupNorm
// Deterministic two-level reduction (no atomics): block (split, frame) writes the (sum, sum of squares) of its
// pixel range for each of the 32 groups to part[frame][split][group]; gn_apply adds the splits in a fixed order.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int HW, int C, double* part) {
    pdl_trigger();
    pdl_wait();
    __shared__ float ssum[960];
    __shared__ float ssq[960];
    const int f = blockIdx.y, tid = threadIdx.x;
    const int cv_n = C >> 2;
    const int P = 256 / cv_n;                       // pixel lanes (10, 5, 2, 1 for C = 96..768)
    const int cv = tid % cv_n, pl = tid / cv_n;
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per;
    const int p1 = min(HW, p0 + per);
    if (pl < P) {
        float4 s4 = make_float4(0, 0, 0, 0), q4 = make_float4(0, 0, 0, 0);
        const float4* base = reinterpret_cast<const float4*>(x + (size_t)f * HW * C) + cv;
        for (int px = p0 + pl; px < p1; px += P) {
            const float4 v = base[(size_t)px * cv_n];
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            q4.x = fmaf(v.x, v.x, q4.x); q4.y = fmaf(v.y, v.y, q4.y); q4.z = fmaf(v.z, v.z, q4.z); q4.w = fmaf(v.w, v.w, q4.w);
        }
        float* ps = ssum + pl * C + 4 * cv;
        float* pq = ssq + pl * C + 4 * cv;
        ps[0] = s4.x; ps[1] = s4.y; ps[2] = s4.z; ps[3] = s4.w;
        pq[0] = q4.x; pq[1] = q4.y; pq[2] = q4.z; pq[3] = q4.w;
    }
    __syncthreads();
    if (tid < 32) {
        const int cg = C >> 5;
        double s = 0.0, q = 0.0;
        for (int l = 0; l < P; ++l)
            for (int k = 0; k < cg; ++k) { s += (double)ssum[l * C + tid * cg + k]; q += (double)ssq[l * C + tid * cg + k]; }
        double* o = part + (size_t)f * 4096 + ((size_t)blockIdx.x * 32 + tid) * 2;
        o[0] = s;
        o[1] = q;
        __threadfence();
    }
    // the last block of a frame to arrive folds the splits (fixed order -> bitwise reproducible) into (mean, rstd)
    __shared__ bool last;
    unsigned* counter = reinterpret_cast<unsigned*>(part + (size_t)f * 4096 + 4064);
    __syncthreads();
    if (tid == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int grp = tid >> 3, sl = tid & 7, S = gridDim.x;
    double sm = 0.0, sq = 0.0;
    for (int k = sl; k < S; k += 8) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(part + (size_t)f * 4096 + ((size_t)k * 32 + grp) * 2));
        sm += v.x;
        sq += v.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (sl == 0) {
        const double n = (double)HW * (C >> 5);
        const double m = sm / n;
        double v = sq / n - m * m;
        if (v < 0.0) v = 0.0;
        float2* fin = reinterpret_cast<float2*>(part + (size_t)f * 4096 + 4032);
        fin[grp] = make_float2((float)m, (float)(1.0 / sqrt(v + 1e-6)));
    }
    if (tid == 0) *counter = 0;
}

// scratch layout, 4096 doubles per frame: partials [S <= 63][32][2] | (mean, rstd) float2 [32] | arrival counter
static int gn_splits(int HW) { int S = HW / 8; return S > 63 ? 63 : (S < 1 ? 1 : S); }

int gn_stats_launch(const float* x, int F, int HW, int C, double* part, cudaStream_t s) {
    if (C % 32 || C > 768 || C < 96) return -30;
    DSB_PDL_LAUNCH(gn_stats_kernel, dim3(gn_splits(HW), F), 256, 0, s, x, HW, C, part);
    DSB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, int HW, int C, int S,
                                                      const double* __restrict__ part_sums, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, bf16* __restrict__ out_act,
                                                      bf16* __restrict__ out_raw) {
    pdl_trigger();
    pdl_wait();
    __shared__ float mean[32];
    __shared__ float rstd[32];
    const int f = blockIdx.y, tid = threadIdx.x;
    const int cg = C >> 5;
    if (tid < 32) {
        const float2 v = reinterpret_cast<const float2*>(part_sums + (size_t)f * 4096 + 4032)[tid];
        mean[tid] = v.x;
        rstd[tid] = v.y;
    }
    __syncthreads();
    // gridDim.x * 256 is a multiple of C / 4 (launch picks gridDim.x % 3 == 0), so a thread keeps one 4-channel vector
    // for its whole walk: the affine (x * a + b) of its channels is hoisted out of the loop
    const int cv_n = C >> 2;
    const int total = HW * cv_n;
    const int i0 = blockIdx.x * 256 + tid, stride = gridDim.x * 256;
    const int cv = i0 % cv_n, c = 4 * cv;
    const float4 g = reinterpret_cast<const float4*>(gamma)[cv];
    const float4 bt = reinterpret_cast<const float4*>(beta)[cv];
    const int g0 = c / cg, g1 = (c + 1) / cg, g2 = (c + 2) / cg, g3 = (c + 3) / cg;
    const float ax = rstd[g0] * g.x, ay = rstd[g1] * g.y, az = rstd[g2] * g.z, aw = rstd[g3] * g.w;
    const float bx = fmaf(-mean[g0], ax, bt.x), by = fmaf(-mean[g1], ay, bt.y), bz = fmaf(-mean[g2], az, bt.z),
                bw = fmaf(-mean[g3], aw, bt.w);
    const float4* xin = reinterpret_cast<const float4*>(x + (size_t)f * HW * C);
    uint2* oa = reinterpret_cast<uint2*>(out_act + (size_t)f * HW * C);
    uint2* orw = out_raw ? reinterpret_cast<uint2*>(out_raw + (size_t)f * HW * C) : nullptr;
#pragma unroll 4
    for (int i = i0; i < total; i += stride) {
        const float4 v = xin[i];
        oa[i] = make_uint2(pack_bf16x2(swishf(fmaf(v.x, ax, bx)), swishf(fmaf(v.y, ay, by))),
                           pack_bf16x2(swishf(fmaf(v.z, az, bz)), swishf(fmaf(v.w, aw, bw))));
        if (orw) orw[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

int gn_apply_launch(const float* x, int F, int HW, int C, const double* acc, const float* gamma, const float* beta,
                    bf16* out_act, bf16* out_raw, cudaStream_t s) {
    long total = (long)HW * (C / 4);
    int gx = (int)((total + 767) / 768) * 3;             // multiple of 3: 768 threads cover whole C/4 periods
    if (gx > 297) gx = 297;
    DSB_PDL_LAUNCH(gn_apply_kernel, dim3(gx, F), 256, 0, s, x, HW, C, gn_splits(HW), acc, gamma, beta, out_act, out_raw);
    DSB_LAUNCH_CHECK();
}

// GroupNorm + swish in ONE launch: a thread-block cluster of CL CTAs owns a frame.  Pass 1: every CTA reduces its pixel
// slice to per-group (sum, sum of squares) in fp64; the partials meet through distributed shared memory (each CTA reads
// all CL partial sets in rank order -> the same bits in every CTA, run to run); pass 2 re-reads the slice (L2-resident:
// the encoder tensors are 2 - 33 MB) and writes bf16(swish(GN(x))) and, optionally, bf16(x).  Replaces the
// gn_stats_kernel / gn_apply_kernel pair (two launches and a last-block fold per GroupNorm, six GroupNorms per evaluation).
// Round 2b: the CTA's slice of the frame (HW / CL pixels x C channels, 32 - 258 KB) is fetched ONCE by 1-D TMA bulk
// copies into shared memory, all requested up front by one thread; both passes then read shared memory.  The first
// version streamed the slice twice through ordinary loads: with one 512-thread CTA per SM only ~12 KB were in flight per
// SM and the six GroupNorms of the (strictly serial) noise encoder ran at 0.4 - 1.2 TB/s.  What does not fit in the
// kGnSmemCap bytes (only the C = 192 map of the first block: 258 KB per CTA) is streamed as before.
constexpr uint32_t kGnSmemCap = 192u * 1024u;
constexpr uint32_t kGnChunk = 32u * 1024u;

template <int CL>
__global__ void __launch_bounds__(512) gn_fused_kernel(const float* __restrict__ x, int HW, int C,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      bf16* __restrict__ out_act, bf16* __restrict__ out_raw) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) float4 gsl[];         // resident part of the slice
    __shared__ float ssum[4][768];
    __shared__ float ssq[4][768];
    __shared__ double part[32][2];
    __shared__ float mean[32], rstd[32];
    __shared__ __align__(8) uint64_t bar;
    const unsigned rank = cluster_ctarank();
    const int f = blockIdx.x / CL, tid = threadIdx.x;
    const int cv_n = C >> 2, cg = C >> 5;
    const int per = (HW + CL - 1) / CL;
    const int p0 = (int)rank * per, p1 = min(HW, p0 + per);
    const float4* base = reinterpret_cast<const float4*>(x + (size_t)f * HW * C);
    const int res_px = max(0, min(p1 - p0, (int)(kGnSmemCap / ((uint32_t)C * 4u))));   // pixels kept in shared memory
    const int pr = p0 + res_px;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)res_px * (uint32_t)C * 4u;
        mbar_expect_tx(&bar, bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(base + (size_t)p0 * cv_n);
        const uint32_t dst = smem_u32(gsl);
        for (uint32_t o = 0; o < bytes; o += kGnChunk) tma_bulk_load(dst + o, src + o, min(kGnChunk, bytes - o), &bar);
    }
    // ---- pass 1: thread = (4-channel vector, pixel lane); 512 threads = P pixel lanes x cv_n vectors (P = 21 .. 2)
    const int P = min(4, 512 / cv_n);                     // smem holds 4 lanes; more lanes fold into them below
    const int PL = 512 / cv_n;                            // pixel lanes actually running
    const int cv = tid % cv_n, pl = tid / cv_n;
    float4 s4 = make_float4(0, 0, 0, 0), q4 = make_float4(0, 0, 0, 0);
    mbar_wait(&bar, 0u);
    if (pl < PL) {
        int px = p0 + pl;
#pragma unroll 8
        for (; px < pr; px += PL) {
            const float4 v = gsl[(size_t)(px - p0) * cv_n + cv];
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            q4.x = fmaf(v.x, v.x, q4.x); q4.y = fmaf(v.y, v.y, q4.y); q4.z = fmaf(v.z, v.z, q4.z); q4.w = fmaf(v.w, v.w, q4.w);
        }
#pragma unroll 8
        for (; px < p1; px += PL) {
            const float4 v = base[(size_t)px * cv_n + cv];
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            q4.x = fmaf(v.x, v.x, q4.x); q4.y = fmaf(v.y, v.y, q4.y); q4.z = fmaf(v.z, v.z, q4.z); q4.w = fmaf(v.w, v.w, q4.w);
        }
    }
    // fold the pixel lanes into <= 4 smem lanes in a fixed order (lane pl goes to slot pl % 4, rounds in sequence)
    for (int round = 0; round * 4 < PL; ++round) {
        if (pl < PL && pl / 4 == round) {
            float* ps = &ssum[pl & 3][4 * cv];
            float* pq = &ssq[pl & 3][4 * cv];
            if (round == 0) {
                ps[0] = s4.x; ps[1] = s4.y; ps[2] = s4.z; ps[3] = s4.w;
                pq[0] = q4.x; pq[1] = q4.y; pq[2] = q4.z; pq[3] = q4.w;
            } else {
                ps[0] += s4.x; ps[1] += s4.y; ps[2] += s4.z; ps[3] += s4.w;
                pq[0] += q4.x; pq[1] += q4.y; pq[2] += q4.z; pq[3] += q4.w;
            }
        }
        __syncthreads();
    }
    if (tid < 32) {
        // fp64 is slow on this part (a dependent DADD chain of 4 x cg = 12 .. 96 links cost microseconds): one chain per
        // smem lane, joined in lane order
        double sml[4] = {0.0, 0.0, 0.0, 0.0}, sql[4] = {0.0, 0.0, 0.0, 0.0};
        const int lanes = min(P, PL);
        for (int k = 0; k < cg; ++k) {
#pragma unroll
            for (int l = 0; l < 4; ++l)
                if (l < lanes) { sml[l] += (double)ssum[l][tid * cg + k]; sql[l] += (double)ssq[l][tid * cg + k]; }
        }
        part[tid][0] = (sml[0] + sml[1]) + (sml[2] + sml[3]);
        part[tid][1] = (sql[0] + sql[1]) + (sql[2] + sql[3]);
    }
    cluster_sync_all();                                    // partials of all CL CTAs are visible cluster-wide
    if (tid < 32) {
        double sm = 0.0, sq = 0.0;
        const uint32_t local = smem_u32(&part[tid][0]);
        double pa[CL], pb[CL];
#pragma unroll
        for (unsigned r = 0; r < CL; ++r) {                // all remote reads in flight together (they were serialised)
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(pa[r]), "=d"(pb[r]) : "r"(remote));
        }
#pragma unroll
        for (unsigned r = 0; r < CL; ++r) {                // rank order: identical bits in every CTA
            sm += pa[r];
            sq += pb[r];
        }
        const double n = (double)HW * cg;
        const double m = sm / n;
        double v = sq / n - m * m;
        if (v < 0.0) v = 0.0;
        mean[tid] = (float)m;
        rstd[tid] = (float)(1.0 / sqrt(v + 1e-6));
    }
    cluster_sync_all();                                    // nobody leaves (or overwrites part) while peers still read it
    // ---- pass 2: apply + swish over the same slice (resident part from shared memory)
    const int total = (p1 - p0) * cv_n, total_res = res_px * cv_n;
    const float4* xin = base + (size_t)p0 * cv_n;
    uint2* oa = reinterpret_cast<uint2*>(out_act + (size_t)f * HW * C) + (size_t)p0 * cv_n;
    uint2* orw = out_raw ? reinterpret_cast<uint2*>(out_raw + (size_t)f * HW * C) + (size_t)p0 * cv_n : nullptr;
    const int stride = (512 / cv_n) * cv_n;                // threads beyond a whole number of channel periods idle
    if (tid < stride) {
        const int c = 4 * (tid % cv_n);
        const float4 g = reinterpret_cast<const float4*>(gamma)[tid % cv_n];
        const float4 bt = reinterpret_cast<const float4*>(beta)[tid % cv_n];
        const int g0 = c / cg, g1 = (c + 1) / cg, g2 = (c + 2) / cg, g3 = (c + 3) / cg;
        const float ax = rstd[g0] * g.x, ay = rstd[g1] * g.y, az = rstd[g2] * g.z, aw = rstd[g3] * g.w;
        const float bx = fmaf(-mean[g0], ax, bt.x), by = fmaf(-mean[g1], ay, bt.y), bz = fmaf(-mean[g2], az, bt.z),
                    bw = fmaf(-mean[g3], aw, bt.w);
#pragma unroll 8
        for (int i = tid; i < total; i += stride) {
            const float4 v = i < total_res ? gsl[i] : xin[i];
            oa[i] = make_uint2(pack_bf16x2(swishf(fmaf(v.x, ax, bx)), swishf(fmaf(v.y, ay, by))),
                               pack_bf16x2(swishf(fmaf(v.z, az, bz)), swishf(fmaf(v.w, aw, bw))));
            if (orw) orw[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
    }
}

template <int CL>
static int gn_fused_launch_cl(const float* x, int F, int HW, int C, const float* gamma, const float* beta, bf16* out_act,
                              bf16* out_raw, cudaStream_t s) {
    static int ok = -1;                                    // -1 unknown, 0 this cluster size cannot be launched, 1 fine
    if (ok < 0) {
        ok = 1;
        if (CL > 8 && cudaFuncSetAttribute(gn_fused_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            (void)cudaGetLastError();
            ok = 0;
        }
        if (ok && cudaFuncSetAttribute(gn_fused_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSmemCap) != cudaSuccess) {
            (void)cudaGetLastError();
            ok = 0;
        }
    }
    if (!ok) return -1000;
    const int per = (HW + CL - 1) / CL;
    const size_t slice = (size_t)per * C * 4;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(F * CL);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = slice < kGnSmemCap ? slice : kGnSmemCap;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_use(false) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gn_fused_kernel<CL>, x, HW, C, gamma, beta, out_act, out_raw);
    if (e != cudaSuccess && CL > 8) { (void)cudaGetLastError(); ok = 0; return -1000; }
    return (int)e;
}

// out_act = bf16(swish(GN(x))), out_raw = bf16(x) (optional); x: [F][HW][C] fp32
int gn_fused_launch(const float* x, int F, int HW, int C, const float* gamma, const float* beta, bf16* out_act, bf16* out_raw,
                    cudaStream_t s) {
    if (C % 32 || C > 768 || C < 96) return -30;
    int r = gn_fused_launch_cl<16>(x, F, HW, C, gamma, beta, out_act, out_raw, s);
    if (r == -1000) r = gn_fused_launch_cl<8>(x, F, HW, C, gamma, beta, out_act, out_raw, s);
    return r;
}

// ------------------------------------------------------------------------------------------ bilinear helpers
// PyTorch area_pixel_compute_source_index (align_corners=False): src = scale*(dst+0.5)-0.5 clamped at 0.
__device__ __forceinline__ void bil_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.0f) src = 0.0f;
    i0 = (int)src;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
}

// Horizontal source offset / weight of output column xo0 + dx for an exact 2^-N bilinear upscale (xo0 a multiple of
// 2^N): src = (dx + 0.5) / 2^N - 0.5 relative to xo0 >> N.  Compile-time functions of dx.
template <int N>
struct PowX {
    static __device__ __forceinline__ constexpr int off(int dx) {
        const int num = 2 * dx + 1 - (1 << N);
        return num >= 0 ? num / (1 << (N + 1)) : -((-num + (1 << (N + 1)) - 1) / (1 << (N + 1)));
    }
    static __device__ __forceinline__ constexpr float lx(int dx) {
        const int num = 2 * dx + 1 - (1 << N);
        return (float)(num - off(dx) * (1 << (N + 1))) / (float)(1 << (N + 1));
    }
};

// thread = (4-channel vector, output row); produces 24 consecutive output columns with the walk fully unrolled: for
// the x2 upscale the reload points and the 0.25 / 0.75 weights are compile-time constants (index clamping at the edges
// reproduces PyTorch's source-coordinate clamp exactly).
template <int DX>
struct Up2Walk {
    static __device__ __forceinline__ void run(const float4* __restrict__ base, int y0, int y1, float ly, int W, int cv_n,
                                               int xb, float4& colL, float4& colR, uint2* __restrict__ orow) {
        constexpr int o = PowX<1>::off(DX);
        constexpr float lxv = PowX<1>::lx(DX);
        constexpr bool reload = (DX == 0) || (PowX<1>::off(DX) != PowX<1>::off(DX > 0 ? DX - 1 : 0));
        if constexpr (reload) {
            const int xu = xb + o;
            const int x1 = min(max(xu + 1, 0), W - 1);
            const float h0 = 1.0f - ly;
            if constexpr (DX == 0) {
                const int x0 = min(max(xu, 0), W - 1);
                const float4 a = base[((size_t)y0 * W + x0) * cv_n], c = base[((size_t)y1 * W + x0) * cv_n];
                colL = make_float4(h0 * a.x + ly * c.x, h0 * a.y + ly * c.y, h0 * a.z + ly * c.z, h0 * a.w + ly * c.w);
            } else {
                colL = colR;                                   // the walk advanced by exactly one source column
            }
            const float4 a = base[((size_t)y0 * W + x1) * cv_n], c = base[((size_t)y1 * W + x1) * cv_n];
            colR = make_float4(h0 * a.x + ly * c.x, h0 * a.y + ly * c.y, h0 * a.z + ly * c.z, h0 * a.w + ly * c.w);
        }
        constexpr float w0 = 1.0f - lxv;
        orow[(size_t)DX * cv_n] = make_uint2(pack_bf16x2(w0 * colL.x + lxv * colR.x, w0 * colL.y + lxv * colR.y),
                                             pack_bf16x2(w0 * colL.z + lxv * colR.z, w0 * colL.w + lxv * colR.w));
        if constexpr (DX + 1 < 24) Up2Walk<DX + 1>::run(base, y0, y1, ly, W, cv_n, xb, colL, colR, orow);
    }
};

__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, int H, int W, int C,
                                                        bf16* __restrict__ out, int XT) {
    pdl_trigger();
    pdl_wait();
    const int cv_n = C >> 2;
    const int lanes = blockDim.x;                        // vector lanes per row group
    const int cv = blockIdx.z * lanes + threadIdx.x;
    const int yo = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.x / ((2 * W) / XT);
    const int xo0 = (blockIdx.x % ((2 * W) / XT)) * XT;  // XT == 24
    if (cv >= cv_n || yo >= 2 * H) return;
    int y0, y1;
    float ly;
    bil_src(yo, 0.5f, H, y0, y1, ly);
    const float4* base = reinterpret_cast<const float4*>(x + (size_t)f * H * W * C) + cv;
    float4 colL, colR;
    uint2* orow = reinterpret_cast<uint2*>(out) + (((size_t)f * 2 * H + yo) * 2 * W + xo0) * cv_n + cv;
    Up2Walk<0>::run(base, y0, y1, ly, W, cv_n, xo0 >> 1, colL, colR, orow);
}

int upsample2x_launch(const float* x, int F, int H, int W, int C, bf16* out, cudaStream_t s) {
    const int cv_n = C / 4;
    const int lanes = (cv_n % 32 == 0) ? 32 : ((cv_n % 24 == 0) ? 24 : 16);
    if (cv_n % lanes) return -35;
    const int rows = 7;                                  // 2H is 14, 28 or 56
    const int XT = 24;                                   // 2W is 24, 48, 96 or 192: all multiples of 24
    if ((2 * W) % XT || (2 * H) % rows) return -35;
    dim3 grid(F * ((2 * W) / XT), (2 * H) / rows, cv_n / lanes);
    DSB_PDL_LAUNCH(upsample2x_kernel, grid, dim3(lanes, rows), 0, s, x, H, W, C, out, XT);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ LayerNorm over channels
// A token's C channels are read as float4 by LPT lanes (8 / 16 / 32 lanes for C = 96 / 192 / >= 384), so a warp covers
// 4 / 2 / 1 tokens with 48-96 bytes in flight per lane; statistics are two-pass (mean, then centred squares) in fp32.
template <int C>
struct LnGeom {
    static constexpr int V = C / 4;
    static constexpr int LPT = (C == 96) ? 8 : ((C == 192) ? 16 : 32);
    static constexpr int NVEC = V / LPT;
    static constexpr int TPW = 32 / LPT;
};

template <int LPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int C, bool APPLY>
__global__ void __launch_bounds__(256) ln_vec_kernel(const float* __restrict__ x, long tokens, float2* __restrict__ stats,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    bf16* __restrict__ out, int hw, int T, int tmax, int f16) {
    pdl_trigger();
    pdl_wait();
    using G = LnGeom<C>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G::LPT, l = lane % G::LPT;
    const long tok = ((long)blockIdx.x * 8 + warp) * G::TPW + sub;
    const bool live = tok < tokens && (int)((tok / hw) % T) < tmax;
    float4 v[G::NVEC];
    const float4* row = reinterpret_cast<const float4*>(x + (live ? tok : 0) * C);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < G::NVEC; ++i) {
        v[i] = live ? row[l + G::LPT * i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = group_sum<G::LPT>(s) * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < G::NVEC; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(group_sum<G::LPT>(q) * (1.0f / C) + 1e-5f);
    if (!live) return;
    if constexpr (APPLY) {
        uint2* o = reinterpret_cast<uint2*>(out + tok * C);
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
        for (int i = 0; i < G::NVEC; ++i) {
            const float4 g = g4[l + G::LPT * i], b = b4[l + G::LPT * i];
            const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
            o[l + G::LPT * i] = f16 ? make_uint2(pack_f16x2(y0, y1), pack_f16x2(y2, y3))
                                    : make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
        }
    } else {
        if (l == 0) stats[tok] = make_float2(mean, rstd);
    }
}

template <bool APPLY>
static int ln_vec_launch(const float* x, long tokens, int C, float2* stats, const float* gamma, const float* beta, bf16* out,
                         int hw, int T, int tmax, cudaStream_t s, int f16 = 0) {
#define DSB_LN_CASE(CC)                                                                                              \
    case CC: {                                                                                                       \
        const int g = (int)((tokens + 8 * LnGeom<CC>::TPW - 1) / (8 * LnGeom<CC>::TPW));                              \
        DSB_PDL_LAUNCH((ln_vec_kernel<CC, APPLY>), g, 256, 0, s, x, tokens, stats, gamma, beta, out, hw, T, tmax, f16);     \
        break;                                                                                                       \
    }
    switch (C) {
        DSB_LN_CASE(96)
        DSB_LN_CASE(192)
        DSB_LN_CASE(384)
        DSB_LN_CASE(512)
        DSB_LN_CASE(768)
        default: return -31;
    }
#undef DSB_LN_CASE
    DSB_LAUNCH_CHECK();
}

int ln_stats_launch(const float* x, long tokens, int C, float2* stats, int hw, int T, int tmax, cudaStream_t s) {
    return ln_vec_launch<false>(x, tokens, C, stats, nullptr, nullptr, nullptr, hw, T, tmax, s);
}

int ln_apply_launch(const float* x, long tokens, int C, const float* gamma, const float* beta, bf16* out, int hw,
                    int T, int tmax, cudaStream_t s, int f16) {
    return ln_vec_launch<true>(x, tokens, C, nullptr, gamma, beta, out, hw, T, tmax, s, f16);
}

// ------------------------------------------------------------------------------------------ q = LN(dw3x3(LN(x)))
// Wide stages (C = 384, 768; few tokens): one warp per token, a lane owns NVEC float4 channel vectors (LnGeom layout).
// The 9 neighbour rows are read straight from L1/L2 with 128-bit loads; both LayerNorms are warp reductions.
template <int C>
__global__ void __launch_bounds__(256) q_dwln_kernel(const float* __restrict__ x, const float2* __restrict__ stats,
                                                    long tokens, int H, int W, const float* __restrict__ ng,
                                                    const float* __restrict__ nb, const float* __restrict__ wq,
                                                    const float* __restrict__ qg, const float* __restrict__ qb,
                                                    bf16* __restrict__ out, int T, int tmax) {
    pdl_trigger();
    pdl_wait();
    using G = LnGeom<C>;
    static_assert(G::LPT == 32, "one token per warp");
    constexpr int NV = G::NVEC, CV = C / 4;
    const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tok >= tokens) return;
    const int hw = H * W;
    const long f = tok / hw;
    if ((int)(f % T) >= tmax) return;
    const int pix = (int)(tok % hw);
    const int y = pix / W, xx = pix % W;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* w4 = reinterpret_cast<const float4*>(wq);
    float4 q[NV], wsum[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { q[i] = make_float4(0.f, 0.f, 0.f, 0.f); wsum[i] = q[i]; }
    // conv(LN(x)) = g * sum_k w_k * n_k + b * sum_{k in image} w_k  with  n_k = (x_k - mean_k) * rstd_k
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xn = xx + k % 3 - 1;
        if (yy < 0 || yy >= H || xn < 0 || xn >= W) continue;          // warp-uniform
        const long nt = f * hw + (long)yy * W + xn;
        const float2 st = stats[nt];
        const float a = st.y, c0 = -st.x * st.y;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 v = x4[nt * CV + lane + 32 * i];
            const float4 w = __ldg(w4 + k * CV + lane + 32 * i);
            q[i].x = fmaf(w.x, fmaf(v.x, a, c0), q[i].x); q[i].y = fmaf(w.y, fmaf(v.y, a, c0), q[i].y);
            q[i].z = fmaf(w.z, fmaf(v.z, a, c0), q[i].z); q[i].w = fmaf(w.w, fmaf(v.w, a, c0), q[i].w);
            wsum[i].x += w.x; wsum[i].y += w.y; wsum[i].z += w.z; wsum[i].w += w.w;
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(ng) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(nb) + lane + 32 * i);
        q[i].x = fmaf(g.x, q[i].x, b.x * wsum[i].x); q[i].y = fmaf(g.y, q[i].y, b.y * wsum[i].y);
        q[i].z = fmaf(g.z, q[i].z, b.z * wsum[i].z); q[i].w = fmaf(g.w, q[i].w, b.w * wsum[i].w);
        s += (q[i].x + q[i].y) + (q[i].z + q[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float v2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = q[i].x - mean, b = q[i].y - mean, c = q[i].z - mean, d = q[i].w - mean;
        v2 = fmaf(a, a, v2); v2 = fmaf(b, b, v2); v2 = fmaf(c, c, v2); v2 = fmaf(d, d, v2);
    }
    const float rstd = rsqrtf(warp_sum(v2) * (1.0f / C) + 1e-5f);
    uint2* o = reinterpret_cast<uint2*>(out + tok * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(qg) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(qb) + lane + 32 * i);
        o[lane + 32 * i] = make_uint2(pack_bf16x2((q[i].x - mean) * rstd * g.x + b.x, (q[i].y - mean) * rstd * g.y + b.y),
                                      pack_bf16x2((q[i].z - mean) * rstd * g.z + b.z, (q[i].w - mean) * rstd * g.w + b.w));
    }
}

// Shared-memory tiled variant for the narrow, token-rich stages (C = 96, 192).  A block owns XT tokens of one image row
// and stages the LayerNormed inputs of the 3 x (XT + 2) neighbourhood once (normalised while loading; zero outside the
// image = the conv's zero padding of LN(x)).  Both phases use the LnGeom layout: LPT lanes x 3 float4 cover a token,
// so a warp works on 4 (C = 96) or 2 (C = 192) tokens at a time with 128-bit shared/global accesses and the channel
// LayerNorm of the conv output is a shuffle reduction inside the LPT lanes.  (The scalar one-token-per-warp version was
// issue-bound: ~150 warp instructions per token against ~50 here.)
template <int C, int XT>
__global__ void __launch_bounds__(256) q_dwln_tiled_kernel(const float* __restrict__ x, const float2* __restrict__ stats,
                                                          int H, int W, const float* __restrict__ ng,
                                                          const float* __restrict__ nb, const float* __restrict__ wq,
                                                          const float* __restrict__ qg, const float* __restrict__ qb,
                                                          bf16* __restrict__ out, int T, int tmax) {
    pdl_trigger();
    pdl_wait();
    using G = LnGeom<C>;
    static_assert(G::NVEC == 3 && XT == 8 * G::TPW, "tile geometry");
    constexpr int TW = XT + 2, CV = C / 4, NT = 3 * TW, PER = 8 * G::TPW, NPASS = (NT + PER - 1) / PER;
    extern __shared__ float4 tile4[];                    // [3][TW][CV]
    const int nseg = W / XT;
    const int seg = blockIdx.x % nseg;
    const int y = (blockIdx.x / nseg) % H;
    const int f = blockIdx.x / (nseg * H);
    if (f % T >= tmax) return;
    const int x0 = seg * XT;
    const size_t fbase = (size_t)f * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G::LPT, l = lane % G::LPT;
    const int slot = warp * G::TPW + sub;                // token slot of this lane group, 0 .. PER-1
    const float4* x4 = reinterpret_cast<const float4*>(x);
    {
        float4 v[NPASS][3];
        float2 st[NPASS];
        bool ok[NPASS];
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
            const int t = slot + p * PER;
            const int r = t / TW, col = t % TW;
            const int yy = y + r - 1, xx = x0 + col - 1;
            ok[p] = t < NT && yy >= 0 && yy < H && xx >= 0 && xx < W;
            const size_t tok = ok[p] ? fbase + (size_t)yy * W + xx : fbase;
            st[p] = stats[tok];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[p][i] = x4[tok * CV + l + G::LPT * i];
        }
        float4 g[3], b[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g[i] = reinterpret_cast<const float4*>(ng)[l + G::LPT * i];
            b[i] = reinterpret_cast<const float4*>(nb)[l + G::LPT * i];
        }
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
            const int t = slot + p * PER;
            if (t < NT) {
                const float m = st[p].x, rs = ok[p] ? st[p].y : 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok[p])
                        o = make_float4((v[p][i].x - m) * rs * g[i].x + b[i].x, (v[p][i].y - m) * rs * g[i].y + b[i].y,
                                        (v[p][i].z - m) * rs * g[i].z + b[i].z, (v[p][i].w - m) * rs * g[i].w + b[i].w);
                    tile4[t * CV + l + G::LPT * i] = o;
                }
            }
        }
    }
    __syncthreads();
    const float4* w4 = reinterpret_cast<const float4*>(wq);
    float4 q[3];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int cvi = l + G::LPT * i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 tv = tile4[(r * TW + slot + k) * CV + cvi];
                const float4 wv = __ldg(w4 + (r * 3 + k) * CV + cvi);
                a.x = fmaf(wv.x, tv.x, a.x); a.y = fmaf(wv.y, tv.y, a.y);
                a.z = fmaf(wv.z, tv.z, a.z); a.w = fmaf(wv.w, tv.w, a.w);
            }
        q[i] = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    const float mean = group_sum<G::LPT>(s) * (1.0f / C);
    float v2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = q[i].x - mean, b = q[i].y - mean, c = q[i].z - mean, d = q[i].w - mean;
        v2 = fmaf(a, a, v2); v2 = fmaf(b, b, v2); v2 = fmaf(c, c, v2); v2 = fmaf(d, d, v2);
    }
    const float rstd = rsqrtf(group_sum<G::LPT>(v2) * (1.0f / C) + 1e-5f);
    uint2* o = reinterpret_cast<uint2*>(out + (fbase + (size_t)y * W + x0 + slot) * C);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int cvi = l + G::LPT * i;
        const float4 gq = __ldg(reinterpret_cast<const float4*>(qg) + cvi), bq = __ldg(reinterpret_cast<const float4*>(qb) + cvi);
        o[cvi] = make_uint2(pack_bf16x2((q[i].x - mean) * rstd * gq.x + bq.x, (q[i].y - mean) * rstd * gq.y + bq.y),
                            pack_bf16x2((q[i].z - mean) * rstd * gq.z + bq.z, (q[i].w - mean) * rstd * gq.w + bq.w));
    }
}

// Second generation of the tiled variant (C = 96, 192): a block owns a TH x TW pixel tile of one frame.
//  * phase 1 stages the plain normalised values n = (x - mean) * rstd of the (TH + 2) x (TW + 2) neighbourhood (one FFMA
//    per element; zero outside the image = the conv's zero padding of LN(x)).  The LayerNorm affine is folded out of the
//    tile:  conv(g * n + b) = sum_k (w_k g) n_k + sum_{k in image} w_k b, with the tables wg = w * g, wb = w * b and
//    wbs = sum_k wb_k prepared once per weight set (q_dw_prep_launch);
//  * phase 2: a lane group (LPT lanes x 3 float4 = one token's channels) owns a tile column and walks its TH rows with the
//    nine taps of one channel vector held in registers, so tap loads, index math and the block's fixed costs are
//    amortised over TH tokens; the out-of-image taps of border tokens are subtracted from wbs.
// Against the first version (one token per lane group and block row: 3.2 staged tokens and 27 tap loads per output token,
// three FP operations per staged element) this stages 1.6 - 2.1 tokens per output token at a third of the arithmetic.
template <int C, int TW, int TH, int MINB, int PB>
__global__ void __launch_bounds__(256, MINB) q_dwln_tile2_kernel(const float* __restrict__ x, const float2* __restrict__ stats,
                                                             int H, int W, const float* __restrict__ wg,
                                                             const float* __restrict__ wb, const float* __restrict__ wbs,
                                                             const float* __restrict__ qg, const float* __restrict__ qb,
                                                             bf16* __restrict__ out, int T, int tmax) {
    pdl_trigger();
    pdl_wait();
    using G = LnGeom<C>;
    static_assert(G::NVEC == 3 && TW == 8 * G::TPW, "tile geometry: one lane group per tile column");
    constexpr int HWT = TW + 2, HHT = TH + 2, CV = C / 4, NT = HHT * HWT, PER = 8 * G::TPW, NPASS = (NT + PER - 1) / PER;
    extern __shared__ float4 tile4[];                    // [HHT][HWT][CV]
    const int nseg = W / TW, nrow = H / TH;
    const int seg = blockIdx.x % nseg;
    const int yb = (blockIdx.x / nseg) % nrow;
    const int f = blockIdx.x / (nseg * nrow);
    if (f % T >= tmax) return;
    const int x0 = seg * TW, y0 = yb * TH;
    const size_t fbase = (size_t)f * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G::LPT, l = lane % G::LPT;
    const int slot = warp * G::TPW + sub;                // tile column of this lane group, 0 .. TW-1
    const float4* x4 = reinterpret_cast<const float4*>(x);
    // ---- phase 1: normalised neighbourhood -> shared memory (batches of PB tokens per lane group keep 3 PB loads in flight)
#pragma unroll 1
    for (int p0 = 0; p0 < NPASS; p0 += PB) {
        float4 v[PB][3];
        float2 st[PB];
        bool ok[PB];
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            const int t = slot + (p0 + p) * PER;
            const int r = t / HWT, col = t - r * HWT;
            const int yy = y0 + r - 1, xx = x0 + col - 1;
            ok[p] = t < NT && yy >= 0 && yy < H && xx >= 0 && xx < W;
            const size_t tok = ok[p] ? fbase + (size_t)yy * W + xx : fbase;
            st[p] = stats[tok];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[p][i] = x4[tok * CV + l + G::LPT * i];
        }
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            const int t = slot + (p0 + p) * PER;
            if (t < NT) {
                const float rs = ok[p] ? st[p].y : 0.0f, c0 = -st[p].x * rs;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    tile4[t * CV + l + G::LPT * i] = make_float4(fmaf(v[p][i].x, rs, c0), fmaf(v[p][i].y, rs, c0),
                                                                 fmaf(v[p][i].z, rs, c0), fmaf(v[p][i].w, rs, c0));
            }
        }
    }
    __syncthreads();
    // ---- phase 2: depthwise 3x3 down the column, then the channel LayerNorm of each of the TH tokens
    const float4* wg4 = reinterpret_cast<const float4*>(wg);
    const float4* wb4 = reinterpret_cast<const float4*>(wb);
    const int xx = x0 + slot;
    float4 q[TH][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int cvi = l + G::LPT * i;
        float4 w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = __ldg(wg4 + k * CV + cvi);
        const float4 bs = __ldg(reinterpret_cast<const float4*>(wbs) + cvi);
#pragma unroll
        for (int r = 0; r < TH; ++r) {
            float4 a = bs;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 tv = tile4[((r + dy) * HWT + slot + dx) * CV + cvi];
                    const float4 wv = w[dy * 3 + dx];
                    a.x = fmaf(wv.x, tv.x, a.x); a.y = fmaf(wv.y, tv.y, a.y);
                    a.z = fmaf(wv.z, tv.z, a.z); a.w = fmaf(wv.w, tv.w, a.w);
                }
            const int yy = y0 + r;
            if (yy == 0 || yy == H - 1 || xx == 0 || xx == W - 1) {          // border token: drop the out-of-image bias taps
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int y2 = yy + dy - 1, x2 = xx + dx - 1;
                        if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) {
                            const float4 t = __ldg(wb4 + (dy * 3 + dx) * CV + cvi);
                            a.x -= t.x; a.y -= t.y; a.z -= t.z; a.w -= t.w;
                        }
                    }
            }
            q[r][i] = a;
        }
    }
    float4 gq[3], bq[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        gq[i] = __ldg(reinterpret_cast<const float4*>(qg) + l + G::LPT * i);
        bq[i] = __ldg(reinterpret_cast<const float4*>(qb) + l + G::LPT * i);
    }
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) s += (q[r][i].x + q[r][i].y) + (q[r][i].z + q[r][i].w);
        const float mean = group_sum<G::LPT>(s) * (1.0f / C);
        float v2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float a = q[r][i].x - mean, b = q[r][i].y - mean, c = q[r][i].z - mean, d = q[r][i].w - mean;
            v2 = fmaf(a, a, v2); v2 = fmaf(b, b, v2); v2 = fmaf(c, c, v2); v2 = fmaf(d, d, v2);
        }
        const float rstd = rsqrtf(group_sum<G::LPT>(v2) * (1.0f / C) + 1e-5f);
        uint2* o = reinterpret_cast<uint2*>(out + (fbase + (size_t)(y0 + r) * W + xx) * C);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            o[l + G::LPT * i] = make_uint2(
                pack_bf16x2((q[r][i].x - mean) * rstd * gq[i].x + bq[i].x, (q[r][i].y - mean) * rstd * gq[i].y + bq[i].y),
                pack_bf16x2((q[r][i].z - mean) * rstd * gq[i].z + bq[i].z, (q[r][i].w - mean) * rstd * gq[i].w + bq[i].w));
    }
}

__device__ __forceinline__ void pooled_ln_store(const float* pre, int C, const float* __restrict__ g,
                                                const float* __restrict__ b, bf16* __restrict__ o, float* red,
                                                const ScoreBias* sbp = nullptr, int tokv = 0);

// Q and V producers of a narrow stage in ONE pass over the stage input (audio-visual mode, C = 96 / 192):
//   q = LN_q(dw3x3(LN(x)))  for every pixel,   v = LN_v(dw SxS stride S (LN(x)))  for the 18 pooling windows of the frame.
// A block owns a TH x TW pixel tile made of whole pooling windows (16 x 16 = one window at C = 96, 8 x 16 = two windows
// at C = 192).  Phase 1 computes the LayerNorm statistics of every staged token itself (a lane group holds the whole
// token; same two-pass fp32 formula as ln_vec_kernel) and stages n = (x - mean) * rstd; phase 2 is q_dwln_tile2's column
// walk (a lane group owns RS = 4 rows of one column); phase 3 pools the windows straight from the staged tile with the
// LayerNorm affine folded into the tap table (wvg = w_v * g, wvbs = b * sum_p w_v).  Replaces ln_stats + q_dwln +
// pool_ln(v): the stage input is read once (x 1.3 - 1.4 halo) instead of three times, and two launches disappear.
template <int C, int TW, int TH, int S_>
__global__ void __launch_bounds__(512, 1) qv_tile_kernel(const float* __restrict__ x, int H, int W,
                                                        const float* __restrict__ wg, const float* __restrict__ wb,
                                                        const float* __restrict__ wbs, const float* __restrict__ qg,
                                                        const float* __restrict__ qb, const float* __restrict__ wvg,
                                                        const float* __restrict__ wvbs, const float* __restrict__ vg,
                                                        const float* __restrict__ vb, bf16* __restrict__ q_out,
                                                        bf16* __restrict__ v_out, int T, int tmax) {
    pdl_trigger();
    pdl_wait();
    using G = LnGeom<C>;
    constexpr int RS = 4, NG = 512 / G::LPT, HWT = TW + 2, HHT = TH + 2, CV = C / 4, NT = HHT * HWT;
    constexpr int NPASS = (NT + NG - 1) / NG, NWIN = (TW / S_) * (TH / S_), PG = 512 / CV;
    static_assert(G::NVEC == 3 && NG == TW * (TH / RS) && TH % S_ == 0 && TW % S_ == 0, "tile geometry");
    extern __shared__ float4 tile4[];                    // [HHT][HWT][CV] normalised tokens
    __shared__ float pool[PG * C];
    __shared__ float red[32];
    const int nseg = W / TW, nrow = (H + TH - 1) / TH;
    const int seg = blockIdx.x % nseg;
    const int yb = (blockIdx.x / nseg) % nrow;
    const int f = blockIdx.x / (nseg * nrow);
    if (f % T >= tmax) return;
    const int x0 = seg * TW, y0 = yb * TH;
    const size_t fbase = (size_t)f * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G::LPT, l = lane % G::LPT;
    const int gidx = warp * G::TPW + sub;                // lane group, 0 .. NG-1
    const float4* x4 = reinterpret_cast<const float4*>(x);
    // ---- phase 1: statistics + normalised neighbourhood -> shared memory
#pragma unroll 1
    for (int p0 = 0; p0 < NPASS; p0 += 3) {
        float4 v[3][3];
        bool ok[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int t = gidx + (p0 + p) * NG;
            const int r = t / HWT, col = t - r * HWT;
            const int yy = y0 + r - 1, xx = x0 + col - 1;
            ok[p] = (p0 + p) < NPASS && t < NT && yy >= 0 && yy < H && xx >= 0 && xx < W;
            const size_t tok = ok[p] ? fbase + (size_t)yy * W + xx : fbase;
#pragma unroll
            for (int i = 0; i < 3; ++i) v[p][i] = x4[tok * CV + l + G::LPT * i];
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int t = gidx + (p0 + p) * NG;
            float s1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) s1 += (v[p][i].x + v[p][i].y) + (v[p][i].z + v[p][i].w);
            const float mean = group_sum<G::LPT>(s1) * (1.0f / C);
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float a = v[p][i].x - mean, b = v[p][i].y - mean, c = v[p][i].z - mean, d = v[p][i].w - mean;
                s2 = fmaf(a, a, s2); s2 = fmaf(b, b, s2); s2 = fmaf(c, c, s2); s2 = fmaf(d, d, s2);
            }
            const float rstd = rsqrtf(group_sum<G::LPT>(s2) * (1.0f / C) + 1e-5f);
            if ((p0 + p) < NPASS && t < NT) {
                const float rs = ok[p] ? rstd : 0.0f, c0 = -mean * rs;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    tile4[t * CV + l + G::LPT * i] = make_float4(fmaf(v[p][i].x, rs, c0), fmaf(v[p][i].y, rs, c0),
                                                                 fmaf(v[p][i].z, rs, c0), fmaf(v[p][i].w, rs, c0));
            }
        }
    }
    __syncthreads();
    // ---- phase 2: q for RS rows of one tile column per lane group
    {
        const float4* wg4 = reinterpret_cast<const float4*>(wg);
        const float4* wb4 = reinterpret_cast<const float4*>(wb);
        const int col = gidx % TW, r0 = (gidx / TW) * RS;
        const int xx = x0 + col;
        float4 q[RS][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int cvi = l + G::LPT * i;
            float4 w[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k] = __ldg(wg4 + k * CV + cvi);
            const float4 bs = __ldg(reinterpret_cast<const float4*>(wbs) + cvi);
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                float4 a = bs;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 tv = tile4[((r0 + r + dy) * HWT + col + dx) * CV + cvi];
                        const float4 wv = w[dy * 3 + dx];
                        a.x = fmaf(wv.x, tv.x, a.x); a.y = fmaf(wv.y, tv.y, a.y);
                        a.z = fmaf(wv.z, tv.z, a.z); a.w = fmaf(wv.w, tv.w, a.w);
                    }
                const int yy = y0 + r0 + r;
                if (yy == 0 || yy == H - 1 || xx == 0 || xx == W - 1) {      // border token: drop the out-of-image bias taps
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int y2 = yy + dy - 1, x2 = xx + dx - 1;
                            if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) {
                                const float4 t = __ldg(wb4 + (dy * 3 + dx) * CV + cvi);
                                a.x -= t.x; a.y -= t.y; a.z -= t.z; a.w -= t.w;
                            }
                        }
                }
                q[r][i] = a;
            }
        }
        float4 gq[3], bq[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gq[i] = __ldg(reinterpret_cast<const float4*>(qg) + l + G::LPT * i);
            bq[i] = __ldg(reinterpret_cast<const float4*>(qb) + l + G::LPT * i);
        }
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            float s1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) s1 += (q[r][i].x + q[r][i].y) + (q[r][i].z + q[r][i].w);
            const float mean = group_sum<G::LPT>(s1) * (1.0f / C);
            float v2 = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float a = q[r][i].x - mean, b = q[r][i].y - mean, c = q[r][i].z - mean, d = q[r][i].w - mean;
                v2 = fmaf(a, a, v2); v2 = fmaf(b, b, v2); v2 = fmaf(c, c, v2); v2 = fmaf(d, d, v2);
            }
            const float rstd = rsqrtf(group_sum<G::LPT>(v2) * (1.0f / C) + 1e-5f);
            const int yy = y0 + r0 + r;
            if (yy < H) {
                uint2* o = reinterpret_cast<uint2*>(q_out + (fbase + (size_t)yy * W + xx) * C);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    o[l + G::LPT * i] = make_uint2(
                        pack_bf16x2((q[r][i].x - mean) * rstd * gq[i].x + bq[i].x, (q[r][i].y - mean) * rstd * gq[i].y + bq[i].y),
                        pack_bf16x2((q[r][i].z - mean) * rstd * gq[i].z + bq[i].z, (q[r][i].w - mean) * rstd * gq[i].w + bq[i].w));
            }
        }
    }
    // ---- phase 3: the pooling windows of this tile (rows below the last whole window are not pooled: 3 x 6 windows)
    if (y0 / S_ >= 3) return;                            // block-uniform
    const int tid = threadIdx.x;
    const int c4 = tid % CV, gq_ = tid / CV;
    const float4* wv4 = reinterpret_cast<const float4*>(wvg) + c4;
#pragma unroll 1
    for (int wdw = 0; wdw < NWIN; ++wdw) {
        const int wy = wdw / (TW / S_), wx = wdw % (TW / S_);
        const int Y = y0 / S_ + wy, X = x0 / S_ + wx;
        if (Y >= 3) break;                               // block-uniform
        if (gq_ < PG) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int p = gq_; p < S_ * S_; p += PG) {
                const int dy = p / S_, dx = p % S_;
                const float4 tv = tile4[((wy * S_ + dy + 1) * HWT + wx * S_ + dx + 1) * CV + c4];
                const float4 w = __ldg(wv4 + p * CV);
                acc.x = fmaf(w.x, tv.x, acc.x); acc.y = fmaf(w.y, tv.y, acc.y);
                acc.z = fmaf(w.z, tv.z, acc.z); acc.w = fmaf(w.w, tv.w, acc.w);
            }
            reinterpret_cast<float4*>(pool)[gq_ * CV + c4] = acc;
        }
        __syncthreads();
        for (int c = tid; c < C; c += 512) {
            float a = __ldg(wvbs + c);
            for (int k = 0; k < PG; ++k) a += pool[k * C + c];
            pool[c] = a;                                  // element c of row 0 is only touched by this thread
        }
        __syncthreads();
        pooled_ln_store(pool, C, vg, vb, v_out + ((size_t)f * 18 + Y * 6 + X) * C, red);
        __syncthreads();
    }
}

// taps with the LayerNorm affine folded in for a depthwise SxS pooling: wg[p][c] = w[p][c] * g[c], wbs[c] = b[c] * sum_p w[p][c]
__global__ void dw_affine_prep_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ b,
                                      int taps, int C, float* __restrict__ wg, float* __restrict__ wbs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.0f;
    for (int k = 0; k < taps; ++k) {
        const float wv = w[k * C + c];
        wg[k * C + c] = wv * g[c];
        s += wv;
    }
    wbs[c] = s * b[c];
}

int dw_affine_prep_launch(const float* w, const float* g, const float* b, int taps, int C, float* wg, float* wbs, cudaStream_t s) {
    dw_affine_prep_kernel<<<(C + 127) / 128, 128, 0, s>>>(w, g, b, taps, C, wg, wbs);
    DSB_LAUNCH_CHECK();
}

template <int C, int TW, int TH, int S_>
static int qv_tile_launch_t(const float* x, int F, int H, int W, const QdwTables& tb, const float* qg, const float* qb,
                            const float* wvg, const float* wvbs, const float* vg, const float* vb, bf16* q_out, bf16* v_out,
                            int T, int tmax, cudaStream_t s) {
    constexpr size_t smem = (size_t)(TH + 2) * (TW + 2) * C * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(qv_tile_kernel<C, TW, TH, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const int grid = F * ((H + TH - 1) / TH) * (W / TW);
    DSB_PDL_LAUNCH((qv_tile_kernel<C, TW, TH, S_>), grid, 512, smem, s, x, H, W, tb.wg, tb.wb, tb.wbs, qg, qb, wvg, wvbs, vg, vb,
                   q_out, v_out, T, tmax);
    DSB_LAUNCH_CHECK();
}

int qv_tile_launch(const float* x, int F, int H, int W, int C, int s_, const QdwTables& tb, const float* qg, const float* qb,
                   const float* wvg, const float* wvbs, const float* vg, const float* vb, bf16* q_out, bf16* v_out, int T,
                   int tmax, cudaStream_t s) {
    if (C == 96 && H == 56 && W == 96 && s_ == 16)
        return qv_tile_launch_t<96, 16, 16, 16>(x, F, H, W, tb, qg, qb, wvg, wvbs, vg, vb, q_out, v_out, T, tmax, s);
    if (C == 192 && H == 28 && W == 48 && s_ == 8)
        return qv_tile_launch_t<192, 16, 8, 8>(x, F, H, W, tb, qg, qb, wvg, wvbs, vg, vb, q_out, v_out, T, tmax, s);
    return -37;
}

// wg[k][c] = w[k][c] * g[c], wb[k][c] = w[k][c] * b[c], wbs[c] = sum_k wb[k][c]   (w: [9][C] depthwise taps)
__global__ void q_dw_prep_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ b, int C,
                                 float* __restrict__ wg, float* __restrict__ wb, float* __restrict__ wbs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.0f;
    for (int k = 0; k < 9; ++k) {
        const float wv = w[k * C + c];
        wg[k * C + c] = wv * g[c];
        const float t = wv * b[c];
        wb[k * C + c] = t;
        s += t;
    }
    wbs[c] = s;
}

int q_dw_prep_launch(const float* w9, const float* ng, const float* nb, int C, float* wg, float* wb, float* wbs, cudaStream_t s) {
    q_dw_prep_kernel<<<(C + 127) / 128, 128, 0, s>>>(w9, ng, nb, C, wg, wb, wbs);
    DSB_LAUNCH_CHECK();
}

template <int C, int TW, int TH, int MINB, int PB>
static int q_dwln_tile2_launch(const float* x, const float2* stats, int F, int H, int W, const QdwTables& tb, const float* qg,
                               const float* qb, bf16* out, int T, int tmax, cudaStream_t s) {
    constexpr size_t smem = (size_t)(TH + 2) * (TW + 2) * C * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(q_dwln_tile2_kernel<C, TW, TH, MINB, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    DSB_PDL_LAUNCH((q_dwln_tile2_kernel<C, TW, TH, MINB, PB>), F * (H / TH) * (W / TW), 256, smem, s, x, stats, H, W, tb.wg, tb.wb, tb.wbs, qg,
                   qb, out, T, tmax);
    DSB_LAUNCH_CHECK();
}

int q_dwln_launch(const float* x, const float2* stats, int F, int H, int W, int C, const float* ng, const float* nb,
                  const float* wq, const QdwTables* tb, const float* qg, const float* qb, bf16* out, int T, int tmax,
                  cudaStream_t s) {
    const long tokens = (long)F * H * W;
    const int g = (int)((tokens + 7) / 8);
    static const int th = [] { const char* e = getenv("DSB_QDW_TH"); return e ? atoi(e) : 4; }();
    if (tb && tb->wg && C == 96 && W % 32 == 0 && H % 4 == 0) {
        if (th == 2) return q_dwln_tile2_launch<96, 32, 2, 3, 2>(x, stats, F, H, W, *tb, qg, qb, out, T, tmax, s);
        if (th == 4) return q_dwln_tile2_launch<96, 32, 4, 2, 4>(x, stats, F, H, W, *tb, qg, qb, out, T, tmax, s);
    }
    if (tb && tb->wg && C == 192 && W % 16 == 0 && H % 4 == 0) {
        if (th == 2) return q_dwln_tile2_launch<192, 16, 2, 3, 2>(x, stats, F, H, W, *tb, qg, qb, out, T, tmax, s);
        if (th == 4) return q_dwln_tile2_launch<192, 16, 4, 2, 4>(x, stats, F, H, W, *tb, qg, qb, out, T, tmax, s);
    }
    if (C == 96 && W % 32 == 0) {
        DSB_PDL_LAUNCH((q_dwln_tiled_kernel<96, 32>), F * H * (W / 32), 256, 3 * 34 * 96 * sizeof(float), s, x, stats, H, W, ng, nb, wq, qg, qb, out, T, tmax);
        DSB_LAUNCH_CHECK();
    }
    if (C == 192 && W % 16 == 0) {
        DSB_PDL_LAUNCH((q_dwln_tiled_kernel<192, 16>), F * H * (W / 16), 256, 3 * 18 * 192 * sizeof(float), s, x, stats, H, W, ng, nb, wq, qg, qb, out, T, tmax);
        DSB_LAUNCH_CHECK();
    }
    switch (C) {
        case 384: DSB_PDL_LAUNCH((q_dwln_kernel<384>), g, 256, 0, s, x, stats, tokens, H, W, ng, nb, wq, qg, qb, out, T, tmax); break;
        case 768: DSB_PDL_LAUNCH((q_dwln_kernel<768>), g, 256, 0, s, x, stats, tokens, H, W, ng, nb, wq, qg, qb, out, T, tmax); break;
        default: return -31;     // C = 96 / 192 need W % 32 / W % 16 == 0 (tiled kernels above)
    }
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ pooled K / V tokens
// Tail shared by both pooling kernels: pre[C] (smem) -> LayerNorm -> bf16 row.
__device__ __forceinline__ void pooled_ln_store(const float* pre, int C, const float* __restrict__ g,
                                                const float* __restrict__ b, bf16* __restrict__ o, float* red,
                                                const ScoreBias* sbp, int tokv) {
    float s = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += pre[c];
    const float mean = block_sum(s, red) / (float)C;
    float q = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { const float d = pre[c] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(block_sum(q, red) / (float)C + 1e-5f);
    float d0 = 0.0f, d1 = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const bf16 y = __float2bfloat16((pre[c] - mean) * rstd * g[c] + b[c]);
        o[c] = y;
        if (sbp && sbp->sb) {                             // folded score bias of this key (see "projections folded into K / V")
            const float yf = __bfloat162float(y);
            d0 = fmaf(yf, __ldg(sbp->mb + c), d0);
            d1 = fmaf(yf, __ldg(sbp->mb + C + c), d1);
        }
    }
    if (sbp && sbp->sb) {
        d0 = block_sum(d0, red);
        d1 = block_sum(d1, red);
        if (threadIdx.x == 0) {
            const int f = tokv / 18, j = tokv - f * 18;
            sbp->sb[(size_t)f * sbp->R + j] = d0 + __ldg(sbp->cb);
            sbp->sb[(size_t)f * sbp->R + 18 + j] = d1 + __ldg(sbp->cb + 1);
        }
    }
}

// part[G][C] (one partial row per pixel group) -> pre[C] in part[0], groups added in index order
__device__ __forceinline__ void pool_fold_groups(float* sm, int C, int G) {
    const int tid = threadIdx.x;
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        float a = sm[c];
        for (int k = 1; k < G; ++k) a += sm[k * C + c];
        sm[c] = a;                                        // element c of row 0 is only touched by this thread
    }
    __syncthreads();
}

// one block per pooled token (f, Y, X)
__global__ void pool_ln_kernel(const float* __restrict__ x, const float2* __restrict__ stats, int H, int W, int C,
                               int s_, const float* __restrict__ ng, const float* __restrict__ nb,
                               const float* __restrict__ wv, const float* __restrict__ vg,
                               const float* __restrict__ vb, bf16* __restrict__ out, int T, int tmax, const ScoreBias sbv) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];          // part[G][C]; pre[C] aliases part[0]
    __shared__ float red[32];
    const int tokv = blockIdx.x;           // f*18 + Y*6 + X
    const int f = tokv / 18, Y = (tokv % 18) / 6, X = tokv % 6;
    if (f % T >= tmax) return;
    const size_t fbase = (size_t)f * H * W;
    const int sl = 31 - __clz(s_);                       // s is a power of two (2, 4, 8, 16)
    const int CV = C >> 2, tid = threadIdx.x;
    const int G = blockDim.x / CV;                       // pixel groups; thread = (4-channel vector, pixel group)
    const int c4 = tid % CV, gq = tid / CV;
    const float4* xw = reinterpret_cast<const float4*>(x + (fbase + (size_t)(Y * s_) * W + X * s_) * C) + c4;
    const float2* sw = stats + fbase + (size_t)(Y * s_) * W + X * s_;
    const float4* w4 = reinterpret_cast<const float4*>(wv) + c4;
    if (gq < G) {
        const float4 g = reinterpret_cast<const float4*>(ng)[c4], bb = reinterpret_cast<const float4*>(nb)[c4];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int npx = s_ * s_;
#pragma unroll 8
        for (int p = gq; p < npx; p += G) {
            const int t = (p >> sl) * W + (p & (s_ - 1));
            const float2 st = sw[t];
            const float4 v = xw[(size_t)t * CV];
            const float4 w = __ldg(w4 + p * CV);
            acc.x = fmaf(w.x, (v.x - st.x) * st.y * g.x + bb.x, acc.x);
            acc.y = fmaf(w.y, (v.y - st.x) * st.y * g.y + bb.y, acc.y);
            acc.z = fmaf(w.z, (v.z - st.x) * st.y * g.z + bb.z, acc.z);
            acc.w = fmaf(w.w, (v.w - st.x) * st.y * g.w + bb.w, acc.w);
        }
        reinterpret_cast<float4*>(sm)[gq * CV + c4] = acc;
    }
    pool_fold_groups(sm, C, G);
    pooled_ln_store(sm, C, vg, vb, out + (size_t)tokv * C, red, &sbv, tokv);
}

// (C / 4) channel vectors x pixel groups; 256-thread blocks keep >= 5 blocks resident per SM (768-thread blocks were
// register-limited to one per SM and ran ~9 serial waves of sync-heavy work)
static int pool_threads(int C) { (void)C; return 256; }

int pool_ln_launch(const float* x, const float2* stats, int F, int H, int W, int C, int s_, const float* ng,
                   const float* nb, const float* wv, const float* vg, const float* vb, bf16* out, int T, int tmax,
                   cudaStream_t s, ScoreBias sbv) {
    const int nthr = pool_threads(C);
    if (C % 4 || C / 4 > nthr) return -34;
    const size_t smem = (size_t)(nthr / (C / 4)) * C * sizeof(float);
    DSB_PDL_LAUNCH(pool_ln_kernel, F * 18, nthr, smem, s, x, stats, H, W, C, s_, ng, nb, wv, vg, vb, out, T, tmax, sbv);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ audio gate
// one block per (y, b, 32-channel chunk): m[x][c] = mean_t a*x ; softmax over x ; g written as [b][c][y][x]
__global__ void __launch_bounds__(256) av_gate_kernel(const float* __restrict__ x, const float* __restrict__ a_low,
                                                     int T, int H, int W, int C, int rshift, float* __restrict__ g) {
    pdl_trigger();
    pdl_wait();
    __shared__ float m[96 * 33];           // [W][33]
    const int y = blockIdx.x, b = blockIdx.y, c0 = blockIdx.z * 32;
    const float invT = 1.0f / (float)T;
    const int CV = C >> 2;
    // thread = (pixel, 4-channel vector): 8 lanes cover the block's 32 channels with 128-bit loads
    for (int e = threadIdx.x; e < W * 8; e += 256) {
        const int c4 = e & 7, xx = e >> 3;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* xp = reinterpret_cast<const float4*>(x) + ((((size_t)b * T) * H + y) * W + xx) * CV + (c0 >> 2) + c4;
        const float4* ap = reinterpret_cast<const float4*>(a_low) +
                           ((((size_t)b * T) * 7 + (y >> rshift)) * 12 + (xx >> rshift)) * CV + (c0 >> 2) + c4;
        const int xs = H * W * CV, as = 84 * CV;
        int t = 0;
        for (; t + 3 <= T; t += 3) {                       // 6 independent loads in flight
            const float4 x0 = xp[t * xs], x1 = xp[(t + 1) * xs], x2 = xp[(t + 2) * xs];
            const float4 a0 = __ldg(ap + t * as), a1 = __ldg(ap + (t + 1) * as), a2 = __ldg(ap + (t + 2) * as);
            acc.x = fmaf(a0.x, x0.x, acc.x); acc.y = fmaf(a0.y, x0.y, acc.y); acc.z = fmaf(a0.z, x0.z, acc.z); acc.w = fmaf(a0.w, x0.w, acc.w);
            acc.x = fmaf(a1.x, x1.x, acc.x); acc.y = fmaf(a1.y, x1.y, acc.y); acc.z = fmaf(a1.z, x1.z, acc.z); acc.w = fmaf(a1.w, x1.w, acc.w);
            acc.x = fmaf(a2.x, x2.x, acc.x); acc.y = fmaf(a2.y, x2.y, acc.y); acc.z = fmaf(a2.z, x2.z, acc.z); acc.w = fmaf(a2.w, x2.w, acc.w);
        }
        for (; t < T; ++t) {
            const float4 x0 = xp[t * xs], a0 = __ldg(ap + t * as);
            acc.x = fmaf(a0.x, x0.x, acc.x); acc.y = fmaf(a0.y, x0.y, acc.y); acc.z = fmaf(a0.z, x0.z, acc.z); acc.w = fmaf(a0.w, x0.w, acc.w);
        }
        float* mp = m + xx * 33 + 4 * c4;
        mp[0] = acc.x * invT; mp[1] = acc.y * invT; mp[2] = acc.z * invT; mp[3] = acc.w * invT;
    }
    __syncthreads();
    {   // softmax over x: 8 threads per channel, shuffle-reduced
        const int c = threadIdx.x >> 3, part = threadIdx.x & 7;
        float mx = -INFINITY;
        for (int xx = part; xx < W; xx += 8) mx = fmaxf(mx, m[xx * 33 + c]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int xx = part; xx < W; xx += 8) { const float e = __expf(m[xx * 33 + c] - mx); m[xx * 33 + c] = e; sum += e; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        for (int xx = part; xx < W; xx += 8) m[xx * 33 + c] *= inv;
    }
    __syncthreads();
    const int WQ = W >> 2;                 // W is 12, 24, 48 or 96
    for (int e = threadIdx.x; e < WQ * 32; e += 256) {
        const int q = e % WQ, c = e / WQ;
        const float* mp = m + (4 * q) * 33 + c;
        reinterpret_cast<float4*>(g + (((size_t)b * C + c0 + c) * H + y) * W)[q] = make_float4(mp[0], mp[33], mp[66], mp[99]);
    }
}

int av_gate_launch(const float* x, const float* a_low, int B, int T, int H, int W, int C, float* g, cudaStream_t s) {
    if (W > 96 || C % 32 || W % 4) return -33;
    const int r = H / 7;
    int rshift = 0;
    while ((1 << rshift) < r) ++rshift;
    if ((1 << rshift) != r || H != 7 * r || W != 12 * r) return -33;
    DSB_PDL_LAUNCH(av_gate_kernel, dim3(H, B, C / 32), 256, 0, s, x, a_low, T, H, W, C, rshift, g);
    DSB_LAUNCH_CHECK();
}

// K source = raw reinterpretation of the contiguous [B][C][T][H][W] buffer (a*g) as [(B T)][H W][C]
// (transformer.py:146) followed by 'b (h w) c -> b c h w' (attention.py:89): element (bt, pix', c') is the flat
// element j = (bt % T)*HW*C + pix'*C + c' of clip b = bt / T, and flat j decodes to (c, t, pix) = [C][T][HW].
// The audio map is read from its channel-major copy a_cm[b][c][t][7][12] (audio_cmajor_launch, once per conditioning):
// for a fixed source channel the 1 / 2 / 4 audio values under 4 consecutive source pixels are then adjacent floats and the
// lanes of a warp walk contiguous memory.  (Read from the token-major a_low[b][t][84][c] the same values were 4-byte
// gathers 4*C bytes apart: one 32-byte sector per value, 8 % of the HBM rate in ncu.)
template <int C, int H, int W, int S_, int T_>
__global__ void kpool_av_kernel(const float* __restrict__ g, const float* __restrict__ a_cm, int tmax,
                                const float* __restrict__ wk, const float* __restrict__ kg,
                                const float* __restrict__ kb, bf16* __restrict__ out, const ScoreBias sbv) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    __shared__ float red[32];
    constexpr int HW = H * W, R = H / 7, CV = C / 4, THW = T_ * HW, NPX = S_ * S_;
    static_assert(HW % 4 == 0 && W % 4 == 0 && C % 4 == 0, "4 consecutive flat indices stay inside one source row");
    const int tokv = blockIdx.x;                 // bt*18 + Y*6 + X
    const int bt = tokv / 18, Y = (tokv % 18) / 6, X = tokv % 6;
    const int b = bt / T_, tt = bt % T_;
    if (tt >= tmax) return;
    const int clip_base = tt * HW * C;           // < 9 * 516096, fits int
    const int tid = threadIdx.x;
    // Thread (4-channel vector c4, pixel group gq).  Token (tt, pixel p') channel c of the reference's raw .view of the
    // gated [C, T, H, W] clip is flat element j = (tt*HW + p')*C + c; 4 consecutive channels are 4 consecutive source
    // pixels of one row (all extents are multiples of 4), so gate values come as one float4 and the audio map (nearest
    // upsample by R) as 1, 2 or 4 scalars.  All divisors are compile-time.
    const int G = blockDim.x / CV;
    const int c4 = tid % CV, gq = tid / CV;
    const float* a_b = a_cm + (size_t)b * C * T_ * 84;
    const float* g_b = g + (size_t)b * C * HW;
    if (gq < G) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int p = gq; p < NPX; p += G) {
            const int dy = p / S_, dx = p % S_;
            const int j = clip_base + ((Y * S_ + dy) * W + X * S_ + dx) * C + 4 * c4;
            const int cs = j / THW, rem = j % THW;
            const int ts = rem / HW, pix = rem % HW;
            const int ys = pix / W, xs = pix % W;
            const float4 gv = *reinterpret_cast<const float4*>(g_b + (size_t)cs * HW + pix);
            const float4 w = __ldg(reinterpret_cast<const float4*>(wk) + p * CV + c4);
            const float* ar = a_b + (((size_t)cs * T_ + ts) * 7 + (ys / R)) * 12;
            float a0, a1, a2, a3;
            if constexpr (R >= 4) {
                a0 = a1 = a2 = a3 = __ldg(ar + xs / R);
            } else if constexpr (R == 2) {
                const float2 t2 = __ldg(reinterpret_cast<const float2*>(ar + xs / 2));      // xs % 4 == 0
                a0 = a1 = t2.x;
                a2 = a3 = t2.y;
            } else {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(ar + xs));
                a0 = t4.x; a1 = t4.y; a2 = t4.z; a3 = t4.w;
            }
            acc.x = fmaf(w.x, a0 * gv.x, acc.x);
            acc.y = fmaf(w.y, a1 * gv.y, acc.y);
            acc.z = fmaf(w.z, a2 * gv.z, acc.z);
            acc.w = fmaf(w.w, a3 * gv.w, acc.w);
        }
        reinterpret_cast<float4*>(sm)[gq * CV + c4] = acc;
    }
    pool_fold_groups(sm, C, G);
    pooled_ln_store(sm, C, kg, kb, out + (size_t)tokv * C, red, &sbv, tokv);
}

// a_low[(b*T + t)*84 + p][c]  ->  a_cm[b][c][t*84 + p]     (32 x 32 shared-memory transpose, coalesced on both sides)
__global__ void __launch_bounds__(256) audio_cmajor_kernel(const float* __restrict__ a_low, int C, int TP,
                                                          float* __restrict__ a_cm) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int p = p0 + k, c = c0 + tx;
        tile[k][tx] = (p < TP && c < C) ? a_low[((size_t)b * TP + p) * C + c] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, p = p0 + tx;
        if (p < TP && c < C) a_cm[((size_t)b * C + c) * TP + p] = tile[tx][k];
    }
}

int audio_cmajor_launch(const float* a_low, int B, int T, int C, float* a_cm, cudaStream_t s) {
    const int TP = T * 84;
    audio_cmajor_kernel<<<dim3((TP + 31) / 32, (C + 31) / 32, B), 256, 0, s>>>(a_low, C, TP, a_cm);
    DSB_LAUNCH_CHECK();
}

int kpool_av_launch(const float* g, const float* a_low, int B, int T, int H, int W, int C, int s_, const float* wk,
                    const float* kg, const float* kb, bf16* out, int tmax, cudaStream_t s, ScoreBias sbv) {
    const int nthr = pool_threads(C);
    if (T != 9 || C / 4 > nthr) return -34;
    const size_t smem = (size_t)(nthr / (C / 4)) * C * sizeof(float);
    const int grid = B * T * 18;
    if (C == 768 && H == 7 && W == 12 && s_ == 2) DSB_PDL_LAUNCH((kpool_av_kernel<768, 7, 12, 2, 9>), grid, nthr, smem, s, g, a_low, tmax, wk, kg, kb, out, sbv);
    else if (C == 384 && H == 14 && W == 24 && s_ == 4) DSB_PDL_LAUNCH((kpool_av_kernel<384, 14, 24, 4, 9>), grid, nthr, smem, s, g, a_low, tmax, wk, kg, kb, out, sbv);
    else if (C == 192 && H == 28 && W == 48 && s_ == 8) DSB_PDL_LAUNCH((kpool_av_kernel<192, 28, 48, 8, 9>), grid, nthr, smem, s, g, a_low, tmax, wk, kg, kb, out, sbv);
    else if (C == 96 && H == 56 && W == 96 && s_ == 16) DSB_PDL_LAUNCH((kpool_av_kernel<96, 56, 96, 16, 9>), grid, nthr, smem, s, g, a_low, tmax, wk, kg, kb, out, sbv);
    else return -34;
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ projections folded into K / V
// The 18 keys / values of a frame only ever meet the queries through  q_ln Wq^T K^T  and  P V Wp^T  (attention.py:97-113), so
// the query projection folds into the keys and the output projection into the values, and -- because K = k_ln Wk^T + bk and
// V = v_ln Wv^T + bv are themselves linear -- both fold into the WEIGHTS once per weight set:
//   K1[h*18+j][c] = scale sum_{c' in head h} K[j,c'] Wq[c',c] = k_ln[j,:] . MK_h[c,:] + cK_h[c]
//   sb[h*18+j]    = scale sum_{c' in head h} K[j,c'] bq[c']   = k_ln[j,:] . mb_h[:]   + cb_h
//   V2[c][h*18+j] = sum_{c' in head h} Wp[c,c'] V[j,c']       = v_ln[j,:] . MV_h[c,:] + cV_h[c]
// with MK_h = scale Wq_h^T Wk_h, MV_h = Wp[:,h] Wv_h (C x C per head).  One GEMM per branch (N = 2C) then yields the
// per-frame operands of the score and P.V products directly; proj_q, proj and the per-evaluation fold kernel disappear.
// grid (C/16, C/16, 2 heads), block (16, 16): out[h][c][e]
__global__ void __launch_bounds__(256) fold_weights_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                                          const float* __restrict__ wp, const float* __restrict__ wv, int C,
                                                          float scale, float* __restrict__ MK, float* __restrict__ MV) {
    const int e = blockIdx.x * 16 + threadIdx.x, c = blockIdx.y * 16 + threadIdx.y, h = blockIdx.z;
    const int d = C >> 1, d0 = h * d;
    float ak = 0.0f, av = 0.0f;
    for (int i = 0; i < d; ++i) {
        const int cp = d0 + i;
        ak = fmaf(wq[(size_t)cp * C + c], wk[(size_t)cp * C + e], ak);
        av = fmaf(wp[(size_t)c * C + cp], wv[(size_t)cp * C + e], av);
    }
    MK[((size_t)h * C + c) * C + e] = ak * scale;
    MV[((size_t)h * C + c) * C + e] = av;
}

// cK[h*C+c] = scale sum Wq[c',c] bk[c'];  cV[h*C+c] = sum Wp[c,c'] bv[c'];  mb[h*C+e] = scale sum bq[c'] Wk[c',e];
// cb[h] = scale sum bq[c'] bk[c']          (sums over c' in head h)
__global__ void __launch_bounds__(256) fold_bias_kernel(const float* __restrict__ wq, const float* __restrict__ bq,
                                                       const float* __restrict__ wk, const float* __restrict__ bk,
                                                       const float* __restrict__ wp, const float* __restrict__ bv, int C, float scale,
                                                       float* __restrict__ cK, float* __restrict__ cV, float* __restrict__ mb,
                                                       float* __restrict__ cb) {
    const int c = blockIdx.x * 256 + threadIdx.x, h = blockIdx.y;
    const int d = C >> 1, d0 = h * d;
    if (c < C) {
        float a = 0.0f, v = 0.0f, m = 0.0f;
        for (int i = 0; i < d; ++i) {
            const int cp = d0 + i;
            a = fmaf(wq[(size_t)cp * C + c], bk[cp], a);
            v = fmaf(wp[(size_t)c * C + cp], bv[cp], v);
            m = fmaf(bq[cp], wk[(size_t)cp * C + c], m);
        }
        cK[h * C + c] = a * scale;
        cV[h * C + c] = v;
        mb[h * C + c] = m * scale;
    }
    if (c == 0) {
        float t = 0.0f;
        for (int i = 0; i < d; ++i) t = fmaf(bq[d0 + i], bk[d0 + i], t);
        cb[h] = t * scale;
    }
}

int fold_weights_launch(const float* wq, const float* bq, const float* wk, const float* bk, const float* wp, const float* wv,
                        const float* bv, int C, float scale, float* MK, float* MV, float* cK, float* cV, float* mb, float* cb,
                        cudaStream_t s) {
    if (C % 16) return -38;
    fold_weights_kernel<<<dim3(C / 16, C / 16, 2), dim3(16, 16), 0, s>>>(wq, wk, wp, wv, C, scale, MK, MV);
    fold_bias_kernel<<<dim3((C + 255) / 256, 2), 256, 0, s>>>(wq, bq, wk, bk, wp, bv, C, scale, cK, cV, mb, cb);
    DSB_LAUNCH_CHECK();
}

// Per frame: the folded GEMM outputs Kf / Vf [F*18][2C] fp32 (column h*C + c) -> the operand layouts of the score and P.V
// products:  K1 bf16 [F][R][C] (row h*18+j; rows 36..R-1 zero), sb fp32 [F][R], V2 bf16 [F][C][64] (column h*18+j; 36..63 zero)
__global__ void __launch_bounds__(256) kv_pack_kernel(const float* __restrict__ Kf, const float* __restrict__ Vf,
                                                     const bf16* __restrict__ k_ln, const float* __restrict__ mb,
                                                     const float* __restrict__ cb, int C, int R, int T, int tmax,
                                                     bf16* __restrict__ K1, float* __restrict__ sb, bf16* __restrict__ V2) {
    pdl_trigger();
    pdl_wait();
    // grid (frame, part): part p owns the 8-channel groups p, p + P, ... of K1, the channels p*256/P.. of V2 and the
    // key rows r = p (mod P) of the score bias
    const int f = blockIdx.x, part = blockIdx.y, P = gridDim.y, tid = threadIdx.x;
    if (f % T >= tmax) return;
    const int c8n = C >> 3, ld = 2 * C;
    for (int it = part * 256 + tid; it < R * c8n; it += 256 * P) {
        const int r = it / c8n, c = (it - r * c8n) * 8;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (r < 36) {
            const float4* src = reinterpret_cast<const float4*>(Kf + ((size_t)f * 18 + r % 18) * ld + (r / 18) * C + c);
            const float4 a = src[0], b = src[1];
            o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
        }
        reinterpret_cast<uint4*>(K1 + ((size_t)f * R + r) * C + c)[0] = o;
    }
    {   // folded score bias: one warp per key row
        const int lane = tid & 31;
        for (int r = part * 8 + (tid >> 5); r < R; r += 8 * P) {
            float a = 0.0f;
            if (r < 36) {
                const bf16* kr = k_ln + ((size_t)f * 18 + r % 18) * C;
                const float* m = mb + (r / 18) * C;
                for (int c = lane; c < C; c += 32) a = fmaf(__bfloat162float(kr[c]), __ldg(m + c), a);
                a = warp_sum(a) + __ldg(cb + r / 18);
            }
            if (lane == 0) sb[(size_t)f * R + r] = a;
        }
    }
    for (int c = part * 256 + tid; c < C; c += 256 * P) {
        uint32_t w[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k] = 0u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float v0 = Vf[((size_t)f * 18 + 2 * k) * ld + hh * C + c], v1 = Vf[((size_t)f * 18 + 2 * k + 1) * ld + hh * C + c];
                w[hh * 9 + k] = pack_bf16x2(v0, v1);
            }
        uint4* dst = reinterpret_cast<uint4*>(V2 + ((size_t)f * C + c) * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}

int kv_pack_launch(const float* Kf, const float* Vf, const bf16* k_ln, const float* mb, const float* cb, int F, int C, int R,
                   int T, int tmax, bf16* K1, float* sb, bf16* V2, cudaStream_t s) {
    if (C % 8 || R < 36) return -38;
    const int P = C >= 768 ? 6 : C >= 384 ? 4 : 2;           // >= 2 blocks per SM at 72 frames
    DSB_PDL_LAUNCH(kv_pack_kernel, dim3(F, P), 256, 0, s, Kf, Vf, k_ln, mb, cb, C, R, T, tmax, K1, sb, V2);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ attention operands
__global__ void __launch_bounds__(256) attn_operands_kernel(const float* __restrict__ kp, const float* __restrict__ vp,
                                                           int C, float scale, bf16* __restrict__ KB,
                                                           bf16* __restrict__ VB) {
    pdl_trigger();
    pdl_wait();
    // blockIdx.y = frame.  Items 0 .. 48*C/8-1: one uint4 (8 channels) of KB[f][row][c]; then one 128-byte row
    // VB[f][c][0..63] per item.  A head owns a contiguous half of the channels (d = C/2, a multiple of 8).
    const int f = blockIdx.y, d = C >> 1, c8n = C >> 3, nK = 48 * c8n;
    const int item = blockIdx.x * 256 + threadIdx.x;
    if (item < nK) {
        const int row = item / c8n, c = (item - row * c8n) * 8;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (row < 36 && (c / d) == (row / 18)) {
            const float4* src = reinterpret_cast<const float4*>(kp + ((size_t)f * 18 + row % 18) * C + c);
            const float4 a = src[0], b = src[1];
            o = make_uint4(pack_bf16x2(a.x * scale, a.y * scale), pack_bf16x2(a.z * scale, a.w * scale),
                           pack_bf16x2(b.x * scale, b.y * scale), pack_bf16x2(b.z * scale, b.w * scale));
        }
        reinterpret_cast<uint4*>(KB + ((size_t)f * 48 + row) * C + c)[0] = o;
    } else if (item < nK + C) {
        const int c = item - nK, h = c / d;
        float v[18];
#pragma unroll
        for (int j = 0; j < 18; ++j) v[j] = vp[((size_t)f * 18 + j) * C + c];
        uint32_t w[32];                                       // 64 bf16 columns: head h owns columns 18h .. 18h+17
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k] = 0u;
        if (h == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) w[9 + k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
        }
        uint4* dst = reinterpret_cast<uint4*>(VB + ((size_t)f * C + c) * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}

int attn_operands_launch(const float* kp, const float* vp, int F, int C, float scale, bf16* KB, bf16* VB,
                         cudaStream_t s) {
    if (C % 16) return -36;
    const int items = 48 * (C / 8) + C;
    DSB_PDL_LAUNCH(attn_operands_kernel, dim3((items + 255) / 256, F), 256, 0, s, kp, vp, C, scale, KB, VB);
    DSB_LAUNCH_CHECK();
}

// Folded attention operands of the narrow stages (consumed by the fused attention chain, mlp_fused.cu mode 1):
//   K1[f][hj][c]   = scale * sum_{d in head h} K[f,j,d] * Wq[d][c]      (bf16 [F][64][C], rows 36..63 zero)
//   sb[f][hj]      = scale * sum_{d in head h} K[f,j,d] * bq[d]         (fp32 [F][64])
//   V2[f][c][hj]   = sum_{d in head h} Wp[c][d] * V[f,j,d]              (bf16 [F][C][64], columns 36..63 zero)
// one block per (head, group of 3 keys, frame), one thread per output channel c: a weight element read from L2 feeds
// 3 keys / 3 values (the first version, one block per key, moved 36 x both weight matrices per frame through L2)
__global__ void __launch_bounds__(384) attn_fold_kernel(const float* __restrict__ kp, const float* __restrict__ vp,
                                                       const float* __restrict__ wq, const float* __restrict__ bq,
                                                       const float* __restrict__ wpT, int C, float scale, int T, int tmax,
                                                       bf16* __restrict__ K1, float* __restrict__ sb, bf16* __restrict__ V2) {
    pdl_trigger();
    pdl_wait();
    constexpr int KPB = 3;
    __shared__ float sk[KPB][192];
    __shared__ float sv[KPB][192];
    const int h = blockIdx.x / 6, j0 = (blockIdx.x % 6) * KPB, f = blockIdx.y, c = threadIdx.x;
    if (f % T >= tmax) return;
    const int d = C >> 1, d0 = h * d;
    for (int e = c; e < KPB * d; e += blockDim.x) {
        const int j = e / d, i = e - j * d;
        sk[j][i] = kp[((size_t)f * 18 + j0 + j) * C + d0 + i];
        sv[j][i] = vp[((size_t)f * 18 + j0 + j) * C + d0 + i];
    }
    __syncthreads();
    float ak[KPB], av[KPB];
#pragma unroll
    for (int j = 0; j < KPB; ++j) { ak[j] = 0.0f; av[j] = 0.0f; }
#pragma unroll 4
    for (int i = 0; i < d; ++i) {
        const float w = wq[(size_t)(d0 + i) * C + c], wp = wpT[(size_t)(d0 + i) * C + c];
#pragma unroll
        for (int j = 0; j < KPB; ++j) { ak[j] = fmaf(sk[j][i], w, ak[j]); av[j] = fmaf(sv[j][i], wp, av[j]); }
    }
    bf16* v2row = V2 + ((size_t)f * C + c) * 64;
#pragma unroll
    for (int j = 0; j < KPB; ++j) {
        K1[((size_t)f * 64 + h * 18 + j0 + j) * C + c] = __float2bfloat16(ak[j] * scale);
        v2row[h * 18 + j0 + j] = __float2bfloat16(av[j]);
    }
    if (blockIdx.x == 0) {                                   // zero padding: key rows / value columns 36..63
        for (int r = 36; r < 64; ++r) K1[((size_t)f * 64 + r) * C + c] = __float2bfloat16(0.0f);
#pragma unroll
        for (int k = 0; k < 7; ++k) reinterpret_cast<uint2*>(v2row + 36)[k] = make_uint2(0u, 0u);
        if (c < 28) sb[(size_t)f * 64 + 36 + c] = 0.0f;
    }
    if (c < KPB) {
        float b = 0.0f;
        for (int i = 0; i < d; ++i) b = fmaf(sk[c][i], bq[d0 + i], b);
        sb[(size_t)f * 64 + h * 18 + j0 + c] = b * scale;
    }
}

int attn_fold_launch(const float* kp, const float* vp, const float* wq, const float* bq, const float* wpT, int F, int C,
                     float scale, int T, int tmax, bf16* K1, float* sb, bf16* V2, cudaStream_t s) {
    if (C > 384 || C < 32) return -36;
    DSB_PDL_LAUNCH(attn_fold_kernel, dim3(12, F), C, 0, s, kp, vp, wq, bq, wpT, C, scale, T, tmax, K1, sb, V2);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ multi-scale sum
struct MsSrc { const float* r[4]; };

// thread = (4-channel vector, output row); produces XT = 16 consecutive output columns.  All four scales are exact powers
// of two (1/2 .. 1/16), so for an output column xo0 + dx (xo0 a multiple of 16) the left source column is
// (xo0 >> n) + off(dx, n) and the horizontal weight lx(dx, n) are COMPILE-TIME constants of dx.  The sources are taken one
// after the other: the 16 outputs touch only NC = 16 / 2^n + 2 source columns, which are loaded (both rows) and
// interpolated vertically ONCE into registers, then every output adds (1 - lx) * V[j] + lx * V[j + 1] with static j and
// immediates for the weights.  Index clamping at the image edges reproduces PyTorch's "clamp the source coordinate at
// 0 / last" exactly (both taps then hit the same pixel).  ~50 SASS instructions per output vector in 14 KB of code (the
// first version walked all four sources column by column: 88 instructions per vector in 45 KB, beyond the 32 KB
// instruction cache).
constexpr int kMsXT = 16;

template <int N, int DX>
__device__ __forceinline__ void ms_out(const float4 (&V)[(kMsXT >> N) + 2], float4 (&acc)[kMsXT]) {
    constexpr int j = PowX<N>::off(DX) + 1;                       // V[0] is source column (xo0 >> N) - 1
    constexpr float lx = PowX<N>::lx(DX), w0 = 1.0f - lx;
    acc[DX].x = fmaf(lx, V[j + 1].x, fmaf(w0, V[j].x, acc[DX].x));
    acc[DX].y = fmaf(lx, V[j + 1].y, fmaf(w0, V[j].y, acc[DX].y));
    acc[DX].z = fmaf(lx, V[j + 1].z, fmaf(w0, V[j].z, acc[DX].z));
    acc[DX].w = fmaf(lx, V[j + 1].w, fmaf(w0, V[j].w, acc[DX].w));
    if constexpr (DX + 1 < kMsXT) ms_out<N, DX + 1>(V, acc);
}

template <int N>
__device__ __forceinline__ void ms_source(const float* __restrict__ r, int b, int cv, int yo, int xo0, float4 (&acc)[kMsXT]) {
    constexpr int H = 112 >> N, W = 192 >> N, CV = 192, NC = (kMsXT >> N) + 2;
    int y0, y1;
    float ly;
    bil_src(yo, 1.0f / (float)(1 << N), H, y0, y1, ly);
    const float4* row0 = reinterpret_cast<const float4*>(r) + ((size_t)b * H + y0) * W * CV + cv;
    const float4* row1 = reinterpret_cast<const float4*>(r) + ((size_t)b * H + y1) * W * CV + cv;
    const int xb = (xo0 >> N) - 1;
    float4 V[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int x = min(max(xb + j, 0), W - 1);
        const float4 a = row0[x * CV], c = row1[x * CV];
        V[j] = make_float4(fmaf(ly, c.x - a.x, a.x), fmaf(ly, c.y - a.y, a.y), fmaf(ly, c.z - a.z, a.z), fmaf(ly, c.w - a.w, a.w));
    }
    ms_out<N, 0>(V, acc);
}

__global__ void __launch_bounds__(256, 2) ms_sum_kernel(MsSrc src, bf16* __restrict__ S) {
    pdl_trigger();
    pdl_wait();
    constexpr int CV = 192, OH = 112, OW = 192;
    const int cv = blockIdx.z % 12 * 16 + (threadIdx.x & 15);
    const int b = blockIdx.z / 12;
    const int yo = blockIdx.y * 16 + (threadIdx.x >> 4);
    const int xo0 = blockIdx.x * kMsXT;
    float4 acc[kMsXT];
#pragma unroll
    for (int i = 0; i < kMsXT; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ms_source<4>(src.r[0], b, cv, yo, xo0, acc);          // 7x12    (1/16)
    ms_source<3>(src.r[1], b, cv, yo, xo0, acc);          // 14x24   (1/8)
    ms_source<2>(src.r[2], b, cv, yo, xo0, acc);          // 28x48   (1/4)
    ms_source<1>(src.r[3], b, cv, yo, xo0, acc);          // 56x96   (1/2)
    uint2* orow = reinterpret_cast<uint2*>(S) + (((size_t)b * OH + yo) * OW + xo0) * CV + cv;
#pragma unroll
    for (int i = 0; i < kMsXT; ++i)
        orow[(size_t)i * CV] = make_uint2(pack_f16x2(acc[i].x, acc[i].y), pack_f16x2(acc[i].z, acc[i].w));   // fp16: mt_proj operand
}

int ms_sum_launch(const float* const r[4], int B, bf16* S, cudaStream_t s) {
    MsSrc src;
    for (int k = 0; k < 4; ++k) src.r[k] = r[k];
    DSB_PDL_LAUNCH(ms_sum_kernel, dim3(192 / kMsXT, 112 / 16, B * 12), 256, 0, s, src, S);
    DSB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) final_up_kernel(const float* __restrict__ p, int B, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const long total = (long)B * 224 * 384;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int xo = (int)(i % 384);
        const int yo = (int)((i / 384) % 224);
        const int b = (int)(i / (384 * 224));
        int y0, y1, x0, x1; float ly, lx;
        bil_src(yo, 0.5f, 112, y0, y1, ly);
        bil_src(xo, 0.5f, 192, x0, x1, lx);
        const float* s = p + (size_t)b * 112 * 192;
        const float v00 = s[y0 * 192 + x0], v01 = s[y0 * 192 + x1], v10 = s[y1 * 192 + x0], v11 = s[y1 * 192 + x1];
        out[i] = (1.0f - ly) * ((1.0f - lx) * v00 + lx * v01) + ly * ((1.0f - lx) * v10 + lx * v11);
    }
}

int final_up_launch(const float* p, int B, float* out, cudaStream_t s) {
    const long total = (long)B * 224 * 384;
    long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    DSB_PDL_LAUNCH(final_up_kernel, (int)g, 256, 0, s, p, B, out);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ sampler update
struct AxpyArgs { const float* in[4]; float c[4]; };

__global__ void __launch_bounds__(256) axpy_kernel(int nin, AxpyArgs a, const float* __restrict__ noise, float cn,
                                                  float* __restrict__ out, long n4) {
    pdl_trigger();
    pdl_wait();
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
        for (int k = 0; k < nin; ++k) {
            const float4 v = reinterpret_cast<const float4*>(a.in[k])[i];
            const float c = a.c[k];
            acc.x = fmaf(c, v.x, acc.x); acc.y = fmaf(c, v.y, acc.y); acc.z = fmaf(c, v.z, acc.z); acc.w = fmaf(c, v.w, acc.w);
        }
        if (noise) {
            const float4 z = reinterpret_cast<const float4*>(noise)[i];
            acc.x = fmaf(cn, z.x, acc.x); acc.y = fmaf(cn, z.y, acc.y); acc.z = fmaf(cn, z.z, acc.z); acc.w = fmaf(cn, z.w, acc.w);
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
}

int axpy_launch(int nin, const float* const in[4], const float c[4], const float* noise, float cn, float* out, long n,
                cudaStream_t s) {
    if (nin < 1 || nin > 4 || (n & 3)) return -32;
    AxpyArgs a;
    for (int k = 0; k < 4; ++k) { a.in[k] = k < nin ? in[k] : nullptr; a.c[k] = k < nin ? c[k] : 0.0f; }
    const long n4 = n >> 2;
    long g = (n4 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    DSB_PDL_LAUNCH(axpy_kernel, (int)g, 256, 0, s, nin, a, noise, cn, out, n4);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ post-processing
// inverse_data_transform (clamp to [0,1], datasets/__init__.py:35) followed by normalize_data (per-map min-max to
// uint8, util/utils.py:11-16): one block per clip, two passes over its 86 016 pixels.
__global__ void __launch_bounds__(1024) postprocess_kernel(const float* __restrict__ x, int n, float* __restrict__ clamped,
                                                          uint8_t* __restrict__ u8) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    __shared__ float smin, smax;
    const float* xb = x + (size_t)blockIdx.x * n;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = fminf(fmaxf(xb[i], 0.0f), 1.0f);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    mn = -warp_max(-mn);
    if (lane == 0) red[wid] = mn;
    __syncthreads();
    if (wid == 0) { float t = lane < (blockDim.x >> 5) ? red[lane] : INFINITY; t = -warp_max(-t); if (lane == 0) smin = t; }
    __syncthreads();
    mx = warp_max(mx);
    if (lane == 0) red[wid] = mx;
    __syncthreads();
    if (wid == 0) { float t = lane < (blockDim.x >> 5) ? red[lane] : -INFINITY; t = warp_max(t); if (lane == 0) smax = t; }
    __syncthreads();
    const float lo = smin;
    const float scale = 255.0f / (smax - lo);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = fminf(fmaxf(xb[i], 0.0f), 1.0f);
        if (clamped) clamped[(size_t)blockIdx.x * n + i] = v;
        if (u8) u8[(size_t)blockIdx.x * n + i] = (uint8_t)fminf(fmaxf((v - lo) * scale, 0.0f), 255.0f);
    }
}

int postprocess_launch(const float* x, int B, int n, float* clamped, uint8_t* u8, cudaStream_t s) {
    DSB_PDL_LAUNCH(postprocess_kernel, B, 1024, 0, s, x, n, clamped, u8);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ adaptive-step error norm
// dpm_solver_adaptive (sampler.py:996-999): per sample  E_b = sqrt(mean(((x_higher - x_lower) / delta)^2)),
// delta = max(atol, rtol * max(|x_lower|, |x_prev|)).  One block per sample, fp32 like the reference.
__global__ void __launch_bounds__(1024) adaptive_error_kernel(const float* __restrict__ xl, const float* __restrict__ xh,
                                                             const float* __restrict__ xp, int n, float atol, float rtol,
                                                             float* __restrict__ out) {
    __shared__ float red[32];
    const size_t base = (size_t)blockIdx.x * n;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = xl[base + i], b = xh[base + i], c = xp[base + i];
        const float delta = fmaxf(atol, rtol * fmaxf(fabsf(a), fabsf(c)));
        const float e = (b - a) / delta;
        acc = fmaf(e, e, acc);
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = sqrtf(tot / (float)n);
}

int adaptive_error_launch(const float* xl, const float* xh, const float* xp, int B, int n, float atol, float rtol, float* out,
                          cudaStream_t s) {
    adaptive_error_kernel<<<B, 1024, 0, s>>>(xl, xh, xp, n, atol, rtol, out);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ content fingerprint
// 64-bit order-independent fingerprint of a device buffer (sum over 16-byte words of a position-keyed splitmix64 of their
// two halves).  Lets the host recognise that the conditioning tensors of this denoiser call hold the same values as the
// last call's (the reference's sample_ddim deep-copies the feature list every step, diffusion_trainer.py:452, so the
// pointers change while the content does not) without re-running the conditioning.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) content_hash_kernel(const uint4* __restrict__ p, size_t n16, unsigned long long seed,
                                                          unsigned long long* __restrict__ out) {
    unsigned long long h = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = p[i];
        const unsigned long long a = ((unsigned long long)v.x << 32) | v.y, b = ((unsigned long long)v.z << 32) | v.w;
        const unsigned long long k = seed + (unsigned long long)i * 0x9E3779B97F4A7C15ull;
        h += mix64(a + k) + mix64(b ^ (k * 0xD6E8FEB86659FD93ull + 1ull));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);          // integer sum: order-independent, deterministic
}

int content_hash_launch(const void* p, size_t bytes, unsigned long long seed, unsigned long long* out, cudaStream_t s) {
    if (bytes % 16 || (reinterpret_cast<uintptr_t>(p) & 15)) return -40;
    const size_t n16 = bytes / 16;
    int grid = (int)((n16 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) return 0;
    content_hash_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(p), n16, seed, out);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ layout conversion
// [B][C][Tv][HW] -> [(b*T+t)][HW][C] through a 32x32 smem transpose (coalesced on both sides)
__global__ void __launch_bounds__(256) nct_to_frames_kernel(const float* __restrict__ vis, int C, int Tv, int HW, int T,
                                                           float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int bt = blockIdx.z;                // b*Tv + t
    const int b = bt / Tv, t = bt % Tv;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, p = p0 + tx;
        tile[k][tx] = (c < C && p < HW) ? vis[(((size_t)b * C + c) * Tv + t) * HW + p] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int p = p0 + k, c = c0 + tx;
        if (c < C && p < HW) dst[(((size_t)b * T + t) * HW + p) * C + c] = tile[tx][k];
    }
}

int nct_to_frames_launch(const float* vis, int B, int C, int Tv, int HW, int T, float* dst, cudaStream_t s) {
    nct_to_frames_kernel<<<dim3((HW + 31) / 32, (C + 31) / 32, B * Tv), 256, 0, s>>>(vis, C, Tv, HW, T, dst);
    DSB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) audio_tokens_kernel(const float* __restrict__ audio, int T,
                                                          bf16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int bt = blockIdx.z;
    const int b = bt / T, t = bt % T;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, p = p0 + tx;
        tile[k][tx] = (p < 84) ? audio[(((size_t)b * 512 + c) * T + t) * 84 + p] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int p = p0 + k, c = c0 + tx;
        if (p < 84) out[(((size_t)b * T + t) * 84 + p) * 512 + c] = __float2bfloat16(tile[tx][k]);
    }
}

int audio_tokens_launch(const float* audio, int B, int T, bf16* out, cudaStream_t s) {
    audio_tokens_kernel<<<dim3(3, 16, B * T), 256, 0, s>>>(audio, T, out);
    DSB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ x, long n, bf16* __restrict__ out) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256)
        out[i] = __float2bfloat16(x[i]);
}

int to_bf16_launch(const float* x, long n, bf16* out, cudaStream_t s) {
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    to_bf16_kernel<<<(int)g, 256, 0, s>>>(x, n, out);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ audio transformer glue
// (models/audio_attention.py:55-69) qkv [tokens][3 * 128] -> per-head operands of the two attention GEMMs:
//   Qh [2][B][n][64] (pre-scaled by 64^-0.5 = 0.125, exact in bf16), Kh [2][B][npad][64], Vt [2][B][64][npad]
// rows / columns n..npad-1 are never written (zeroed once at allocation).
__global__ void __launch_bounds__(256) qkv_split_kernel(const bf16* __restrict__ qkv, int B, int n, int npad,
                                                       bf16* __restrict__ Qh, bf16* __restrict__ Kh,
                                                       bf16* __restrict__ Vt) {
    __shared__ bf16 vt[32][130];
    const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x;
    const __nv_bfloat162 eighth = __floats2bfloat162_rn(0.125f, 0.125f);
    // q, k: 32 tokens x 2 x 16 uint4 ; v: 32 tokens x 16 uint4 staged for the transpose
    for (int e = tid; e < 32 * 48; e += 256) {
        const int tk = e / 48, u = e % 48, tok = t0 + tk;
        if (tok >= n) continue;
        uint4 v = reinterpret_cast<const uint4*>(qkv + ((size_t)b * n + tok) * 384)[u];
        const int part = u >> 4, hh = (u >> 3) & 1, d8 = u & 7;     // part: 0 q, 1 k, 2 v ; 8 uint4 per head
        if (part == 0) {
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int i = 0; i < 4; ++i) h2[i] = __hmul2(h2[i], eighth);
            reinterpret_cast<uint4*>(Qh + (((size_t)hh * B + b) * n + tok) * 64)[d8] = v;
        } else if (part == 1) {
            reinterpret_cast<uint4*>(Kh + (((size_t)hh * B + b) * npad + tok) * 64)[d8] = v;
        } else {
            const bf16* e8 = reinterpret_cast<const bf16*>(&v);
#pragma unroll
            for (int i = 0; i < 8; ++i) vt[tk][hh * 64 + d8 * 8 + i] = e8[i];
        }
    }
    __syncthreads();
    for (int e = tid; e < 128 * 32; e += 256) {
        const int tk = e & 31, r = e >> 5, tok = t0 + tk;          // r = head * 64 + d
        if (tok < n) Vt[(((size_t)(r >> 6) * B + b) * 64 + (r & 63)) * npad + tok] = vt[tk][r];
    }
}

int qkv_split_launch(const bf16* qkv, int B, int n, int npad, bf16* Qh, bf16* Kh, bf16* Vt, cudaStream_t s) {
    qkv_split_kernel<<<dim3((n + 31) / 32, B), 256, 0, s>>>(qkv, B, n, npad, Qh, Kh, Vt);
    DSB_LAUNCH_CHECK();
}

// softmax over the first `valid` of `ld` (= 768) columns of each fp32 row -> bf16 probabilities (padding columns 0)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long rows, int valid,
                                                          bf16* __restrict__ P) {
    constexpr int LD = 768, NV = LD / 128;               // 6 float4 per lane
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(S + row * LD);
    float4 v[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = src[lane + 32 * i];
        const int c = 4 * (lane + 32 * i);
        if (c + 0 >= valid) v[i].x = -INFINITY;
        if (c + 1 >= valid) v[i].y = -INFINITY;
        if (c + 2 >= valid) v[i].z = -INFINITY;
        if (c + 3 >= valid) v[i].w = -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx);
        v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float inv = 1.0f / warp_sum(sum);
    uint2* dst = reinterpret_cast<uint2*>(P + row * LD);
#pragma unroll
    for (int i = 0; i < NV; ++i)
        dst[lane + 32 * i] = make_uint2(pack_bf16x2(v[i].x * inv, v[i].y * inv), pack_bf16x2(v[i].z * inv, v[i].w * inv));
}

int softmax_rows_launch(const float* S, long rows, int valid, int ld, bf16* P, cudaStream_t s) {
    if (ld != 768 || valid < 1 || valid > ld) return -37;
    softmax_rows_kernel<<<(int)((rows + 7) / 8), 256, 0, s>>>(S, rows, valid, P);
    DSB_LAUNCH_CHECK();
}

// final LayerNorm of the audio transformer, written back channels-first: x [B][n][512] -> out [B][512][n]
// (audio_attention.py:90,141: 'b (t h w) c -> b c t h w').  16 tokens per block through a padded smem tile.
__global__ void __launch_bounds__(256) ln_nct_kernel(const float* __restrict__ x, int n, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ out) {
    constexpr int C = 512;
    __shared__ float tile[16][C + 1];
    const int b = blockIdx.y, t0 = blockIdx.x * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tk = warp; tk < 16; tk += 8) {
        const int tok = t0 + tk;
        if (tok >= n) continue;
        const float4* row = reinterpret_cast<const float4*>(x + ((size_t)b * n + tok) * C);
        float4 v[4];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = row[lane + 32 * i]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q = fmaf(a, a, q); q = fmaf(bb, bb, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c4 = lane + 32 * i;
            const float4 g = reinterpret_cast<const float4*>(gamma)[c4], be = reinterpret_cast<const float4*>(beta)[c4];
            float* d = &tile[tk][4 * c4];
            d[0] = (v[i].x - mean) * rstd * g.x + be.x; d[1] = (v[i].y - mean) * rstd * g.y + be.y;
            d[2] = (v[i].z - mean) * rstd * g.z + be.z; d[3] = (v[i].w - mean) * rstd * g.w + be.w;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * 16; e += 256) {
        const int tk = e & 15, c = e >> 4, tok = t0 + tk;
        if (tok < n) out[((size_t)b * C + c) * n + tok] = tile[tk][c];
    }
}

int ln_nct_launch(const float* x, int B, int n, int C, const float* gamma, const float* beta, float* out, cudaStream_t s) {
    if (C != 512) return -38;
    ln_nct_kernel<<<dim3((n + 15) / 16, B), 256, 0, s>>>(x, n, gamma, beta, out);
    DSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------ sampler correctors
__global__ void __launch_bounds__(256) clamp_kernel(float* __restrict__ x, long n4, float lo, float hi) {
    pdl_trigger();
    pdl_wait();
    float4* x4 = reinterpret_cast<float4*>(x);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
        float4 v = x4[i];
        v.x = fminf(fmaxf(v.x, lo), hi); v.y = fminf(fmaxf(v.y, lo), hi);
        v.z = fminf(fmaxf(v.z, lo), hi); v.w = fminf(fmaxf(v.w, lo), hi);
        x4[i] = v;
    }
}

int clamp_launch(float* x, long n, float lo, float hi, cudaStream_t s) {
    if (n & 3) return -39;
    long g = (n / 4 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    DSB_PDL_LAUNCH(clamp_kernel, (int)g, 256, 0, s, x, n / 4, lo, hi);
    DSB_LAUNCH_CHECK();
}

// DPM_Solver.dynamic_thresholding_fn (models/dpm_solver/sampler.py:417-426): per clip s = max(quantile(|x0|, p),
// max_val); x0 = clamp(x0, -s, s) / s.  torch.quantile's 'linear' rule: rank = p * (n - 1) in fp32, value =
// lerp(sorted[floor(rank)], sorted[ceil(rank)], frac(rank)); the host passes k = floor(rank) and w = frac(rank).
// One block per clip: exact order statistics by a 4-pass radix select on the bit patterns of |x| (monotone for
// non-negative floats), no sort.
__global__ void __launch_bounds__(1024) dyn_threshold_kernel(float* __restrict__ x, int n, int k, float w, float max_val) {
    pdl_trigger();
    pdl_wait();
    __shared__ unsigned hist[256];
    __shared__ unsigned sh_prefix, sh_rank, sh_eq, sh_next;
    float* xb = x + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    unsigned prefix = 0, mask = 0, rank = (unsigned)k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 1024) {
            const unsigned key = __float_as_uint(fabsf(xb[i]));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned cum = 0;
            int d = 0;
            for (; d < 255; ++d) {
                if (cum + hist[d] > rank) break;
                cum += hist[d];
            }
            sh_prefix = prefix | ((unsigned)d << shift);
            sh_rank = rank - cum;
            sh_eq = hist[d];
            sh_next = 0xffffffffu;
        }
        __syncthreads();
        prefix = sh_prefix;
        rank = sh_rank;
        mask |= 255u << shift;
    }
    // prefix = bits of the k-th smallest |x|; rank = its index among the sh_eq elements equal to it
    unsigned above = prefix;
    if (w != 0.0f && rank + 1 >= sh_eq) {
        unsigned mn = 0xffffffffu;
        for (int i = tid; i < n; i += 1024) {
            const unsigned key = __float_as_uint(fabsf(xb[i]));
            if (key > prefix) mn = min(mn, key);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) atomicMin(&sh_next, mn);
        __syncthreads();
        above = sh_next == 0xffffffffu ? prefix : sh_next;
    }
    const float a = __uint_as_float(prefix), b = __uint_as_float(above);
    // at::lerp: a + w * (b - a) for |w| < 0.5, else b - (b - a) * (1 - w)
    const float diff = __fsub_rn(b, a);
    const float q = (fabsf(w) < 0.5f) ? __fadd_rn(a, __fmul_rn(w, diff)) : __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, w)));
    const float sc = fmaxf(q, max_val);
    for (int i = tid; i < n; i += 1024) xb[i] = __fdiv_rn(fminf(fmaxf(xb[i], -sc), sc), sc);
}

int dyn_threshold_launch(float* x, int B, int n, int k, float w, float max_val, cudaStream_t s) {
    if (k < 0 || k >= n || !(w >= 0.0f && w < 1.0f) || (w != 0.0f && k + 1 >= n)) return -40;
    DSB_PDL_LAUNCH(dyn_threshold_kernel, B, 1024, 0, s, x, n, k, w, max_val);
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb
