// Fused transformer MLP for the narrow stages (C = 96, 192):
//
//     out[m, :] = x1[m, :] + W2 * gelu(W1 * ln2[m, :] + b1) + b2          (common_block.py:125-147, transformer.py:157)
//
// One CTA owns a tile of 128 tokens.  The hidden activation (2C wide) never leaves the SM: it is produced 64 columns at
// a time into a TMEM accumulator, passed through bias + exact-erf GELU by the epilogue warps, written as a bf16 K-major
// SWIZZLE_128B operand tile into shared memory and consumed by the second GEMM, whose accumulator (C columns) stays in
// TMEM for the whole tile.  HBM traffic per token-channel drops from 24 B (ln2 read, hidden write + read, residual
// read, output write, through two kernels) to 10 B.
//
//   warp 0     : TMA producer for the A tile (once per tile) and the W1 chunk ring;  warp 10: W2 chunk ring
//   warp 1     : tcgen05.mma issuer (GEMM1 of chunk h+1 is issued before GEMM2 of chunk h)
//   warps 2..9 : epilogue (stage 1: TMEM -> GELU -> smem operand;  stage 2: TMEM + bias + residual -> global)
//
// The same two-GEMM chain also runs the narrow stages' attention (mode 1): GEMM1 = LayerNormed query tokens times the
// per-frame folded key operand K'_f = scale * K_f Wq (N = 2 heads x 18 keys), "activation" = + folded bias and per-head
// softmax, GEMM2 = P times the per-frame folded value operand V''_f = Wp V_f (K = 64), + proj bias + residual.  Folding
// Wq / Wp into the 18-token K / V operands removes the Q-projection and output-projection GEMMs and keeps Q, the
// scores and the probabilities on chip (attention.py:97-113).
#include "mlp_fused.cuh"

#include <string.h>

namespace dsb {

static constexpr int kMlpThreads = 480;           // warps: 0 A+W1 producer, 1 MMA, 2..9 GELU stage, 10..13 output stage, 14 W2 producer
// (Measured and rejected: a second group of eight GELU-stage warps taking every other hidden chunk.  736 threads cap the
// kernel at 80 registers, the output stage spills its residual prefetch, and the chain kernels ran 1.6x SLOWER: 126 / 123 us
// against 77 / 62 us for the MLP / attention chain of stage 2.)
static constexpr int kHC = 64;                    // hidden columns per chunk
static constexpr uint32_t kAcc2Col = 128;         // TMEM column of the second accumulator (acc1 buffers at 0 and 64)

struct __align__(16) MlpBarriers {
    uint64_t a_full, a_empty;
    uint64_t w1_full[4], w1_empty[4];     // W1 chunk ring (freed as soon as GEMM1 of the chunk has read it)
    uint64_t w2_full[4], w2_empty[4];     // W2 chunk ring (freed after GEMM2 of the chunk)
    uint64_t acc1_full[2], acc1_empty[2];
    uint64_t a2_full[2], a2_empty[2];
    uint64_t acc2_full[2], acc2_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

// MODE and C are compile-time (four instantiations): every loop bound, operand size and descriptor step below is then a
// constant, as in the specialised epilogues of gemm_tc.cu
template <int MODE, int C>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const MlpParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const int NS) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int bk = (C % 64 == 0) ? 64 : 32;
    constexpr int ksub = C / bk;                   // K sub-blocks of GEMM1
    constexpr int NC = MODE ? 1 : 2 * C / kHC;     // hidden chunks per tile (attention: one chunk of 2 x 18 (+28) keys)
    const uint32_t a_sub = 128u * bk * 2u;
    const uint32_t a_bytes = a_sub * ksub;
    const uint32_t w1_sub = (uint32_t)kHC * bk * 2u;
    const uint32_t w1_bytes = w1_sub * ksub;
    const uint32_t w2_bytes = (uint32_t)C * kHC * 2u;
    const uint32_t a2_bytes = 128u * kHC * 2u;
    uint8_t* sA = smem;
    uint8_t* sA2 = sA + a_bytes;
    uint8_t* sW1 = sA2 + 2 * a2_bytes;
    uint8_t* sW2 = sW1 + (size_t)NS * w1_bytes;
    MlpBarriers* bars = reinterpret_cast<MlpBarriers*>(sW2 + (size_t)NS * w2_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars->w1_full[i], 1);
            mbar_init(&bars->w1_empty[i], 1);
            mbar_init(&bars->w2_full[i], 1);
            mbar_init(&bars->w2_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->acc1_full[i], 1);
            mbar_init(&bars->acc1_empty[i], 8);       // one arrive per GELU-stage warp
            mbar_init(&bars->a2_full[i], 8);
            mbar_init(&bars->a2_empty[i], 1);
            mbar_init(&bars->acc2_full[i], 1);
            mbar_init(&bars->acc2_empty[i], 4);       // one arrive per output-stage warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_trigger();                                    // PDL (common.cuh): resources held -> the next kernel may queue up
    pdl_wait();                                       // ... and nothing below runs before the previous kernel is done

    const int tiles_per_frame = (p.HW + 127) >> 7;
    const int total_tiles = tiles_per_frame * p.F;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        uint32_t g = 0, it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int f = tile / tiles_per_frame;
            const int x0 = (tile % tiles_per_frame) << 7;
            const int fs = p.f_group ? (f / p.f_used) * p.f_group + f % p.f_used : f;
            mbar_wait(&bars->a_empty, (it & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(&bars->a_full, a_bytes);
                for (int j = 0; j < ksub; ++j) tma_load_3d(sA + j * a_sub, &tmA, &bars->a_full, j * bk, x0, fs);
            }
            __syncwarp();
            for (int h = 0; h < NC; ++h, ++g) {
                const uint32_t s = g % (uint32_t)NS, ph = (g / (uint32_t)NS) & 1u;
                mbar_wait(&bars->w1_empty[s], ph ^ 1u);
                if (elect_one()) {
                    uint8_t* w = sW1 + s * w1_bytes;
                    mbar_expect_tx(&bars->w1_full[s], w1_bytes);
                    for (int j = 0; j < ksub; ++j)
                        tma_load_2d(w + j * w1_sub, &tmW1, &bars->w1_full[s], j * bk, h * kHC + (MODE ? fs * kHC : 0));
                }
                __syncwarp();
            }
        }
    } else if (warp == 14) {
        // ------------------------------------------------------------------ W2 producer (its ring drains later than W1's)
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int f = tile / tiles_per_frame;
            const int fs = p.f_group ? (f / p.f_used) * p.f_group + f % p.f_used : f;
            for (int h = 0; h < NC; ++h, ++g) {
                const uint32_t s = g % (uint32_t)NS, ph = (g / (uint32_t)NS) & 1u;
                mbar_wait(&bars->w2_empty[s], ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(&bars->w2_full[s], w2_bytes);
                    tma_load_2d(sW2 + s * w2_bytes, &tmW2, &bars->w2_full[s], h * kHC, MODE ? fs * C : 0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc1 = umma_idesc_bf16(kHC);
        const uint32_t idesc2 = umma_idesc_bf16((uint32_t)C);
        const uint32_t rb1 = (uint32_t)bk * 2u;
        const uint64_t dA = umma_smem_desc(smem_u32(sA), rb1);
        const uint64_t dW1 = umma_smem_desc(smem_u32(sW1), rb1);
        const uint64_t dW2 = umma_smem_desc(smem_u32(sW2), 128u);
        const uint64_t dA2 = umma_smem_desc(smem_u32(sA2), 128u);
        constexpr int k1 = bk / 16;
        uint32_t gbase = 0, it = 0;

        auto issue_g1 = [&](int h) {
            const uint32_t gg = gbase + (uint32_t)h, s = gg & 1u, ph = (gg >> 1) & 1u;
            const uint32_t ws = gg % (uint32_t)NS, wph = (gg / (uint32_t)NS) & 1u;
            mbar_wait(&bars->w1_full[ws], wph);
            if (h == 0) mbar_wait(&bars->a_full, it & 1u);
            mbar_wait(&bars->acc1_empty[s], ph ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + s * kHC;
                for (int j = 0; j < ksub; ++j) {
                    const uint64_t da = dA + (uint64_t)((j * a_sub) >> 4);
                    const uint64_t dw = dW1 + (uint64_t)((ws * w1_bytes + j * w1_sub) >> 4);
                    for (int k = 0; k < k1; ++k) umma_bf16(d, da + 2 * k, dw + 2 * k, idesc1, (j | k) ? 1u : 0u);
                }
                umma_commit(&bars->w1_empty[ws]);
                umma_commit(&bars->acc1_full[s]);
                if (h == NC - 1) umma_commit(&bars->a_empty);      // the A tile is free once every GEMM1 has read it
            }
            __syncwarp();
        };

        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            issue_g1(0);
            for (int h = 0; h < NC; ++h) {
                if (h + 1 < NC) issue_g1(h + 1);
                const uint32_t gg = gbase + (uint32_t)h, s = gg & 1u, ph = (gg >> 1) & 1u;
                const uint32_t ws = gg % (uint32_t)NS, wph = (gg / (uint32_t)NS) & 1u;
                const uint32_t ab = it & 1u;                               // acc2 buffer of this tile
                if (h == 0) mbar_wait(&bars->acc2_empty[ab], ((it >> 1) & 1u) ^ 1u);
                mbar_wait(&bars->w2_full[ws], wph);
                mbar_wait(&bars->a2_full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + kAcc2Col + ab * (uint32_t)C;
                    const uint64_t da = dA2 + (uint64_t)((s * a2_bytes) >> 4);
                    const uint64_t dw = dW2 + (uint64_t)((ws * w2_bytes) >> 4);
#pragma unroll
                    for (int k = 0; k < kHC / 16; ++k) umma_bf16(d, da + 2 * k, dw + 2 * k, idesc2, (h | k) ? 1u : 0u);
                    umma_commit(&bars->w2_empty[ws]);
                    umma_commit(&bars->a2_empty[s]);
                    if (h == NC - 1) umma_commit(&bars->acc2_full[ab]);
                }
                __syncwarp();
            }
            gbase += (uint32_t)NC;
        }
    } else if (warp < 10) {
        // ------------------------------------------------------------------ GELU stage (warps 2..9)
        // hidden chunk: TMEM -> + b1 -> exact-erf GELU -> bf16 -> K-major SWIZZLE_128B operand tile of GEMM2
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t gg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int f_t = tile / tiles_per_frame;
            const int fs_row = p.f_group ? (f_t / p.f_used) * p.f_group + f_t % p.f_used : f_t;
            for (int h = 0; h < NC; ++h, ++gg) {
                const uint32_t s = gg & 1u, ph = (gg >> 1) & 1u;
                mbar_wait(&bars->acc1_full[s], ph);
                tc_fence_after();
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_addr + s * kHC + half * 32, raw);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc1_empty[s]);
                uint32_t packed[16];
                if (MODE == 0) {
                    const float* b1 = p.b1 + h * kHC + half * 32;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 bb = __ldg(reinterpret_cast<const float2*>(b1) + j);
                        const float2 v = gelu_erf2(__fadd2_rn(make_float2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1])), bb));
                        packed[j] = pack_bf16x2(v.x, v.y);
                    }
                } else {
                    // scores of head 0 live in columns 0..17, head 1 in 18..35: the warp with half == 0 holds columns
                    // 0..31, the other 32..63; the four head-1 scores 32..35 are exchanged through shared memory
                    // (sx: [8 warps][32 rows][4]) so that each thread can normalise its own columns
                    float v[32];
                    const float* rb = p.b1 + (size_t)fs_row * kHC + half * 32;        // folded bias, per frame
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) + __ldg(rb + j);
                    float* sx = reinterpret_cast<float*>(bars + 1);                     // [2 halves][128 rows][4]
                    // half 0 publishes (max, sum-exp partials need both halves): do it in two passes
                    // pass 1: per-head partial maxima
                    float m0 = -INFINITY, m1 = -INFINITY;
                    if (half == 0) {
#pragma unroll
                        for (int j = 0; j < 18; ++j) m0 = fmaxf(m0, v[j]);
#pragma unroll
                        for (int j = 18; j < 32; ++j) m1 = fmaxf(m1, v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) m1 = fmaxf(m1, v[j]);
                    }
                    sx[(half * 128 + row) * 4 + 0] = m1;
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    m1 = fmaxf(m1, sx[((half ^ 1) * 128 + row) * 4 + 0]);
                    // pass 2: exponentials and per-head partial sums
                    float s0 = 0.0f, s1 = 0.0f;
                    if (half == 0) {
#pragma unroll
                        for (int j = 0; j < 18; ++j) { v[j] = __expf(v[j] - m0); s0 += v[j]; }
#pragma unroll
                        for (int j = 18; j < 32; ++j) { v[j] = __expf(v[j] - m1); s1 += v[j]; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { v[j] = __expf(v[j] - m1); s1 += v[j]; }
#pragma unroll
                        for (int j = 4; j < 32; ++j) v[j] = 0.0f;
                    }
                    sx[(half * 128 + row) * 4 + 1] = s1;
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    s1 += sx[((half ^ 1) * 128 + row) * 4 + 1];
                    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
                    if (half == 0) {
#pragma unroll
                        for (int j = 0; j < 18; ++j) v[j] *= i0;
#pragma unroll
                        for (int j = 18; j < 32; ++j) v[j] *= i1;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] *= i1;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                }
                mbar_wait(&bars->a2_empty[s], ph ^ 1u);
                uint8_t* dst = sA2 + s * a2_bytes + row * 128;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int piece = half * 4 + jj;                      // 16-byte piece inside the 128-byte row
                    *reinterpret_cast<uint4*>(dst + ((piece ^ (row & 7)) << 4)) =
                        make_uint4(packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a2_full[s]);
            }
        }
    } else if (warp < 14) {
        // ------------------------------------------------------------------ output stage (warps 10..13)
        // out = acc2 + b2 + residual, overlapped with the GELU stage of the next tile (acc2 is double-buffered).
        // Each 32-row x 32-column chunk is transposed through a padded smem tile so that one warp instruction touches
        // whole 128-byte rows (4 rows per instruction): residual loads and output stores are fully coalesced.  The
        // residual of chunk c + 1 is requested before chunk c is processed.
        const int q = warp & 3;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        constexpr int nch = C / 32;
        float* tile = reinterpret_cast<float*>(bars + 1) + 1024 + (warp - 10) * (32 * 36);   // behind the 4 KB softmax exchange
        const int cq = (lane & 7) * 4;                      // this lane's 4 columns inside a chunk
        const int rq = lane >> 3;                           // row offset inside a group of 4 rows
        uint32_t it = 0;
        for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x, ++it) {
            const int f = tile_i / tiles_per_frame;
            const int x0 = ((tile_i % tiles_per_frame) << 7) + q * 32;      // first token (in the frame) of this warp's rows
            const int fs = p.f_group ? (f / p.f_used) * p.f_group + f % p.f_used : f;
            const int nvalid = min(32, p.HW - x0);                        // <= 0: no valid row
            const size_t tok0 = (size_t)fs * p.HW + x0;
            const uint32_t ab = it & 1u;
            const float* rbase = p.residual + tok0 * C + cq;
            float* obase = p.out + tok0 * C + cq;
            float4 res[2][8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = k * 4 + rq;
                res[0][k] = row < nvalid ? *reinterpret_cast<const float4*>(rbase + (size_t)row * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&bars->acc2_full[ab], (it >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < nch; cc += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c = cc + u;
                    if (c >= nch) break;
                    if (c + 1 < nch) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int row = k * 4 + rq;
                            res[(u + 1) & 1][k] = row < nvalid ? *reinterpret_cast<const float4*>(rbase + (size_t)row * C + (c + 1) * 32)
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    uint32_t raw[32];
                    tmem_ld32(tmem_base + lane_addr + kAcc2Col + ab * (uint32_t)C + c * 32, raw);
                    tmem_ld_wait();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(tile + lane * 36 + 4 * j) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
                    __syncwarp();
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.b2 + c * 32 + cq));
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int row = k * 4 + rq;
                        if (row < nvalid) {
                            const float4 v = *reinterpret_cast<const float4*>(tile + row * 36 + cq);
                            const float4 r = res[u][k];
                            *reinterpret_cast<float4*>(obase + (size_t)row * C + c * 32) =
                                make_float4(v.x + b.x + r.x, v.y + b.y + r.y, v.z + b.z + r.z, v.w + b.w + r.w);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc2_empty[ab]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int mlp_fused_lower(const MlpOp& op, MlpLaunch* out) {
    memset(out, 0, sizeof(*out));
    const int C = op.C;
    if (C != 96 && C != 192) return -40;
    MlpParams& p = out->p;
    p.mode = op.mode;
    p.C = C;
    p.bk = (C % 64 == 0) ? 64 : 32;
    p.HW = op.HW; p.F = op.F; p.f_group = op.f_group; p.f_used = op.f_used;
    p.b1 = op.b1; p.b2 = op.b2; p.residual = op.residual; p.out = op.out;
    const uint64_t e = 2;
    const int src_frames = op.f_group ? ((op.F + op.f_used - 1) / op.f_used) * op.f_group : op.F;
    {
        uint64_t dims[3] = {(uint64_t)C, (uint64_t)op.HW, (uint64_t)src_frames};
        uint64_t str[2] = {C * e, (uint64_t)op.HW * C * e};
        uint32_t box[3] = {(uint32_t)p.bk, 128u, 1u};
        if (int r = make_tensor_map(&out->tmA, op.A, 3, dims, str, box)) return r;
    }
    {
        // MLP: fc1.weight [2C][C];  attention: folded keys [src frames * 64][C]
        uint64_t dims[2] = {(uint64_t)C, op.mode ? (uint64_t)src_frames * kHC : (uint64_t)(2 * C)};
        uint64_t str[1] = {C * e};
        uint32_t box[2] = {(uint32_t)p.bk, (uint32_t)kHC};
        if (int r = make_tensor_map(&out->tmW1, op.W1, 2, dims, str, box)) return r;
    }
    {
        // MLP: fc2.weight [C][2C];  attention: folded values [src frames * C][64]
        uint64_t dims[2] = {op.mode ? (uint64_t)kHC : (uint64_t)(2 * C), op.mode ? (uint64_t)src_frames * C : (uint64_t)C};
        uint64_t str[1] = {(op.mode ? (uint64_t)kHC : (uint64_t)(2 * C)) * e};
        uint32_t box[2] = {(uint32_t)kHC, (uint32_t)C};
        if (int r = make_tensor_map(&out->tmW2, op.W2, 2, dims, str, box)) return r;
    }
    return 0;
}

int mlp_fused_run(const MlpLaunch& l, int num_sms, cudaStream_t stream) {
    const MlpParams& p = l.p;
    const int ksub = p.C / p.bk;
    const size_t a_bytes = (size_t)128 * p.bk * 2 * ksub;
    const size_t w_stage = (size_t)kHC * p.bk * 2 * ksub + (size_t)p.C * kHC * 2;
    const size_t fixed = a_bytes + 2 * (128 * kHC * 2) + sizeof(MlpBarriers) + 4096 + 4 * 32 * 36 * 4 + 1024;   // + softmax exchange + output transpose tiles
    int NS = (int)((230000 - fixed) / w_stage);
    if (NS > 4) NS = 4;
    if (NS < 2) return -42;
    const size_t smem = fixed + (size_t)NS * w_stage;
    static bool attr = false;
    if (!attr) {
        const void* fns[4] = {(const void*)mlp_fused_kernel<0, 96>, (const void*)mlp_fused_kernel<0, 192>,
                              (const void*)mlp_fused_kernel<1, 96>, (const void*)mlp_fused_kernel<1, 192>};
        for (const void* f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
        }
        attr = true;
    }
    const int tiles = ((p.HW + 127) / 128) * p.F;
    int grid = tiles < num_sms ? tiles : num_sms;
    if (grid < 1) return -41;
    if (p.mode == 0 && p.C == 96) return (int)launch_pdl_ex(true, mlp_fused_kernel<0, 96>, dim3(grid), dim3(kMlpThreads), smem, stream, p, l.tmA, l.tmW1, l.tmW2, NS);
    if (p.mode == 0 && p.C == 192) return (int)launch_pdl_ex(true, mlp_fused_kernel<0, 192>, dim3(grid), dim3(kMlpThreads), smem, stream, p, l.tmA, l.tmW1, l.tmW2, NS);
    if (p.mode == 1 && p.C == 96) return (int)launch_pdl_ex(true, mlp_fused_kernel<1, 96>, dim3(grid), dim3(kMlpThreads), smem, stream, p, l.tmA, l.tmW1, l.tmW2, NS);
    if (p.mode == 1 && p.C == 192) return (int)launch_pdl_ex(true, mlp_fused_kernel<1, 192>, dim3(grid), dim3(kMlpThreads), smem, stream, p, l.tmA, l.tmW1, l.tmW2, NS);
    return -40;
}

}  // namespace dsb
