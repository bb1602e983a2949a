#include "conv_plan.cuh"

#include <stdlib.h>
#include <string.h>

namespace dsb {

static int pick_bn(int N) {
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    for (int bn = 256; bn >= 32; bn -= 32)               // the general epilogues work on 32-column chunks
        if (N % bn == 0) return bn;
    for (int bn = 240; bn >= 16; bn -= 16)               // 48-column attention-score tiles
        if (N % bn == 0) return bn;
    return 0;
}

// Choose the 128-pixel tile box (bw x bh x bf, powers of two) that wastes the fewest accumulator rows.
static void pick_tile(int W, int H, int F, bool single_frame, int* bw_log2, int* bh_log2) {
    long best = -1;
    int bx = 7, by = 0;
    for (int lx = 7; lx >= 0; --lx) {
        for (int ly = 0; lx + ly <= 7; ++ly) {
            if (single_frame && lx + ly != 7) continue;
            const int bw = 1 << lx, bh = 1 << ly, bf = 128 >> (lx + ly);
            const long tiles = (long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((F + bf - 1) / bf);
            if (best < 0 || tiles < best) {
                best = tiles;
                bx = lx;
                by = ly;
            }
        }
    }
    *bw_log2 = bx;
    *bh_log2 = by;
}

int conv_lower(const ConvOp& op, ConvLaunch* out) {
    memset(out, 0, sizeof(*out));
    GemmParams& p = out->p;
    const int C = op.Cin;
    const int bk = (C % 64 == 0) ? 64 : ((C % 32 == 0) ? 32 : 0);
    if (!bk) return -20;
    int bn = pick_bn(op.N);
    if (!bn) return -21;
    p.W = op.W; p.H = op.H; p.F = op.F;
    pick_tile(op.W, op.H, op.F, op.b_rows_per_frame != 0 || op.f_group != 0, &p.bw_log2, &p.bh_log2);
    const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2, bf = 128 >> (p.bw_log2 + p.bh_log2);
    p.tiles_x = (op.W + bw - 1) / bw;
    p.tiles_y = (op.H + bh - 1) / bh;
    p.tiles_f = (op.F + bf - 1) / bf;
    p.cin_blocks = C / bk;
    p.bk = bk;
    p.N = op.N;
    p.bn = bn;
    p.ydim = 2;
    // K sub-blocks per stage: long stages amortise the barrier round trip of the producer / MMA-issue loops.  A per-role
    // clock64 trace of upembed.conv1 of the last stage (N = 96, one 22.5 KB sub-block per stage) showed ~500 cycles per
    // stage against 192 cycles of MMA; three sub-blocks per stage (67.5 KB, 3-deep ring) took it from 175 to 135 us.
    // Swept on the whole evaluation: <= 72 KB per stage and >= 768 tensor cycles was the best setting.
    auto pick_ksub = [&](bool pair) {
        const int n_eff = pair ? bn / 2 : bn;
        const int sub_bytes = (128 + n_eff) * bk * 2;
        const int sub_cycles = (bk / 16) * (bn / 2);
        int ks = 1;
        for (int cand = 1; cand <= p.cin_blocks; ++cand) {
            if (p.cin_blocks % cand) continue;
            if (cand * sub_bytes > 73728) break;
            ks = cand;
            if (cand * sub_cycles >= 768) break;
        }
        return ks;
    };
    const int taps_ = (op.kind == CONV_3X3 || op.kind == CONV_3X3_S2) ? 9 : (op.kind == CONV_TEMPORAL ? op.kt : 1);
    const bool pair_ok = !op.b_rows_per_frame && !op.out_softmax && (bn / 2) % 8 == 0;

    // ---- split-K plan (opt-in): few tiles x long K (the encoder's deep convs, ReduceTemp of the small stages).
    // Everything here is a function of the layer geometry at the NOMINAL frame count only -- tile box, pairing, stage
    // size, slice count -- so the K partition, and with it the rounding, never depends on the actual batch size.
    int S_plan = 1;
    if (op.split_ws && !op.b_rows_per_frame && !op.out_softmax && !op.head_w && !op.f_group) {
        const int f_nom = op.split_frames_nominal > 0 ? op.split_frames_nominal : op.F;
        int nbw = 0, nbh = 0;
        pick_tile(op.W, op.H, f_nom, false, &nbw, &nbh);
        const int bf_ = 128 >> (nbw + nbh);
        const long tiles = (long)((op.W + (1 << nbw) - 1) >> nbw) * ((op.H + (1 << nbh) - 1) >> nbh) * ((f_nom + bf_ - 1) / bf_) *
                           (op.N / bn);
        if (!(tiles >= 148 && pair_ok)) {                    // the nominal launch would not run as CTA pairs
            const int nk = taps_ * p.cin_blocks / pick_ksub(false);
            for (int cand = 2; cand <= 16; ++cand) {
                if (nk % cand) continue;
                if (tiles * cand > 148) break;
                if (nk / cand < 3) break;                    // keep at least 3 pipeline stages of work per slice
                S_plan = cand;
            }
        }
        if (S_plan > 1 && (long)op.F * op.H * op.W * op.N * S_plan > op.split_ws_elems) return -24;   // scratch too small
    }
    if (S_plan > 1) {
        p.two_cta = 0;                                       // slices run on single CTAs at every batch size
    } else {
        const long tiles = (long)p.tiles_x * p.tiles_y * p.tiles_f * (op.N / bn);
        p.two_cta = (op.two_cta > 0 || (op.two_cta == 0 && tiles >= 148)) && pair_ok ? 1 : 0;
    }
    p.ksub = pick_ksub(p.two_cta != 0);

    // ---- halo-tile path (opt-in): the nine taps of a 3x3 conv re-read every input pixel from L2 nine times; with a small N
    // the launch is bound by the chip-wide L2 throughput (~43 B/cycle/SM), not by the tensor pipe.  One TMA box per
    // (8x16-pixel tile, 64-channel block) holding the tile plus its dilation halo replaces the nine per-tap boxes (23-31 KB
    // instead of 144 KB); the taps become row-shifted UMMA descriptors into it (tests: test_umma_row_shifted_descriptor).
    bool halo = false;
    if (op.halo && op.kind == CONV_3X3 && S_plan == 1 && !op.b_rows_per_frame && !op.out_softmax && op.two_cta >= 0 &&
        !op.f_group && op.dilation >= 1 && op.dilation <= 4) {
        // N tile: at most 128 columns (a stage holds the nine weight tiles of a channel block); wider layers are cut into
        // up to 4 N tiles, each of which re-reads the (cheap) halo tiles
        // two stages must fit beside the epilogue's shared memory (61 KB when the row-coalesced epilogue is used)
        const bool epi_t = op.out_f32 && !op.out_bf16 && !op.head_w;
        const long budget = 225L * 1024 - 2048 - (epi_t ? 12L * (32 * 36 + 128) * 4 : 0) - 3 * 128 * 4;
        const long a_h = (((long)(8 + 2 * op.dilation) * (16 + 2 * op.dilation) * bk * 2) + 1023) & ~1023L;
        int bn_h = 0;
        for (int cand = 128; cand >= 32; cand -= 32) {
            if (op.N % cand || op.N / cand > 4) continue;
            if (2 * (a_h + 9L * (cand / 2) * bk * 2) > budget) continue;
            bn_h = cand;
            break;
        }
        if (op.head_w && bn_h != op.N) bn_h = 0;             // the fused head needs the whole row in one tile
        const long tiles8 = bn_h ? (long)((op.W + 7) / 8) * ((op.H + 15) / 16) * op.F * (op.N / bn_h) : 0;
        const long rows_box = (long)((op.W + 7) / 8) * 8 * ((op.H + 15) / 16) * 16;
        // worth it when the 8x16 tiling wastes few rows and there are enough tiles for CTA pairs
        if (bn_h && tiles8 >= 148 && rows_box * 100 <= (long)op.W * op.H * 120) {
            halo = true;
            bn = bn_h;
            p.bn = bn;
            p.halo = 1;
            p.halo_d = op.dilation;
            p.halo_pw = 8 + 2 * op.dilation;
            p.halo_ph = 16 + 2 * op.dilation;
            p.bw_log2 = 3; p.bh_log2 = 4;
            p.tiles_x = (op.W + 7) / 8; p.tiles_y = (op.H + 15) / 16; p.tiles_f = op.F;
            p.two_cta = 1;
            p.ksub = 1;
        }
    }

    uint64_t dims[5], strides[4];
    uint32_t box[5];
    const uint64_t e = 2;   // bytes per bf16
    if (op.kind == CONV_3X3) {
        p.taps = 9;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                int* o = p.tap_off[dy * 3 + dx];
                o[0] = 0; o[1] = (dx - 1) * op.dilation; o[2] = (dy - 1) * op.dilation; o[3] = 0;
            }
        dims[0] = C; dims[1] = op.W; dims[2] = op.H; dims[3] = 1; dims[4] = op.F;
        strides[0] = C * e; strides[1] = (uint64_t)op.W * C * e; strides[2] = (uint64_t)op.H * op.W * C * e;
        strides[3] = (uint64_t)op.H * op.W * C * e;
        box[0] = bk; box[1] = bw; box[2] = bh; box[3] = 1; box[4] = bf;
        if (halo) { box[1] = p.halo_pw; box[2] = p.halo_ph; box[4] = 1; }
    } else if (op.kind == CONV_1X1) {
        p.taps = 1;
        dims[0] = C; dims[1] = op.W; dims[2] = op.H; dims[3] = 1; dims[4] = op.F;
        strides[0] = C * e; strides[1] = (uint64_t)op.W * C * e; strides[2] = (uint64_t)op.H * op.W * C * e;
        strides[3] = (uint64_t)op.H * op.W * C * e;
        box[0] = bk; box[1] = bw; box[2] = bh; box[3] = 1; box[4] = bf;
    } else if (op.kind == CONV_3X3_S2) {
        // input [F, 2H, 2W, C] viewed as [F][H][2][W][2C]: x_in = 2x + dx -> (x + dx/2, parity dx%2)
        p.taps = 9;
        p.ydim = 3;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                int* o = p.tap_off[dy * 3 + dx];
                o[0] = (dx & 1) * C; o[1] = dx >> 1; o[2] = dy & 1; o[3] = dy >> 1;
            }
        const uint64_t Win = 2 * (uint64_t)op.W, Hin = 2 * (uint64_t)op.H;
        dims[0] = 2 * C; dims[1] = op.W; dims[2] = 2; dims[3] = op.H; dims[4] = op.F;
        strides[0] = 2 * C * e; strides[1] = Win * C * e; strides[2] = 2 * Win * C * e; strides[3] = Hin * Win * C * e;
        box[0] = bk; box[1] = bw; box[2] = 1; box[3] = bh; box[4] = bf;
    } else if (op.kind == CONV_TEMPORAL) {
        if (op.kt < 1 || op.kt > 9 || op.kt > op.T) return -22;
        p.taps = op.kt;
        for (int t = 0; t < op.kt; ++t) {
            int* o = p.tap_off[t];
            o[0] = 0; o[1] = 0; o[2] = 0; o[3] = t;
        }
        dims[0] = C; dims[1] = op.W; dims[2] = op.H; dims[3] = op.T; dims[4] = op.F;
        strides[0] = C * e; strides[1] = (uint64_t)op.W * C * e; strides[2] = (uint64_t)op.H * op.W * C * e;
        strides[3] = (uint64_t)op.T * op.H * op.W * C * e;
        box[0] = bk; box[1] = bw; box[2] = bh; box[3] = 1; box[4] = bf;
    } else {
        return -23;
    }
    if (op.f_group) dims[4] = (uint64_t)((op.F + op.f_used - 1) / op.f_used) * op.f_group;   // source frames
    int r = make_tensor_map(&out->tmA, op.A, 5, dims, strides, box);
    if (r) return r;
    const uint64_t K = (uint64_t)p.taps * C;
    const uint64_t src_frames = op.f_group ? (uint64_t)((op.F + op.f_used - 1) / op.f_used) * op.f_group : (uint64_t)op.F;
    const uint64_t wrows = op.b_rows_per_frame ? src_frames * op.b_rows_per_frame : (uint64_t)op.N;
    uint64_t wd[2] = {K, wrows};
    uint64_t ws[1] = {K * e};
    uint32_t wb[2] = {(uint32_t)bk, (uint32_t)(p.two_cta ? bn / 2 : bn)};
    r = make_tensor_map(&out->tmB, op.Wt, 2, wd, ws, wb);
    if (r) return r;

    p.scale = op.scale; p.shift = op.shift; p.rowbias = op.rowbias; p.residual = op.residual;
    p.act = op.act;
    p.ab_f16 = op.ab_f16;
    p.out_f16 = op.out_f16;
    p.out_f32 = op.out_f32; p.out_bf16 = op.out_bf16;
    p.ldo = op.ldo ? op.ldo : op.N;
    p.out_fmul = op.out_fmul ? op.out_fmul : 1;
    p.out_fadd = op.out_fadd;
    p.out2_f32 = op.out2_f32; p.out2_fmul = op.out2_fmul ? op.out2_fmul : 1; p.out2_fadd = op.out2_fadd;
    p.head_w = op.head_w; p.head_b = op.head_b; p.out_head = op.out_head;
    p.b_rows_per_frame = op.b_rows_per_frame;
    p.out_softmax = op.out_softmax;
    // fp32 outputs (with or without an fp32 residual) go through the smem-transposed, row-coalesced epilogue; bf16
    // outputs keep the direct thread-per-row epilogue (measured: upembed.conv1 +7 %, fc1 +18 % slower when transposed)
    p.epi_transposed = (op.out_f32 && !op.out_bf16 && !op.head_w && !op.out_softmax && bn % 32 == 0) ? 1 : 0;
    p.f_group = op.f_group; p.f_used = op.f_used; p.out_remap = op.out_remap;
    p.kv_mode = op.kv_mode; p.kv_R = op.kv_R; p.kv_C = op.kv_C;
    p.trace = op.trace;
    if (op.kv_mode && (!op.out_bf16 || op.out_f32 || S_plan > 1 || op.kind != CONV_1X1 || op.kv_C % 16 || op.N != 2 * op.kv_C ||
                       op.kv_R < 36 || op.H != 1 || op.F != 1 || op.W % 18))
        return -26;

    if (S_plan > 1) {
        const int S = S_plan;
        const long rows = (long)op.F * op.H * op.W;
        SplitReduce& r = out->split;
        r.ws = op.split_ws; r.S = S; r.slab = rows * op.N;
        r.rows = (int)rows; r.N = op.N; r.HW = op.H * op.W;
        r.scale = p.scale; r.shift = p.shift; r.rowbias = p.rowbias; r.residual = p.residual; r.act = p.act;
        r.out_f32 = p.out_f32; r.out_bf16 = p.out_bf16; r.out2_f32 = p.out2_f32;
        r.ldo = p.ldo; r.out_fmul = p.out_fmul; r.out_fadd = p.out_fadd; r.out2_fmul = p.out2_fmul; r.out2_fadd = p.out2_fadd;
        // the GEMM pass writes raw partials, compact rows, plain direct epilogue
        p.scale = p.shift = p.rowbias = p.residual = nullptr;
        p.act = ACT_NONE;
        p.out_f32 = op.split_ws; p.out_bf16 = nullptr; p.out2_f32 = nullptr;
        p.ldo = op.N; p.out_fmul = 1; p.out_fadd = 0; p.out2_fmul = 1; p.out2_fadd = 0;
        p.epi_transposed = 0;
        p.ksplit = S;
        p.split_stride = r.slab;
    }
    return 0;
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(SplitReduce r) {
    pdl_trigger();
    pdl_wait();
    const int nv = r.N >> 2;
    const long total = (long)r.rows * nv;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int row = (int)(i / nv), n = (int)(i % nv) * 4;
        const int f = row / r.HW, pix = row - f * r.HW;
        const float4* src = reinterpret_cast<const float4*>(r.ws + (size_t)row * r.N + n);
        float4 v = src[0];
        for (int s = 1; s < r.S; ++s) {                      // fixed order: bitwise reproducible
            const float4 t = *reinterpret_cast<const float4*>(r.ws + (size_t)s * r.slab + (size_t)row * r.N + n);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (r.scale) { const float4 t = __ldg(reinterpret_cast<const float4*>(r.scale + n)); v.x *= t.x; v.y *= t.y; v.z *= t.z; v.w *= t.w; }
        if (r.shift) { const float4 t = __ldg(reinterpret_cast<const float4*>(r.shift + n)); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        if (r.rowbias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(r.rowbias + (size_t)f * r.N + n));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (r.act == ACT_RELU) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        } else if (r.act == ACT_GELU) {
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        }
        if (r.residual) {
            const float4 t = *reinterpret_cast<const float4*>(r.residual + (size_t)row * r.N + n);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        const size_t po = ((size_t)(f * r.out_fmul + r.out_fadd) * r.HW + pix) * r.ldo + n;
        if (r.out_f32) *reinterpret_cast<float4*>(r.out_f32 + po) = v;
        if (r.out_bf16) *reinterpret_cast<uint2*>(r.out_bf16 + po) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        if (r.out2_f32) {
            const size_t p2 = ((size_t)(f * r.out2_fmul + r.out2_fadd) * r.HW + pix) * r.ldo + n;
            *reinterpret_cast<float4*>(r.out2_f32 + p2) = v;
        }
    }
}

int conv_run(const ConvLaunch& l, int num_sms, cudaStream_t stream) {
    if (int r = gemm_launch(l.p, l.tmA, l.tmB, num_sms, stream)) return r;
    if (l.split.S > 1) {
        long g = ((long)l.split.rows * (l.split.N >> 2) + 255) / 256;
        if (g > 148 * 8) g = 148 * 8;
        pdl_allow_next() = true;                       // follows its own GEMM on the same stream
        return (int)launch_pdl(splitk_reduce_kernel, dim3((int)g), dim3(256), 0, stream, l.split);
    }
    return 0;
}

}  // namespace dsb
