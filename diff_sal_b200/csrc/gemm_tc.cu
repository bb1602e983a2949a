// tcgen05 / TMEM / TMA implicit-GEMM kernel (see gemm_tc.cuh).  Warp-specialised, persistent:
//   warp 0      : TMA producer (one lane)           smem ring  full[] / empty[]
//   warp 1      : tcgen05.mma issuer (one lane)      TMEM double buffer  tmem_full[] / tmem_empty[]
//   warps 2..9  : epilogue, two per TMEM lane quadrant (tcgen05.ld -> registers -> fused epilogue -> global)
#include "gemm_tc.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace dsb {

static constexpr int kEpiWarps = 12;           // kPerQuad warps per TMEM lane quadrant, splitting the column chunks
static constexpr int kPerQuad = kEpiWarps / 4;
static constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
static constexpr int kMaxStages = 8;
static constexpr uint32_t kAccStride = 256;   // TMEM columns between the two accumulator buffers
static constexpr uint32_t kTmemCols = 512;

struct __align__(16) GemmBarriers {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// Compile-time epilogue flavours.  The generic kernel (EPI = 0) tests every epilogue option at run time inside its
// innermost loops: the row-coalesced store loop came to ~135 SASS instructions per (4 rows x 128 B) store and the twelve
// epilogue warps saturated the issue slots -- a per-role clock trace (tools/gemm_trace.py) showed 3.4 k cycles per
// 32-column chunk, 2-5x the MMA time of the mid-size GEMMs.  A specialised instantiation has bit 0 set and states the
// options in the other bits, so its loops hold only the work of that flavour.
enum : int {
    kEpiSpecial = 1, kEpiTransposed = 2, kEpiScale = 4, kEpiShift = 8, kEpiRowbias = 16, kEpiResidual = 32,
    kEpiRelu = 64, kEpiGelu = 128, kEpiOut32 = 256, kEpiOut16 = 512, kEpiOut2 = 1024, kEpiKvK = 2048, kEpiKvV = 4096
};

template <bool TWO, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const int num_stages) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // two_cta: the CTA pair of a cluster shares one M = 256 tcgen05.mma (cta_group::2); each CTA stages its own 128
    // A rows and HALF of the B rows, which halves the shared-memory operand traffic per MMA
    constexpr bool two = TWO;            // separate instantiations: the 1-CTA kernel holds no cta_group::2 code
    constexpr bool SP = (EPI & kEpiSpecial) != 0;
#define F_SCALE (SP ? (EPI & kEpiScale) != 0 : p.scale != nullptr)
#define F_SHIFT (SP ? (EPI & kEpiShift) != 0 : p.shift != nullptr)
#define F_ROWBIAS (SP ? (EPI & kEpiRowbias) != 0 : p.rowbias != nullptr)
#define F_RES (SP ? (EPI & kEpiResidual) != 0 : p.residual != nullptr)
#define F_ACT (SP ? ((EPI & kEpiRelu) ? (int)ACT_RELU : (EPI & kEpiGelu) ? (int)ACT_GELU : (int)ACT_NONE) : p.act)
#define F_OUT32 (SP ? (EPI & kEpiOut32) != 0 : p.out_f32 != nullptr)
#define F_OUT16 (SP ? (EPI & kEpiOut16) != 0 : p.out_bf16 != nullptr)
#define F_OUT2 (SP ? (EPI & kEpiOut2) != 0 : p.out2_f32 != nullptr)
#define F_KV (SP ? ((EPI & kEpiKvK) ? 1 : (EPI & kEpiKvV) ? 2 : 0) : p.kv_mode)
    uint32_t rank = 0u;
    if constexpr (two) rank = cluster_ctarank();
    // a pipeline stage holds `ksub` K sub-blocks of bk channels: [ksub][128][bk] of A then [ksub][bn_local][bk] of B
    const uint32_t a_sub = 128u * p.bk * 2u;
    const uint32_t bn_local = two ? (uint32_t)p.bn / 2u : (uint32_t)p.bn;
    const uint32_t b_sub = bn_local * p.bk * 2u;
    // halo mode: a stage = the halo tile of one 64-channel block (rounded up to the 1024-byte swizzle atom) + the nine
    // per-tap weight tiles of that block
    const bool halo = p.halo != 0;
    const uint32_t halo_tx = (uint32_t)(p.halo_pw * p.halo_ph) * (uint32_t)p.bk * 2u;
    const uint32_t a_bytes = halo ? ((halo_tx + 1023u) & ~1023u) : a_sub * p.ksub;
    const uint32_t b_bytes = halo ? 9u * b_sub : b_sub * p.ksub;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t tx_bytes = halo ? halo_tx + b_bytes : stage_bytes;      // bytes TMA actually delivers per stage
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(smem + (size_t)num_stages * stage_bytes);
    // per-epilogue-warp transpose tile [32 rows][36 words] + row table, behind the barriers
    float* epi_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + sizeof(GemmBarriers));
    float* epi_cst = epi_base + 128 * kPerQuad;            // [2 units in flight][scale 256 | shift 256] of the unit's N tile

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars->full[s], two ? 2 : 1);         // pair: one arrive per CTA's producer, on the leader
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], (two ? 2 : 1) * 32 * kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if constexpr (two) tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
        else tmem_alloc(&bars->tmem_base, kTmemCols);
    }
    tc_fence_before();
    if constexpr (two) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // PDL: this CTA now holds its shared memory and TMEM, so the next kernel of the stream may be scheduled behind it;
    // nothing above touched global memory written by the previous kernel, everything below may
    pdl_trigger();
    pdl_wait();
    const bool tracing = p.trace != nullptr && blockIdx.x < 4;
    unsigned long long* const trc = tracing ? p.trace + (size_t)blockIdx.x * 3 * 16 * 4 : nullptr;
    if (tracing && threadIdx.x == 0) p.trace[768 + blockIdx.x * 2] = clock64();

    const int n_tiles = p.N / p.bn;
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_f;
    // work units: one (M tile, N tile) per CTA, or one (M-tile pair, N tile) per CTA pair
    const int total_tiles = two ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
    const int unit0 = two ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = two ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int ksplit = p.ksplit > 1 ? p.ksplit : 1;         // K slices per tile (split-K for tile-poor, K-long GEMMs)
    const int total_units = total_tiles * ksplit;
    const int nk = halo ? p.cin_blocks : p.taps * p.cin_blocks / p.ksub / ksplit;   // pipeline stages per work unit
    const int bh_log2 = p.bh_log2, bw_log2 = p.bw_log2;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // (whole warp convergent, one elected lane issues: keeps coordinates in uniform registers)
        int s = 0;
        uint32_t phase = 0;
        int ucount = 0;
        for (int unit = unit0; unit < total_units; unit += unit_step, ++ucount) {
            if (tracing && ucount < 16 && lane == 0) trc[(0 * 16 + ucount) * 4 + 0] = clock64();
            const int tile = unit / ksplit, ks = unit - tile * ksplit;
            const int nt = tile % n_tiles;
            const int mt = two ? 2 * (tile / n_tiles) + (int)rank : tile / n_tiles;   // may be one past the end
            const int tx = mt % p.tiles_x;
            const int ty = (mt / p.tiles_x) % p.tiles_y;
            const int tf = mt / (p.tiles_x * p.tiles_y);
            const int x0 = tx << bw_log2;
            const int y0 = ty << bh_log2;
            int f0 = tf << (7 - bw_log2 - bh_log2);
            if (p.f_group) f0 = (f0 / p.f_used) * p.f_group + f0 % p.f_used;
            const int y2 = (p.ydim == 2) ? y0 : 0;
            const int y3 = (p.ydim == 3) ? y0 : 0;
            // per-frame B operand is indexed by the SOURCE frame; a pair splits the B rows between its CTAs
            const int brow = nt * p.bn + (two ? (int)(rank * bn_local) : f0 * p.b_rows_per_frame);
            const int sub0 = ks * nk * p.ksub;                 // first K sub-block of this slice
            // K order: channel block outer, tap inner -- the order the halo path needs (one box per channel block), used by
            // both paths so that a layer's rounding does not depend on which of them a batch size selects
            int cb = sub0 / p.taps, tap = sub0 % p.taps;
            for (int kb = 0; kb < nk; ++kb) {
                mbar_wait(&bars->empty[s], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    if constexpr (!two) {
                        mbar_expect_tx(&bars->full[s], tx_bytes);
                    } else {
                        // both CTAs' bytes are credited to the leader's barrier
                        if (rank == 0) mbar_expect_tx(&bars->full[s], 2u * tx_bytes);
                        else mbar_arrive_leader(&bars->full[s]);
                    }
                    if (halo) {
                        // one box = the tile's pixels plus the dilation halo (out-of-image parts zero-filled = padding)
                        const int c0 = kb * p.bk, hx = x0 - p.halo_d, hy = y0 - p.halo_d;
                        if constexpr (!two) tma_load_5d(sa, &tmA, &bars->full[s], c0, hx, hy, 0, f0);
                        else tma_load_5d_2sm(sa, &tmA, &bars->full[s], c0, hx, hy, 0, f0);
                        for (int t = 0; t < 9; ++t) {
                            const int kcol = (t * p.cin_blocks + kb) * p.bk;
                            if constexpr (!two) tma_load_2d(sa + a_bytes + t * b_sub, &tmB, &bars->full[s], kcol, brow);
                            else tma_load_2d_2sm(sa + a_bytes + t * b_sub, &tmB, &bars->full[s], kcol, brow);
                        }
                    }
                    int tp = tap, c = cb;
                    for (int j = 0; j < (halo ? 0 : p.ksub); ++j) {
                        const int c0 = p.tap_off[tp][0] + c * p.bk;
                        const int c1 = p.tap_off[tp][1] + x0;
                        const int c2 = p.tap_off[tp][2] + y2;
                        const int c3 = p.tap_off[tp][3] + y3;
                        const int kcol = (tp * p.cin_blocks + c) * p.bk;
                        if constexpr (!two) {
                            tma_load_5d(sa + j * a_sub, &tmA, &bars->full[s], c0, c1, c2, c3, f0);
                            tma_load_2d(sa + a_bytes + j * b_sub, &tmB, &bars->full[s], kcol, brow);
                        } else {
                            tma_load_5d_2sm(sa + j * a_sub, &tmA, &bars->full[s], c0, c1, c2, c3, f0);
                            tma_load_2d_2sm(sa + a_bytes + j * b_sub, &tmB, &bars->full[s], kcol, brow);
                        }
                        if (++tp == p.taps) { tp = 0; ++c; }
                    }
                }
                __syncwarp();
                tap += p.ksub;
                while (tap >= p.taps) { tap -= p.taps; ++cb; }
                if (++s == num_stages) { s = 0; phase ^= 1u; }
            }
            if (tracing && ucount < 16 && lane == 0) trc[(0 * 16 + ucount) * 4 + 1] = clock64();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop convergently (uniform registers, no per-instruction election); one elected
        // lane issues.  Descriptors are base + constant offsets: the issue loop must stay far shorter than the MMA
        // time of a K block (192 cycles at N = 96), or the tensor pipe starves on instruction issue.
        if (rank == 0) {                                    // in a pair only the leader issues
            const uint32_t idesc = (two ? umma_idesc_bf16_m256((uint32_t)p.bn) : umma_idesc_bf16((uint32_t)p.bn)) &
                                   (p.ab_f16 ? ~kIdescAbBf16 : ~0u);
            const uint32_t row_bytes = (uint32_t)p.bk * 2u;
            const uint64_t da0 = umma_smem_desc(smem_u32(smem), row_bytes);
            const uint64_t db0 = umma_smem_desc(smem_u32(smem) + a_bytes, row_bytes);
            const uint32_t stage_step = stage_bytes >> 4;   // descriptor start-address units (16 B)
            const uint64_t a_sub_step = a_sub >> 4, b_sub_step = b_sub >> 4;
            const bool k64 = p.bk == 64;
            // halo mode: descriptor field for the 8-row group stride (= one halo row) and the per-tap start offsets
            const uint32_t row16 = row_bytes >> 4;                                  // 16-byte units per pixel row
            const uint64_t halo_sbo = (uint64_t)((uint32_t)p.halo_pw * row16) << 32;
            const uint32_t halo_step_x = (uint32_t)p.halo_d * row16, halo_step_y = halo_step_x * (uint32_t)p.halo_pw;
            int s = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int ucount = 0;
            for (int unit = unit0; unit < total_units; unit += unit_step, ++ucount) {
                mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                if (tracing && ucount < 16 && lane == 0) trc[(1 * 16 + ucount) * 4 + 0] = clock64();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&bars->full[s], phase);
                    tc_fence_after();
                    if (tracing && kb == 0 && ucount < 16 && lane == 0) trc[(1 * 16 + ucount) * 4 + 1] = clock64();
                    if (elect_one()) {
                        uint64_t da = da0 + (uint64_t)((uint32_t)s * stage_step);
                        uint64_t db = db0 + (uint64_t)((uint32_t)s * stage_step);
                        if (halo) {
                            // nine taps = nine row-shifted views of the halo tile: start + ((ty*d)*PW + tx*d) rows of 128 B,
                            // 8-row groups (one image row of the 8-pixel-wide tile) PW rows apart
                            // (fully unrolled, constants folded: this single-thread loop must stay well below the
                            // 48-cycle MMA it issues)
                            const uint64_t dah = (da & ~(0x3FFFull << 32)) | halo_sbo;
#pragma unroll
                            for (int t = 0; t < 9; ++t) {
                                const uint64_t a_t = dah + (uint64_t)((uint32_t)(t / 3) * halo_step_y + (uint32_t)(t % 3) * halo_step_x);
                                const uint64_t b_t = db + (uint64_t)((uint32_t)t * (uint32_t)b_sub_step);
                                const uint32_t first = (kb | t) ? 1u : 0u;
                                if constexpr (two) {
                                    umma_bf16_2sm(d_tmem, a_t, b_t, idesc, first);
                                    umma_bf16_2sm(d_tmem, a_t + 2, b_t + 2, idesc, 1u);
                                    if (k64) {
                                        umma_bf16_2sm(d_tmem, a_t + 4, b_t + 4, idesc, 1u);
                                        umma_bf16_2sm(d_tmem, a_t + 6, b_t + 6, idesc, 1u);
                                    }
                                } else {
                                    umma_bf16(d_tmem, a_t, b_t, idesc, first);
                                    umma_bf16(d_tmem, a_t + 2, b_t + 2, idesc, 1u);
                                    if (k64) {
                                        umma_bf16(d_tmem, a_t + 4, b_t + 4, idesc, 1u);
                                        umma_bf16(d_tmem, a_t + 6, b_t + 6, idesc, 1u);
                                    }
                                }
                            }
                        }
                        for (int j = 0; j < (halo ? 0 : p.ksub); ++j) {
                            const uint32_t first = (kb | j) ? 1u : 0u;
                            if constexpr (two) {
                                umma_bf16_2sm(d_tmem, da, db, idesc, first);
                                umma_bf16_2sm(d_tmem, da + 2, db + 2, idesc, 1u);
                                if (k64) {
                                    umma_bf16_2sm(d_tmem, da + 4, db + 4, idesc, 1u);
                                    umma_bf16_2sm(d_tmem, da + 6, db + 6, idesc, 1u);
                                }
                            } else {
                                umma_bf16(d_tmem, da, db, idesc, first);
                                umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
                                if (k64) {
                                    umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
                                    umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
                                }
                            }
                            da += a_sub_step;
                            db += b_sub_step;
                        }
                        // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
                        if constexpr (two) umma_commit_2sm(&bars->empty[s]); else umma_commit(&bars->empty[s]);
                    }
                    __syncwarp();
                    if (++s == num_stages) { s = 0; phase ^= 1u; }
                }
                if (elect_one()) {                                   // accumulator complete -> epilogue(s)
                    if constexpr (two) umma_commit_2sm(&bars->tmem_full[acc]); else umma_commit(&bars->tmem_full[acc]);
                }
                __syncwarp();
                if (tracing && ucount < 16 && lane == 0) trc[(1 * 16 + ucount) * 4 + 2] = clock64();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;             // which of the quadrant's two warps: takes every other chunk
        const int r = q * 32 + lane;                  // accumulator row == pixel within the tile
        const int rx = r & ((1 << bw_log2) - 1);
        const int ry = (r >> bw_log2) & ((1 << bh_log2) - 1);
        const int rf = r >> (bw_log2 + bh_log2);
        int acc = 0;
        uint32_t acc_phase = 0;
        int ucount = 0;
        const bool tr_e = tracing && warp == 2 && lane == 0;
        const int dbg = p.trace ? (int)p.trace[780] : 0;      // tools/gemm_trace.py experiments: 1 = no stores, 2 = no TMEM loads
        for (int unit = unit0; unit < total_units; unit += unit_step, ++ucount) {
            if (tr_e && ucount < 16) trc[(2 * 16 + ucount) * 4 + 0] = clock64();
            const int tile = unit / ksplit, ks = unit - tile * ksplit;
            const int nt = tile % n_tiles;
            const int mt = two ? 2 * (tile / n_tiles) + (int)rank : tile / n_tiles;
            const int tx = mt % p.tiles_x;
            const int ty = (mt / p.tiles_x) % p.tiles_y;
            const int tf = mt / (p.tiles_x * p.tiles_y);
            const int x = (tx << bw_log2) + rx;
            const int y = (ty << bh_log2) + ry;
            const int f = (tf << (7 - bw_log2 - bh_log2)) + rf;
            const bool valid = (x < p.W) && (y < p.H) && (f < p.F);
            const int fs = p.f_group ? (f / p.f_used) * p.f_group + f % p.f_used : f;     // source frame
            const int fo = p.out_remap ? fs : f;
            const size_t pix_in = ((size_t)fs * p.H + y) * p.W + x;
            const size_t pix_out = ((size_t)(fo * p.out_fmul + p.out_fadd) * p.H + y) * p.W + x;
            const size_t pix_out2 = ((size_t)(fo * p.out2_fmul + p.out2_fadd) * p.H + y) * p.W + x;
            const int n0 = nt * p.bn;

            // The unit's per-column constants go to shared memory while the MMAs of the unit are still running: read from
            // global inside the chunk loops, every chunk paid an exposed L2 round trip for a new cache line of them
            // (the dominant stall of the bf16 epilogues in the ncu source view).  Two buffers: a warp that runs ahead
            // into the next unit must not overwrite what a slower warp still reads.
            const uint32_t cs = smem_u32(epi_cst + (ucount & 1) * 512);   // shared-window byte address
            if (F_SCALE || F_SHIFT) {
                const int et = (int)threadIdx.x - 64;
                if (et < p.bn) {
                    sts32(cs + 4u * et, F_SCALE ? __ldg(p.scale + n0 + et) : 1.0f);
                    sts32(cs + 4u * (256 + et), F_SHIFT ? __ldg(p.shift + n0 + et) : 0.0f);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");   // the epilogue warps only
            }
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            if (tr_e && ucount < 16) trc[(2 * 16 + ucount) * 4 + 1] = clock64();
            const uint32_t t_addr = tmem_base + (uint32_t)acc * kAccStride + ((uint32_t)(q * 32) << 16);
            float head = 0.0f;
            if (!SP && p.out_softmax) {
                // 2 heads x 18 keys: the whole score row lives in this thread (one warp per quadrant does it)
                if (half == 0) {
                uint32_t raw[48];
                tmem_ld16(t_addr + 0, *reinterpret_cast<uint32_t(*)[16]>(raw + 0));
                tmem_ld16(t_addr + 16, *reinterpret_cast<uint32_t(*)[16]>(raw + 16));
                tmem_ld16(t_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(raw + 32));
                tmem_ld_wait();
                if (valid) {
                    float pr[36];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 18; ++j) {
                            float sc = __uint_as_float(raw[h * 18 + j]);
                            if (p.rowbias) sc += __ldg(p.rowbias + (size_t)f * p.N + h * 18 + j);
                            pr[h * 18 + j] = sc;
                            mx = fmaxf(mx, sc);
                        }
                        float sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < 18; ++j) { pr[h * 18 + j] = expf(pr[h * 18 + j] - mx); sum += pr[h * 18 + j]; }
                        const float inv = 1.0f / sum;
#pragma unroll
                        for (int j = 0; j < 18; ++j) pr[h * 18 + j] *= inv;
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.out_softmax + pix_out * 64);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        op[k] = make_uint4(pack_bf16x2(pr[8 * k], pr[8 * k + 1]), pack_bf16x2(pr[8 * k + 2], pr[8 * k + 3]),
                                           pack_bf16x2(pr[8 * k + 4], pr[8 * k + 5]), pack_bf16x2(pr[8 * k + 6], pr[8 * k + 7]));
                    op[4] = make_uint4(pack_bf16x2(pr[32], pr[33]), pack_bf16x2(pr[34], pr[35]), 0u, 0u);
                    op[5] = make_uint4(0u, 0u, 0u, 0u);
                    op[6] = make_uint4(0u, 0u, 0u, 0u);
                    op[7] = make_uint4(0u, 0u, 0u, 0u);
                }
                }
            } else if (!SP && p.head_w) {
                // fused 96 -> 1 head: each of the quadrant's two warps reduces its chunks, partials meet in smem
                for (int c = half * 16; c < p.bn; c += 16 * kPerQuad) {
                    uint32_t raw[16];
                    tmem_ld16(t_addr + c, raw);
                    tmem_ld_wait();
                    if (valid) {
                        const int n = n0 + c;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float v = __uint_as_float(raw[j]);
                            if (p.scale) v *= __ldg(p.scale + n + j);
                            if (p.shift) v += __ldg(p.shift + n + j);
                            if (p.act == ACT_RELU) v = fmaxf(v, 0.0f);
                            head = fmaf(v, __ldg(p.head_w + n + j), head);
                        }
                    }
                }
            } else if (SP ? !(EPI & kEpiTransposed) : !p.epi_transposed) {
                // direct epilogue: thread = one output row, 16-column chunks; the residual of a chunk is requested
                // before the TMEM load so that its DRAM latency overlaps the tcgen05.ld round trip
                for (int c = half * 16; c < p.bn; c += 16 * kPerQuad) {
                    const int n = n0 + c;
                    float4 res[4];
                    if (F_RES && valid) {
                        const float4* rp = reinterpret_cast<const float4*>(p.residual + pix_in * p.N + n);
#pragma unroll
                        for (int j = 0; j < 4; ++j) res[j] = rp[j];
                    }
                    uint32_t raw[16];
                    tmem_ld16(t_addr + c, raw);
                    tmem_ld_wait();
                    if (valid) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                        if (F_SCALE) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t = lds128f(cs + 4u * (c + 4 * j));
                                v[4 * j] *= t.x; v[4 * j + 1] *= t.y; v[4 * j + 2] *= t.z; v[4 * j + 3] *= t.w;
                            }
                        }
                        if (F_SHIFT) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t = lds128f(cs + 4u * (256 + c + 4 * j));
                                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                            }
                        }
                        if (F_ROWBIAS) {
                            const float4* rb = reinterpret_cast<const float4*>(p.rowbias + (size_t)fs * p.N + n);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t = __ldg(rb + j);
                                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                            }
                        }
                        if (F_ACT == ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
                        } else if (F_ACT == ACT_GELU) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 g = gelu_erf2(make_float2(v[2 * j], v[2 * j + 1]));
                                v[2 * j] = g.x; v[2 * j + 1] = g.y;
                            }
                        }
                        if (F_RES) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                v[4 * j] += res[j].x; v[4 * j + 1] += res[j].y; v[4 * j + 2] += res[j].z; v[4 * j + 3] += res[j].w;
                            }
                        }
                        if (F_OUT32) {
                            float4* op = reinterpret_cast<float4*>(p.out_f32 + (size_t)ks * p.split_stride + pix_out * p.ldo + n);
#pragma unroll
                            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                        if (F_OUT2) {
                            float4* op = reinterpret_cast<float4*>(p.out2_f32 + pix_out2 * p.ldo + n);
#pragma unroll
                            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                        if (F_KV) {
                            // attention operands: row = frame * 18 + key, 16 columns of one head (kv_C % 16 == 0)
                            const int fr = (int)pix_out / 18, j = (int)pix_out - fr * 18;
                            const int hh = n / p.kv_C, cc = n - hh * p.kv_C;
                            if (F_KV == 1) {
                                uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + ((size_t)fr * p.kv_R + hh * 18 + j) * p.kv_C + cc);
                                op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                   pack_bf16x2(v[6], v[7]));
                                op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                   pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                            } else {
                                // transposed: the lanes of a warp (consecutive keys) write neighbouring 2-byte elements
                                bf16* op = p.out_bf16 + ((size_t)fr * p.kv_C + cc) * 64 + hh * 18 + j;
#pragma unroll
                                for (int k = 0; k < 16; ++k) op[k * 64] = __float2bfloat16(v[k]);
                            }
                        } else if (F_OUT16) {
                            uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + pix_out * p.ldo + n);
                            if (!SP && p.out_f16) {
                                op[0] = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]),
                                                   pack_f16x2(v[6], v[7]));
                                op[1] = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]),
                                                   pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
                            } else {
                                op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                   pack_bf16x2(v[6], v[7]));
                                op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                   pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                            }
                        }
                    }
                }
            } else {
                // general epilogue: 32-column chunks are transposed through shared memory so that one warp
                // instruction touches whole 128-byte rows (coalesced residual loads and output stores)
                const uint32_t tile = smem_u32(epi_base + 128 * kPerQuad + 1024 + (warp - 2) * (32 * 36 + 128));   // [32][36] floats
                const int cq = (lane & 7) * 4;                             // this lane's 4 columns inside the chunk
                const int rq = lane >> 3;                                  // row offset inside a group of 4 rows
                // where the 8 rows this lane stores (row = it * 4 + rq) go: fetched once per unit from the lanes that own
                // those rows (shuffles; the store loop below has no load-dependent branch or address)
                int po[8];
                {
                    const int my_po = valid ? (int)pix_out : -1;
#pragma unroll
                    for (int it = 0; it < 8; ++it) po[it] = __shfl_sync(0xffffffffu, my_po, it * 4 + rq);
                }
                for (int c = half * 32; c < p.bn; c += 32 * kPerQuad) {
                    // the chunk's residual rows (coalesced: 4 rows x 128 B per instruction) and per-column constants are
                    // requested before the TMEM load so that their latency overlaps the tcgen05.ld + smem transpose
                    const int n = n0 + c + cq;
                    float4 resv[8];
                    if (F_RES) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            resv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            const int pin = __shfl_sync(0xffffffffu, (int)pix_in, it * 4 + rq);
                            if (po[it] >= 0) resv[it] = *reinterpret_cast<const float4*>(p.residual + (size_t)pin * p.N + n);
                        }
                    }
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (F_SCALE) sc = lds128f(cs + 4u * (c + cq));
                    if (F_SHIFT) sh = lds128f(cs + 4u * (256 + c + cq));
                    uint32_t raw[32];
                    if (dbg != 2) {
                        tmem_ld32(t_addr + c, raw);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) raw[j] = 0u;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        sts128(tile + 4u * (lane * 36 + 4 * j), make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]));
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rq;
                        float4 v = lds128f(tile + 4u * (row * 36 + cq));
                        if (F_SCALE) {
                            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                            v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        } else if (F_SHIFT) {
                            v.x += sh.x; v.y += sh.y; v.z += sh.z; v.w += sh.w;
                        }
                        if (F_ROWBIAS) {
                            const int fr = __shfl_sync(0xffffffffu, fs, row);
                            if (po[it] >= 0) {
                                const float4 rb = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)fr * p.N + n));
                                v.x += rb.x; v.y += rb.y; v.z += rb.z; v.w += rb.w;
                            }
                        }
                        if (F_ACT == ACT_RELU) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        } else if (F_ACT == ACT_GELU) {
                            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
                        }
                        if (F_RES) {
                            const float4 t = resv[it];
                            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                        }
                        if (F_OUT2) {
                            const int po2 = __shfl_sync(0xffffffffu, (int)pix_out2, row);
                            if (po[it] >= 0) *reinterpret_cast<float4*>(p.out2_f32 + (size_t)po2 * p.ldo + n) = v;
                        }
                        if (po[it] >= 0 && dbg != 1) {
                            if (F_OUT32) *reinterpret_cast<float4*>(p.out_f32 + (size_t)po[it] * p.ldo + n) = v;
                            if (F_OUT16)
                                *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)po[it] * p.ldo + n) =
                                    (!SP && p.out_f16) ? make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w))
                                                       : make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                        }
                    }
                }
            }
            if (!SP && p.head_w) {
                float* hp = epi_base + q * 32 + lane;                           // [4 quadrants][32 rows]
                if (half > 0) hp[(half - 1) * 128] = head;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");   // the epilogue warps only
                if (half == 0 && valid) {
#pragma unroll
                    for (int k = 1; k < kPerQuad; ++k) head += hp[(k - 1) * 128];
                    p.out_head[pix_out] = 1.0f / (1.0f + __expf(-(head + p.head_b)));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            }
            tc_fence_before();
            if (tr_e && ucount < 16) trc[(2 * 16 + ucount) * 4 + 2] = clock64();
            if constexpr (two) mbar_arrive_leader(&bars->tmem_empty[acc]); else mbar_arrive(&bars->tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    if constexpr (two) cluster_sync_all(); else __syncthreads();
    if (tracing && threadIdx.x == 0) p.trace[768 + blockIdx.x * 2 + 1] = clock64();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (two) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}

#undef F_SCALE
#undef F_SHIFT
#undef F_ROWBIAS
#undef F_RES
#undef F_ACT
#undef F_OUT32
#undef F_OUT16
#undef F_OUT2
#undef F_KV

// the flavours the denoiser's program uses (anything else runs the generic kernel)
constexpr int kET = kEpiSpecial | kEpiTransposed | kEpiOut32;      // fp32 output, row-coalesced epilogue
constexpr int kED = kEpiSpecial | kEpiOut16;                       // bf16 output, direct epilogue
#define DSB_EPI_LIST(X)                                                                                             \
    X(0)                                                                       /* generic                         */ \
    X(kET | kEpiShift)                                                         /* 1x1 shortcut, plain linears     */ \
    X(kET | kEpiShift | kEpiResidual)                                          /* P.V + proj bias, fc2            */ \
    X(kET | kEpiScale | kEpiShift | kEpiRelu)                                  /* upembed.conv2 (last stage)      */ \
    X(kET | kEpiScale | kEpiShift | kEpiRelu | kEpiResidual)                   /* upembed.conv2 + skip            */ \
    X(kET | kEpiRelu)                                                          /* ReduceTemp                      */ \
    X(kET | kEpiShift | kEpiRowbias)                                           /* res.conv1 + temb                */ \
    X(kET | kEpiShift | kEpiOut2)                                              /* Downsample, two destinations    */ \
    X(kEpiSpecial | kEpiOut32)                                                 /* split-K partial sums            */ \
    X(kED | kEpiShift | kEpiGelu)                                              /* fc1                             */ \
    X(kED | kEpiScale | kEpiShift | kEpiRelu)                                  /* upembed.conv1                   */ \
    X(kED | kEpiShift | kEpiResidual)                                          /* res.conv2 + shortcut            */ \
    X(kED | kEpiShift | kEpiKvK)                                               /* folded K projection             */ \
    X(kED | kEpiShift | kEpiKvV)                                               /* folded V projection             */

struct GemmVariant {
    int epi;
    const void* fn[2];      // [two_cta]
};
#define DSB_EPI_ROW(E) {(E), {(const void*)gemm_tc_kernel<false, (E)>, (const void*)gemm_tc_kernel<true, (E)>}},
static const GemmVariant kVariants[] = {DSB_EPI_LIST(DSB_EPI_ROW)};
static constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

// epilogue flavour of a launch (0: generic)
static int epi_of(const GemmParams& p) {
    if (p.out_softmax || p.head_w || p.out_f16) return 0;
    int e = kEpiSpecial;
    if (p.epi_transposed) e |= kEpiTransposed;
    if (p.scale) e |= kEpiScale;
    if (p.shift) e |= kEpiShift;
    if (p.rowbias) e |= kEpiRowbias;
    if (p.residual) e |= kEpiResidual;
    if (p.act == ACT_RELU) e |= kEpiRelu;
    if (p.act == ACT_GELU) e |= kEpiGelu;
    if (p.out_f32) e |= kEpiOut32;
    if (p.out_bf16) e |= kEpiOut16;
    if (p.out2_f32) e |= kEpiOut2;
    if (p.kv_mode == 1) e |= kEpiKvK;
    if (p.kv_mode == 2) e |= kEpiKvV;
    static const bool off = [] { const char* v = getenv("DSB_EPI_GENERIC"); return v && v[0] == '1'; }();
    if (off) return 0;
    for (int i = 1; i < kNumVariants; ++i)
        if (kVariants[i].epi == e) return i;
    return 0;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    const uint32_t inner_bytes = box[0] * 2u;
    CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_NONE) return -2;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(100 + (int)r);
}

int gemm_init() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < kNumVariants; ++i)
            for (int t = 0; t < 2; ++t) {
                cudaError_t e = cudaFuncSetAttribute(kVariants[i].fn[t], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return (int)e;
            }
        attr_set = true;
    }
    return 0;
}

int gemm_launch(const GemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                cudaStream_t stream) {
    if (p.bn % 16 || p.bn > 256 || p.bn < 16 || p.N % p.bn) return -10;
    if (p.bk != 64 && p.bk != 32) return -11;
    if (p.taps < 1 || p.taps > 9 || p.cin_blocks < 1) return -12;
    if (p.bw_log2 + p.bh_log2 > 7) return -13;
    if (p.head_w && p.N != p.bn) return -14;
    if (p.out_softmax && (p.N != 48 || p.bn != 48)) return -17;
    if ((p.b_rows_per_frame || p.f_group) && (p.bw_log2 + p.bh_log2 != 7)) return -18;
    if (p.bn % 32 && !p.out_softmax && !p.head_w) return -19;
    if (p.f_group && (p.f_used < 1 || p.f_used > p.f_group)) return -19;
    if (p.two_cta && (p.b_rows_per_frame || p.out_softmax || (p.bn / 2) % 8)) return -20;
    if (p.ksub < 1 || (p.taps * p.cin_blocks) % p.ksub) return -21;
    if (p.out_f16 && (p.out_softmax || p.ksplit > 1)) return -25;
    if (p.ksplit > 1) {
        if (p.two_cta || p.epi_transposed || p.head_w || p.out_softmax || p.out_bf16 || p.out2_f32 || p.scale || p.shift ||
            p.rowbias || p.residual || p.act != ACT_NONE || !p.out_f32 || p.f_group)
            return -22;
        if ((p.taps * p.cin_blocks / p.ksub) % p.ksplit) return -22;
    }
    uint32_t stage_bytes = (128u * p.bk * 2u + (uint32_t)(p.two_cta ? p.bn / 2 : p.bn) * p.bk * 2u) * p.ksub;
    if (p.halo) {
        if (p.taps != 9 || p.bw_log2 != 3 || p.bh_log2 != 4 || p.ksplit > 1 || p.ydim != 2 || p.halo_d < 1 ||
            p.halo_pw != 8 + 2 * p.halo_d || p.halo_ph != 16 + 2 * p.halo_d)
            return -23;
        stage_bytes = (((uint32_t)(p.halo_pw * p.halo_ph) * (uint32_t)p.bk * 2u + 1023u) & ~1023u) +
                      9u * (uint32_t)(p.two_cta ? p.bn / 2 : p.bn) * (uint32_t)p.bk * 2u;
    }
    const uint32_t epi_bytes = (uint32_t)((p.epi_transposed ? kEpiWarps * (32 * 36 + 128) : 0) + 128 * kPerQuad + 1024) * sizeof(float);
    const uint32_t budget = 225u * 1024u - 1024u - (uint32_t)sizeof(GemmBarriers) - epi_bytes;
    int stages = (int)(budget / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return -15;
    const size_t smem = (size_t)stages * stage_bytes + sizeof(GemmBarriers) + epi_bytes + 1024;
    if (int e = gemm_init()) return e;
    const int vi = epi_of(p);
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_f;
    if (!p.two_cta) {
        const int total = m_tiles * (p.N / p.bn) * (p.ksplit > 1 ? p.ksplit : 1);
        int grid = total < num_sms ? total : num_sms;
        if (grid < 1) return -16;
        void* args[4] = {(void*)&p, (void*)&tmA, (void*)&tmB, (void*)&stages};
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_use(true) ? 1 : 0;
        return (int)cudaLaunchKernelExC(&cfg, kVariants[vi].fn[0], args);
    }
    const int pairs = ((m_tiles + 1) / 2) * (p.N / p.bn);
    int clusters = pairs < num_sms / 2 ? pairs : num_sms / 2;
    if (clusters < 1) return -16;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_use(true) ? 2 : 1;
    void* args[4] = {(void*)&p, (void*)&tmA, (void*)&tmB, (void*)&stages};
    return (int)cudaLaunchKernelExC(&cfg, kVariants[vi].fn[1], args);
}

}  // namespace dsb
