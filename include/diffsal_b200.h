/* libdiffsal_b200 -- C ABI of the B200-native DiffSal sampling hot path.
 *
 * Drop-in boundary for the reference's denoiser + sampler update.  Plain pointers and sizes only;
 * every call returns 0 on success or a negative error class (message via dsb_last_error); nothing throws, nothing
 * falls back to the CPU.  A handle is bound to the CUDA device that is current at dsb_create, is not thread-safe,
 * and all work is enqueued asynchronously on the caller's stream.  Device tensors passed in are never modified
 * (the reference's in-place feature-list mutation, sal_unet.py:317, is not reproduced) and never freed.
 *
 * What each entry point replaces in /root/reference:
 *   dsb_create / dsb_load_weight / dsb_finalize_weights
 *        SalUNet.__init__ + load_state_dict with the reference key names
 *        (models/saliency_decoder/sal_unet.py:146-277, model.py:17-22)
 *   dsb_set_condition
 *        the per-clip, loop-invariant part of SalUNet.forward: feature list / audio features handed to the
 *        decoder (models/diff_model.py:106-113, diffusion_trainer.py:556-565) incl. TransformerBlock.align_conv
 *        on the audio features (models/saliency_decoder/transformer.py:128-131)
 *   dsb_denoise
 *        SalUNet.forward(x, t, feat_list, audio_feat_list)  (sal_unet.py:302-328) = one denoiser evaluation
 *   dsb_sampler_update
 *        the elementwise state updates of DiffusionTrainer.sample_ddim (diffusion_trainer.py:434-437,470-478)
 *        and DPM_Solver.{dpm_solver_first_update, multistep_dpm_solver_second/third_update, data_prediction_fn}
 *        (models/dpm_solver/sampler.py:548-593,797-905,434-443), with host-computed scalar coefficients
 *   dsb_sample
 *        the whole loop: DiffusionTrainer.sample_ddim (diffusion_trainer.py:439-480) /
 *        DPM_Solver.sample(method="multistep") (sampler.py:1048-1247) as a program of EVAL / AXPY ops
 *   dsb_audio_create / dsb_audio_load_weight / dsb_audio_finalize / dsb_audio_forward
 *        AudioAttnNet.__init__ + load_state_dict + forward (models/audio_attention.py:93-143), the once-per-clip
 *        audio transformer whose output VideoSaliencyModel.forward_vggish hands to the decoder
 *        (models/diff_model.py:70-81,97-113) -- SURVEY.md 8f row N1
 *   dsb_vggish_create / dsb_vggish_load_weight / dsb_vggish_finalize / dsb_vggish_forward_feat
 *        VGGish.__init__ + load_state_dict + forward_feat (models/vggish.py:87-124) -- SURVEY.md 8f row N2 (audio half)
 *   dsb_mvit_create / dsb_mvit_load_weight / dsb_mvit_finalize / dsb_mvit_forward
 *        MViT.__init__ + load_state_dict + forward (models/mvit.py:796-1152) -- SURVEY.md 8f row N2 (video half)
 *   dsb_metrics
 *        metrics.metrics.{CC, SIM, NSS, AUC_Judd} (metrics/metrics.py:7-64,178-252) -- SURVEY.md 8f row N3
 *   dsb_val_losses
 *        models.sal_losses.{kldiv2, cc_s2, similarity2, nss2} behind get_kl_cc_sim_loss_wo_weight
 *        (models/sal_losses.py:14-176,207-233) -- SURVEY.md 8f row N3
 */
#ifndef DIFFSAL_B200_H
#define DIFFSAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsb_handle dsb_handle;

typedef struct dsb_config {
    int max_batch;      /* clips per denoiser evaluation the workspace is sized for            */
    int audio_visual;   /* 1: cfgs/audio_visual.py (audio-gated K), 0: cfgs/visual.py           */
    int reserved[6];
} dsb_config;

/* error classes */
#define DSB_OK 0
#define DSB_ERR_ARG (-1)         /* bad argument / shape / state             */
#define DSB_ERR_UNSUPPORTED (-2) /* configuration outside the hot path       */
#define DSB_ERR_CUDA (-3)        /* CUDA runtime / driver error              */
#define DSB_ERR_WEIGHT (-4)      /* missing / mis-shaped weight              */

int dsb_create(const dsb_config* cfg, dsb_handle** out);
void dsb_destroy(dsb_handle* h);
const char* dsb_last_error(const dsb_handle* h);
/* bytes of device memory the handle owns for a batch of B clips (workspace + repacked weights) */
size_t dsb_workspace_bytes(const dsb_handle* h, int B);

/* fp32 tensor under its reference state_dict key; `data` may be a host or a device pointer. */
int dsb_load_weight(dsb_handle* h, const char* ref_key, const void* data, const int64_t* shape, int ndim);
/* BN folding, bf16 repack to K-major [N][tap*Cin+c], conv_in o down1 composition, depthwise tap tables. */
int dsb_finalize_weights(dsb_handle* h);

/* feat[i]: device fp32 [B, C_i, 8, h_i, w_i] (C = 768,384,192,96; feat[3] is accepted and ignored exactly like
 * the reference); audio: device fp32 [B,512,9,7,12] or NULL for the visual-only configuration. */
int dsb_set_condition(dsb_handle* h, const void* const feat[4], const void* audio_or_null, int B, void* stream);

/* 64-bit fingerprint of the values dsb_set_condition would read (same arguments).  The reference's sample_ddim hands
 * the decoder a fresh deep copy of the feature list on every step (diffusion_trainer.py:452): the host side compares
 * fingerprints instead of pointers to skip the (loop-invariant) conditioning.  Synchronises `stream`. */
int dsb_condition_hash(dsb_handle* h, const void* const feat[4], const void* audio_or_null, int B, uint64_t* out,
                       void* stream);

/* x: device fp32 [B,1,224,384]; t: device fp32 [B] (model time, may be fractional); out: device fp32 [B,1,224,384] */
int dsb_denoise(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream);

/* out = sum_k coef[k] * in[k] (nin <= 4) + noise_coef * noise; n elements (multiple of 4), device fp32. */
int dsb_sampler_update(dsb_handle* h, const float* coef, const float* const* in, int nin, const float* noise_or_null,
                       float noise_coef, float* out, int64_t n, void* stream);

/* ---- whole sampling loop as a small program over device buffers -------------------------------------------
 * Buffer ids: 0 = x (the state, in/out), 1 = raw network output of the last EVAL, 2..7 = scratch (model history).
 * DSB_OP_EVAL : buf[1] = SalUNet(buf[src[0]], t)                 (t = op.t for every clip; src[0] = 0 unless the solver
 *               evaluates the network at an intermediate point, as the singlestep DPM-Solver updates do)
 * DSB_OP_AXPY : buf[dst] = sum_k coef[k] * buf[src[k]] + noise_coef * noise[noise_index]
 */
#define DSB_OP_EVAL 0
#define DSB_OP_AXPY 1
#define DSB_OP_CLAMP 2     /* buf[dst] = clamp(buf[dst], coef[0], coef[1])            (util/denoising.py:54)         */
#define DSB_OP_DYNTHRESH 3 /* DPM_Solver.dynamic_thresholding_fn (sampler.py:417-426) on buf[dst], per clip:
                            * s = max(lerp(|x|_(k), |x|_(k+1), coef[0]), coef[1]) with k = noise_index the floor of
                            * torch.quantile's fp32 rank p * (n - 1); buf = clamp(buf, -s, s) / s                     */
#define DSB_OP_POSTPROCESS 4 /* output side of the loop fused into the program: buf[dst] =
                            * clamp(buf[dst], 0, 1) = inverse_data_transform (datasets/__init__.py:26-35) and
                            * desc.out_u8 = normalize_data(buf[dst]) (util/utils.py:11-16, per-map min-max -> uint8)      */
typedef struct dsb_sampler_op {
    int kind;
    float t;
    int dst;
    int nin;
    int src[4];
    float coef[4];
    float noise_coef;
    int noise_index;    /* -1: none; else index of a [B,1,224,384] slab in `noise` */
} dsb_sampler_op;

typedef struct dsb_sampler_desc {
    const dsb_sampler_op* ops;
    int n_ops;
    const float* noise; /* device fp32 [n_slabs][B,1,224,384] or NULL (copied to handle-owned staging: a cached graph
                         * stays valid when the caller passes a fresh randn tensor on every call)                       */
    int use_graph;      /* 1: capture the program into a CUDA graph (cached per program, LRU of 8) and replay it */
    uint8_t* out_u8;    /* device uint8 [B][224*384], written by DSB_OP_POSTPROCESS; NULL if the program has none     */
} dsb_sampler_desc;

int dsb_sample(dsb_handle* h, const dsb_sampler_desc* desc, float* x_inout, int B, void* stream);

/* the two corrector ops as stand-alone calls (no handle state), for loops driven around an arbitrary denoiser:
 * in-place clamp (util/denoising.py:54 `torch.clamp(x0_from_e, -1, 1)`) and DPM_Solver.dynamic_thresholding_fn
 * (models/dpm_solver/sampler.py:417-426; x: [B][n], k / w = floor / frac of torch.quantile's fp32 rank p*(n-1)) */
int dsb_sampler_clamp(float* x, int64_t n, float lo, float hi, void* stream);
int dsb_sampler_dynamic_threshold(float* x, int B, int64_t n, int k, float w, float max_val, void* stream);
/* error norm of DPM_Solver.dpm_solver_adaptive's step-size control (models/dpm_solver/sampler.py:996-999), per sample:
 * out[b] = sqrt(mean(((x_higher - x_lower) / max(atol, rtol * max(|x_lower|, |x_prev|)))^2)); all device fp32 [B][n] */
int dsb_sampler_adaptive_error(const float* x_lower, const float* x_higher, const float* x_prev, int B, int64_t n, float atol,
                               float rtol, float* out, void* stream);

/* output side of the loop: clamp(x,0,1) = inverse_data_transform (datasets/__init__.py:26-35, cfgs/diffusion.yml data.*)
 * and the per-map min-max -> uint8 of normalize_data (util/utils.py:11-16).  x: device fp32 [B][pixels_per_map];
 * either output may be NULL. */
int dsb_postprocess(const float* x, int B, int64_t pixels_per_map, float* clamped_or_null, uint8_t* u8_or_null,
                    void* stream);

/* saliency metrics on the device: CC, SIM (against a density map), NSS, AUC-Judd (against a binary
 * fixation map, at most 1024 fixations per clip) of metrics/metrics.py:7-64,178-252 with metrics/utils.py:11-52
 * normalisation, fp64 accumulation.  pred / density / fixations: device fp32 [B][pixels_per_map]; jitter: device fp64
 * [B][pixels_per_map] holding the reference's `rand * 1e-7` AUC-J jitter (metrics.py:44-45) or NULL;
 * out4: device fp64 [B][4] = CC, SIM, NSS, AUC_J. */
int dsb_metrics(const float* pred, const float* density, const float* fixations, const double* jitter_or_null, int B,
                int64_t pixels_per_map, double* out4, void* stream);

/* validation losses of the reference's sampling loop on the device: kldiv2 / cc_s2 / similarity2 /
 * nss2 of models/sal_losses.py:14-176 (get_kl_cc_sim_loss_wo_weight :207-233, diffusion_trainer.py:741,797,868), PER CLIP:
 * pred / gt: device fp32 [B][elems_per_clip]; out4: device fp64 [B][4] = kl, cc, sim, nss (the reference returns their
 * batch means). */
int dsb_val_losses(const float* pred, const float* gt, int B, int64_t elems_per_clip, double* out4, void* stream);

/* number of kernel launches enqueued by the last dsb_denoise / dsb_sample call (for bench.py's gpu_launches) */
int64_t dsb_last_launch_count(const dsb_handle* h);

/* process-wide mode of programmatic dependent launch for the kernels of a program: 0 off, 1 the memory-bound kernels
 * only (default), 2 also the tcgen05 kernels (DSB_PDL in the environment sets the initial value).  Changing it does not
 * invalidate graphs that were already captured. */
void dsb_set_pdl(int mode);
int dsb_get_pdl(void);

/* kernel launches enqueued by the last dsb_set_condition call */
int64_t dsb_condition_launch_count(const dsb_handle* h);

/* ---- measurement: one denoiser evaluation with a CUDA-event pair around every launch (bench.py's roofline) ---
 * ms[i] = device time of launch i, flops[i] = its algorithmic FLOPs (2*M*N*K as torch's flop counter counts the
 * reference graph; 0 for memory-bound kernels), bytes[i] = algorithmic bytes where stated.  Returns #launches. */
int dsb_profile_denoise(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream, float* ms,
                        double* flops, double* bytes, int cap);
/* the same evaluation as it really runs (side streams, event joins): start / end of every launch in ms since the first,
 * and the stream it ran on (0 = caller's).  Names through dsb_profile_name.  Returns #launches. */
int dsb_profile_timeline(dsb_handle* h, const float* x, const float* t, float* out, int B, void* stream, float* start_ms,
                         float* end_ms, int* stream_idx, int cap);
const char* dsb_profile_name(const dsb_handle* h, int i);

/* ---- debugging / parity taps (read-only views of the workspace after dsb_denoise) ------------------------- */
/* copies an internal fp32 buffer to `dst` (device); names: "noise0".."noise2" ([B,hw,C] frame-8 slices),
 * "x0".."x3" (stage outputs [B*9,hw,C]), "r0".."r3" ([B,hw,768]), "p" ([B,112,192]).  Returns element count. */
int64_t dsb_debug_read(dsb_handle* h, const char* name, float* dst, int64_t max_elems, void* stream);

/* ---- once-per-clip audio transformer (models/audio_attention.py:93-143; cfgs/audio_visual.py:34-48) -------------
 * Own opaque handle (weights + workspace for up to max_batch clips).  Weight keys are AudioAttnNet.state_dict()'s
 * ("transformer.layers.0.0.to_qkv.weight", ...); to_patch_embedding.* / pos_embedding may be loaded but are unused,
 * exactly as in the reference, whose forward discards the patch embedding (audio_attention.py:134-141).  Only the
 * shipped geometry is accepted: dim 512, 2 heads x 64, mlp_dim 256, 9 x 7 x 12 tokens; depth is taken from the keys.
 * audio / out: [B, 512, 9, 7, 12] fp32 device tensors (out may not alias audio). */
typedef struct dsb_audio dsb_audio;
int dsb_audio_create(int max_batch, dsb_audio** out);
void dsb_audio_destroy(dsb_audio* h);
const char* dsb_audio_last_error(const dsb_audio* h);
int dsb_audio_load_weight(dsb_audio* h, const char* ref_key, const void* data, const int64_t* shape, int ndim);
int dsb_audio_finalize(dsb_audio* h);
int dsb_audio_forward(dsb_audio* h, const float* audio, float* out, int B, void* stream);
int dsb_audio_last_launch_count(const dsb_audio* h);

/* ---- VGGish feature stack (models/vggish.py:87-103 `forward_feat`, called once per clip by
 * VideoSaliencyModel.forward_vggish, models/diff_model.py:70-76) -- SURVEY.md 8f row N2, audio half --------------
 * Weight keys are VGGish.state_dict()'s ("features.0.weight" ...); "embeddings.*" keys are accepted and ignored
 * (forward_feat never runs that MLP).  audio: device fp32 [frames][1][112][192] (= audio.view(-1, 1, 112, 192));
 * out: device fp32 [frames][512][7][12]. */
typedef struct dsb_vggish dsb_vggish;
int dsb_vggish_create(int max_frames, dsb_vggish** out);
void dsb_vggish_destroy(dsb_vggish* h);
const char* dsb_vggish_last_error(const dsb_vggish* h);
int dsb_vggish_load_weight(dsb_vggish* h, const char* ref_key, const void* data, const int64_t* shape, int ndim);
int dsb_vggish_finalize(dsb_vggish* h);
int dsb_vggish_forward_feat(dsb_vggish* h, const float* audio, float* out, int frames, void* stream);
int dsb_vggish_last_launch_count(const dsb_vggish* h);

/* ---- MViTv2-S video encoder (models/mvit.py:796-1152 `MViT(arch="small", out_scales=[0,1,2,3])`, called once per clip as
 * `visual_net(imgs)` by VideoSaliencyModel.forward / DiffusionTrainer.sample_image, models/diff_model.py:103-104,
 * diffusion_trainer.py:559-562) -- SURVEY.md 8f row N2, video half --------------------------------------------------------
 * Weight keys are MViT.state_dict()'s ("blocks.3.attn.rel_pos_h", ...; 401 tensors).  video: device fp32
 * [B][3][16][224][384] (the only geometry accepted); out[0..3]: device fp32 feature tensors in the order the reference
 * returns them: [B,768,8,7,12], [B,384,8,14,24], [B,192,8,28,48], [B,96,8,56,96]. */
typedef struct dsb_mvit dsb_mvit;
int dsb_mvit_create(int max_batch, dsb_mvit** out);
void dsb_mvit_destroy(dsb_mvit* h);
const char* dsb_mvit_last_error(const dsb_mvit* h);
int dsb_mvit_load_weight(dsb_mvit* h, const char* ref_key, const void* data, const int64_t* shape, int ndim);
int dsb_mvit_finalize(dsb_mvit* h);
int dsb_mvit_forward(dsb_mvit* h, const float* video, float* const out[4], int B, void* stream);
int dsb_mvit_last_launch_count(const dsb_mvit* h);

#ifdef __cplusplus
}
#endif
#endif
