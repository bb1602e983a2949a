"""Drop-in replacement of the reference's ``SalUNet`` decoder (models/saliency_decoder/sal_unet.py:145-328).

Same constructor kwargs (cfgs/audio_visual.py:50-82 / cfgs/visual.py:33-70), same reference-keyed
``load_state_dict`` and the same ``forward(x, t, feat_list, audio_feat_list=None) -> [B,1,H,W]`` signature, so it
can be registered under the reference's registry name and used by the unmodified ``DiffusionTrainer.sample_ddim``
(diffusion_trainer.py:458), ``VideoSaliencyModel.forward`` (models/diff_model.py:113) and the DPM-solver closure
(diffusion_trainer.py:593-595).  All arithmetic runs in libdiffsal_b200 (hand-written sm_100a kernels); there is
no PyTorch fallback.  Unlike the reference, ``feat_list`` is never mutated (sal_unet.py:317).
"""
import torch
import torch.nn as nn

from . import synth
from ._module import EngineModule
from .engine import DsbError, Engine

_SUPPORTED = dict(
    image_based=True, img_size=(224, 384), frames_len=1, mid_num_stages=4, temporal_size=9,
    temporal_list=[5, 5, 5, 5], futr_num_stages=0, ori_embed_dim=768, down_embed_dim=96,
    patch_size=[0, 3, 3, 3], patch_stride=[0, 1, 1, 1], patch_padding=[0, 2, 2, 2], up_channel=[768, 384, 192, 96],
    num_heads=[2, 2, 2, 2], mlp_ratio=[2.0, 2.0, 2.0, 2.0], qkv_bias=[True] * 4, kv_proj_method=["avg"] * 4,
    kernel_kv=[2, 4, 8, 16], padding_kv=[0] * 4, stride_kv=[2, 4, 8, 16], q_proj_method=["dw_bn"] * 4,
    kernel_q=[3] * 4, padding_q=[1] * 4, stride_q=[1] * 4)


def _norm(v):
    if isinstance(v, (list, tuple)):
        return [_norm(e) for e in v]
    if isinstance(v, float) and v == int(v):
        return int(v)
    return v


class SalUNetB200(EngineModule):
    def __init__(self, max_batch=8, audio_visual=True, **kwargs):
        super().__init__()
        # the kernels are specialised for the one decoder configuration both reference configs use
        for k, want in _SUPPORTED.items():
            if k in kwargs and _norm(kwargs[k]) != _norm(want):
                raise DsbError("SalUNetB200: %s=%r is outside the supported hot-path configuration (%r)" % (k, kwargs[k], want))
        self.kwargs = dict(kwargs)
        self.img_size = (224, 384)
        self.max_batch = int(max_batch)
        self.audio_visual = bool(audio_visual)
        self._cond_ident = None
        self._cond_hash = None
        self._cond_refs = None

    # ------------------------------------------------------------------ weights (see _module.EngineModule)
    def _spec(self):
        return synth.state_dict_spec()

    def _make_engine(self):
        return Engine(self.max_batch, self.audio_visual)

    def _weights_changed(self):
        self._cond_ident = self._cond_hash = self._cond_refs = None

    # ------------------------------------------------------------------ condition cache
    @staticmethod
    def _identity(feat_list, audio):
        """(pointer, version, shape, dtype) of the caller's tensors, or None when a tensor carries no version counter
        (inference-mode tensors): then only the content fingerprint can tell whether the conditioning changed."""
        try:
            key = tuple((f.data_ptr(), f._version, tuple(f.shape), f.dtype, f.device) for f in feat_list[:3])
            key += ((audio.data_ptr(), audio._version, tuple(audio.shape), audio.dtype, audio.device)
                    if audio is not None else None,)
            return key
        except RuntimeError:
            return None

    def _condition(self, feat_list, audio):
        """Runs dsb_set_condition only when the conditioning VALUES changed.  Fast path: the very same, unmodified tensor
        objects as last time (strong references are kept, so an equal address cannot be a recycled allocation).
        Otherwise -- e.g. the reference's ``copy.deepcopy(tmp_img)`` per step (diffusion_trainer.py:452), fp16 / CPU /
        non-contiguous inputs that are converted to fresh tensors -- a 64-bit content fingerprint decides."""
        ident = self._identity(feat_list, audio)
        if ident is not None and ident == self._cond_ident:
            return
        feats, aud = self.engine.prepare_condition(feat_list, audio)
        h = self.engine.condition_hash(feats, aud)
        if h != self._cond_hash:
            self.engine.set_condition(feats, aud, prepared=True)
            self._cond_hash = h
        self._cond_ident = ident
        self._cond_refs = (list(feat_list[:3]), audio)

    # ------------------------------------------------------------------ reference signature
    @torch.no_grad()
    def forward(self, x, t, feat_list, audio_feat_list=None):
        self._condition(feat_list, audio_feat_list)
        return self.engine.denoise(x, t)

    # ------------------------------------------------------------------ fused loop (used by diff_sal_b200.sampler)
    @torch.no_grad()
    def _dsb_fused_sample(self, ops, x, feat_list, model_kwargs=None, noise=None, use_graph=True):
        audio = (model_kwargs or {}).get("audio_feat_list")
        self._condition(feat_list, audio)
        out = x.to(device=self.engine.device, dtype=torch.float32).clone().contiguous()
        if noise is not None:
            noise = noise.to(device=self.engine.device, dtype=torch.float32).contiguous()
        return self.engine.sample(ops, out, noise=noise, use_graph=use_graph)


def register_as_salunet(registry, name="SalUNet", force=True):
    """Registers SalUNetB200 in an mmcv-style registry under the reference's class name so that
    ``decoder_net=dict(type="SalUNet", ...)`` builds the B200 path."""
    registry.register_module(name=name, force=force, module=SalUNetB200)
    return SalUNetB200
