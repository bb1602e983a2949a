"""Drop-in replacement of the reference's ``AudioAttnNet`` (models/audio_attention.py:93-143), the once-per-clip audio
transformer that ``VideoSaliencyModel.forward_vggish`` runs on the VGGish feature map before the decoder is conditioned
on it (models/diff_model.py:70-81,97-113).  SURVEY 8f row N1.

Same constructor kwargs as the reference class (cfgs/audio_visual.py:34-48), same ``load_state_dict`` keys and the same
``forward(audio[B,512,9,7,12]) -> [B,512,9,7,12]``.  All arithmetic runs in libdiffsal_b200 (tcgen05 GEMMs + fused
LayerNorm / softmax kernels); there is no PyTorch fallback.  As in the reference, ``to_patch_embedding`` and
``pos_embedding`` are accepted by ``load_state_dict`` but do not influence the output (audio_attention.py:134-141
discards the embedded tokens).
"""
import torch
import torch.nn as nn

from . import synth
from ._module import EngineModule
from .engine import AudioEngine, DsbError

_SUPPORTED = dict(heads=2, dim=512, mlp_dim=256, dim_head=64, height=7, width=12)


class AudioAttnNetB200(EngineModule):
    def __init__(self, depth, heads, mlp_dim, dim=512, patch_dim=768, num_patches=16, height=7, width=7, pool="cls",
                 dim_head=64, dropout=0.0, emb_dropout=0.0, max_batch=8):
        super().__init__()
        assert pool in {"cls", "mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        got = dict(heads=heads, dim=dim, mlp_dim=mlp_dim, dim_head=dim_head, height=height, width=width)
        for k, want in _SUPPORTED.items():
            if got[k] != want:
                raise DsbError("AudioAttnNetB200: %s=%r is outside the supported hot-path configuration (%r)" % (k, got[k], want))
        if depth < 1:
            raise DsbError("AudioAttnNetB200: depth must be >= 1")
        self.depth = int(depth)
        self.num_patches = num_patches
        self.max_batch = int(max_batch)

    # weights: see _module.EngineModule (to_patch_embedding.* / pos_embedding never influence the output, exactly as in the
    # reference, but are kept for state_dict round trips)
    def _spec(self):
        return synth.audio_attn_state_dict_spec(self.depth)

    def _required(self, key):
        return key.startswith("transformer.")

    def _make_engine(self):
        return AudioEngine(self.max_batch)

    @torch.no_grad()
    def forward(self, audio):
        return self.engine.forward(audio)


def register_as_audio_attn_net(registry, name="AudioAttnNet", force=True):
    """Registers AudioAttnNetB200 under the reference's class name so that ``spatiotemp_net=dict(type="AudioAttnNet",
    ...)`` (cfgs/audio_visual.py:34) builds the B200 path."""
    registry.register_module(name=name, force=force, module=AudioAttnNetB200)
    return AudioAttnNetB200
