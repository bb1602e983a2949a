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
from .engine import AudioEngine, DsbError

_SUPPORTED = dict(heads=2, dim=512, mlp_dim=256, dim_head=64, height=7, width=12)


class AudioAttnNetB200(nn.Module):
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
        self._engine = None
        self._sd = None

    def load_state_dict(self, state_dict, strict=True, prefix=""):
        want = [k for k, _ in synth.audio_attn_state_dict_spec(self.depth)]
        have = {k[len(prefix):] for k in state_dict if k.startswith(prefix)}
        used = [k for k in want if k.startswith("transformer.")]
        missing = [k for k in (want if strict else used) if k not in have]
        unexpected = [k for k in have if k not in set(want)]
        if missing or (strict and unexpected):
            raise DsbError("load_state_dict: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
        self._sd = {k: state_dict[prefix + k].detach().float().cpu().clone() for k in want if prefix + k in state_dict}
        if self._engine is not None:
            self._engine.close()
        self._engine = AudioEngine(self.max_batch)
        self._engine.load_state_dict(self._sd)
        return self

    def state_dict(self, *args, **kwargs):
        return dict(self._sd) if self._sd is not None else {}

    @property
    def engine(self):
        if self._engine is None:
            raise DsbError("AudioAttnNetB200 has no weights: call load_state_dict(reference_state_dict) first")
        return self._engine

    @torch.no_grad()
    def forward(self, audio):
        return self.engine.forward(audio)


def register_as_audio_attn_net(registry, name="AudioAttnNet", force=True):
    """Registers AudioAttnNetB200 under the reference's class name so that ``spatiotemp_net=dict(type="AudioAttnNet",
    ...)`` (cfgs/audio_visual.py:34) builds the B200 path."""
    registry.register_module(name=name, force=force, module=AudioAttnNetB200)
    return AudioAttnNetB200
