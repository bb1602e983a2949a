"""Drop-in replacement of the reference's ``VGGish`` audio feature network (models/vggish.py:106-124) for the one
method the saliency model calls: ``forward_feat`` (models/vggish.py:87-90, used by
``VideoSaliencyModel.forward_vggish``, models/diff_model.py:70-76).  SURVEY 8f row N2, audio half.

Same ``load_state_dict`` keys (``features.*``; the 113 M-parameter ``embeddings`` MLP is accepted and ignored because
``forward_feat`` never runs it), same ``forward_feat(x[(B*T),1,112,192]) -> [(B*T),512,7,12]``.  All arithmetic runs in
libdiffsal_b200 (one direct conv+pool kernel, five tcgen05 implicit-GEMM convolutions, max pools); no PyTorch fallback.
``forward`` (the AudioSet embedding head) is outside the hot path and raises.
"""
import torch
import torch.nn as nn

from . import synth
from .engine import DsbError, VggishEngine


class VGGishB200(nn.Module):
    def __init__(self, pretrained=False, max_frames=72):
        super().__init__()
        if pretrained:
            raise DsbError("VGGishB200(pretrained=True): load the checkpoint yourself and call load_state_dict "
                           "(the reference reads data/pretrained_models/vggish.pth, models/vggish.py:110-119)")
        self.max_frames = int(max_frames)
        self._engine = None
        self._sd = None

    def load_state_dict(self, state_dict, strict=True, prefix=""):
        want = [k for k, _ in synth.vggish_state_dict_spec()]
        have = {k[len(prefix):] for k in state_dict if k.startswith(prefix)}
        missing = [k for k in want if k.startswith("features.") and k not in have]
        unexpected = [k for k in have if k not in set(want)]
        if missing or (strict and unexpected):
            raise DsbError("load_state_dict: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
        self._sd = {k: state_dict[prefix + k].detach().float().cpu().clone() for k in want
                    if k.startswith("features.") and prefix + k in state_dict}
        if self._engine is not None:
            self._engine.close()
        self._engine = VggishEngine(self.max_frames)
        self._engine.load_state_dict(self._sd)
        return self

    def state_dict(self, *args, **kwargs):
        return dict(self._sd) if self._sd is not None else {}

    @property
    def engine(self):
        if self._engine is None:
            raise DsbError("VGGishB200 has no weights: call load_state_dict(reference_state_dict) first")
        return self._engine

    @torch.no_grad()
    def forward_feat(self, x):
        return self.engine.forward_feat(x)

    def forward(self, x):
        raise DsbError("VGGishB200.forward (AudioSet embedding head) is outside the DiffSal hot path; use forward_feat")
