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
from ._module import EngineModule
from .engine import DsbError, VggishEngine


class VGGishB200(EngineModule):
    def __init__(self, pretrained=False, max_frames=72):
        super().__init__()
        if pretrained:
            raise DsbError("VGGishB200(pretrained=True): load the checkpoint yourself and call load_state_dict "
                           "(the reference reads data/pretrained_models/vggish.pth, models/vggish.py:110-119)")
        self.max_frames = int(max_frames)

    # weights: see _module.EngineModule (``embeddings.*`` is accepted, kept for state_dict round trips, never uploaded)
    def _spec(self):
        return synth.vggish_state_dict_spec()

    def _required(self, key):
        return key.startswith("features.")

    def _optional(self, key):
        return key.startswith("embeddings.")            # the 113 M-parameter AudioSet head forward_feat never runs

    def _make_engine(self):
        return VggishEngine(self.max_frames)

    @torch.no_grad()
    def forward_feat(self, x):
        return self.engine.forward_feat(x)

    def forward(self, x):
        raise DsbError("VGGishB200.forward (AudioSet embedding head) is outside the DiffSal hot path; use forward_feat")
