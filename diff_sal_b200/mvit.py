"""Drop-in replacement of the reference's ``MViT`` video encoder (models/mvit.py:796-1152) as DiffSal configures it:
``visual_net=dict(type=MViT, arch="small", pretrained=..., out_scales=[0, 1, 2, 3])`` (cfgs/audio_visual.py:27-32), called
once per clip as ``visual_net(imgs)`` (models/diff_model.py:103-104, diffusion_trainer.py:559-562).  SURVEY 8f row N2,
video half.

Same ``load_state_dict`` keys (401 tensors), same ``forward(x) -> [scale3, scale2, scale1, scale0]`` on
``x[B,3,16,224,384]`` (or the loader's raw 4-D view).  All arithmetic runs in libdiffsal_b200 (tcgen05 GEMMs for every
linear and both attention products, hand-written kernels for the patch embedding, the pooling attention's depthwise
convolutions, the decomposed relative-position softmax and the norms); there is no PyTorch fallback.
"""
import torch

from . import synth
from ._module import EngineModule
from .engine import DsbError, MvitEngine


class MViTB200(EngineModule):
    def __init__(self, arch="small", pretrained=None, out_scales=(0, 1, 2, 3), max_batch=2, **kwargs):
        super().__init__()
        if str(arch).lower() != "small" or sorted(out_scales) != [0, 1, 2, 3] or kwargs:
            raise DsbError("MViTB200 implements the configuration DiffSal uses: arch='small', out_scales=[0,1,2,3] "
                           "(got arch=%r, out_scales=%r, %r)" % (arch, out_scales, kwargs))
        if pretrained:
            raise DsbError("MViTB200(pretrained=...): load the checkpoint yourself (strip the 'backbone.' prefix, "
                           "models/mvit.py:1071-1104) and call load_state_dict")
        self.max_batch = int(max_batch)

    def _spec(self):
        return synth.mvit_state_dict_spec()

    def _make_engine(self):
        return MvitEngine(self.max_batch)

    @torch.no_grad()
    def forward(self, x):
        return self.engine.forward(x)


def register_as_mvit(registry, name="MViT", force=True):
    registry.register_module(name=name, force=force, module=MViTB200)
    return MViTB200
