"""The model container around the hot path: ``VideoSaliencyModel`` (models/diff_model.py:8-114) for the B200 modules.

The reference builds its sub-networks from an mmcv config through ``OBJECT_REGISTRY.build`` (diff_model.py:19-58) and
its sampling driver reaches them as ``model.module.forward_vggish`` / ``.visual_net`` / ``.decoder_net``
(diffusion_trainer.py:556-565,458).  Two ways to put the B200 path behind that interface:

* keep the reference's own ``VideoSaliencyModel`` and swap the classes its config names --
  ``register_b200_modules(OBJECT_REGISTRY)`` registers ``SalUNetB200`` / ``AudioAttnNetB200`` / ``VGGishB200`` under the
  reference's names (or import them in the cfg file in place of the reference classes); checkpoints load through the
  parent's ``load_state_dict`` (see _module.EngineModule);  tests/test_container_gpu.py does exactly this with the
  unmodified reference classes;
* or use ``VideoSaliencyModelB200`` below, the same container without the reference tree: same constructor arguments,
  same ``forward_vggish(audio)`` / ``forward(data, t)``.

``visual_net=dict(type="MViT", arch="small", out_scales=[0, 1, 2, 3])`` builds the B200 video encoder (``MViTB200``); with
``visual_net=None`` the container draws random placeholder features exactly like the reference does (diff_model.py:105-111).
"""
import torch
import torch.nn as nn

from .audio_attention import AudioAttnNetB200
from .mvit import MViTB200
from .salunet import SalUNetB200
from .vggish import VGGishB200


class Registry:
    """Minimal stand-in for ``mmcv.utils.Registry`` (util/registry.py:1-4): ``register_module`` as decorator or call,
    ``build(cfg)`` with ``cfg['type']`` a registered name or a class."""

    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force and self._module_dict[key] is not cls:
                raise KeyError("%s is already registered in %s" % (key, self.name))
            self._module_dict[key] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        t = cfg.pop("type")
        cls = self._module_dict[t] if isinstance(t, str) else t
        return cls(**cfg)


OBJECT_REGISTRY = Registry("object")


def register_b200_modules(registry, force=True):
    """Registers the drop-in modules under the reference's class names (cfgs/audio_visual.py:27-82)."""
    registry.register_module(name="SalUNet", force=force, module=SalUNetB200)
    registry.register_module(name="AudioAttnNet", force=force, module=AudioAttnNetB200)
    registry.register_module(name="VGGish", force=force, module=VGGishB200)
    registry.register_module(name="MViT", force=force, module=MViTB200)
    return registry


register_b200_modules(OBJECT_REGISTRY)


def forward_vggish(audio_net, spatiotemp_net, audio):
    """models/diff_model.py:70-81: audio [B,1,T,112,192] -> VGGish feature stack per frame -> 'b c t h w' -> the audio
    transformer.  Returns the (identical) pair the reference returns."""
    bs, T = audio.shape[0], audio.shape[2]
    frames = audio.reshape(-1, audio.shape[1], audio.shape[3], audio.shape[4])      # audio.view(-1, C, H, W)
    with torch.no_grad():
        feat = audio_net.forward_feat(frames)                                       # [(b t), 512, 7, 12]
    feat = feat.reshape(bs, T, feat.shape[1], feat.shape[2], feat.shape[3]).permute(0, 2, 1, 3, 4).contiguous()
    if spatiotemp_net is not None:
        feat = spatiotemp_net(feat)
    return feat, feat


@OBJECT_REGISTRY.register_module(name="VideoSaliencyModel")
class VideoSaliencyModelB200(nn.Module):
    """Same constructor / attributes / methods as the reference ``VideoSaliencyModel`` (diff_model.py:8-114)."""

    def __init__(self, channel_list, visual_net=None, spatiotemp_net=None, audio_net=None, decoder_net=None,
                 registry=None):
        super().__init__()
        reg = registry or OBJECT_REGISTRY

        def build(cfg):
            if cfg is None or isinstance(cfg, nn.Module):
                return cfg
            return reg.build(cfg)

        self.visual_net = build(visual_net)
        self.spatiotemp_net = build(spatiotemp_net)
        self.audio_net = build(audio_net)
        if self.audio_net is not None:
            # the reference also owns this (unused on the sampling path) 128 -> 512 -> 768 head (diff_model.py:40-46);
            # kept so that reference checkpoints load without unexpected keys
            self.fc = nn.Sequential(nn.Linear(128, 512), nn.ReLU(inplace=True), nn.Linear(512, 768))
        self.decoder_net = build(decoder_net)
        if channel_list is not None:
            self.channel_list = channel_list

    def forward_vggish(self, audio):
        return forward_vggish(self.audio_net, self.spatiotemp_net, audio)

    @torch.no_grad()
    def forward(self, data, t):
        imgs = data.get("img", None)
        x = data["input"]
        if self.audio_net is not None:
            audio_feat, audio_feat_embed = self.forward_vggish(data.get("audio", None))
        else:
            audio_feat, audio_feat_embed = None, None
        if self.visual_net is not None and imgs is not None:
            vis_list = self.visual_net(imgs)
        else:
            B, dev = x.shape[0], x.device
            vis_list = [torch.randn((B, 768, 8, 7, 12), device=dev), torch.randn((B, 384, 8, 14, 24), device=dev),
                        torch.randn((B, 192, 8, 28, 48), device=dev), torch.randn((B, 96, 8, 56, 96), device=dev)]
        return self.decoder_net(x, t, vis_list, audio_feat_embed)
