"""TEST / BASELINE INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference.

In the build container the reference is imported from /root/reference to (a) validate the
restatement in oracle/ against the real reference and (b) generate the golden fixtures under
tests/golden/.  On the GPU box /root/reference does not exist; there the same unmodified
modules are imported from oracle/_ref/, the sourceless-bytecode build of the reference that
oracle/build_ref.py produces (git-ignored, travels with gpurun).  Nothing in the product path
imports this module: only tests/, bench.py's reference arm / cpu_baseline leg do.

The reference needs mmcv / timm / skimage / soundfile / resampy which are absent in
this image; tiny stand-ins are injected in sys.modules (no arithmetic lives in them,
except timm's trunc_normal_ / DropPath which are init-only / train-only).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("DIFFSAL_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models", "saliency_decoder")):
            return cand
    return os.environ.get("DIFFSAL_REFERENCE", "/root/reference")


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "saliency_decoder"))


class _BytecodeFinder:
    """Imports modules from the sourceless build oracle/_ref/ (files ``<module>.refbc`` = py_compile output)."""

    def __init__(self, root):
        self.root = root

    def _locate(self, fullname):
        base = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(os.path.join(base, "__init__.refbc")):
            return os.path.join(base, "__init__.refbc"), True
        if os.path.isfile(base + ".refbc"):
            return base + ".refbc", False
        if os.path.isdir(base) and any(f.endswith(".refbc") for f in os.listdir(base)):
            return None, True                             # namespace-style directory without __init__
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        path_, is_pkg = self._locate(fullname)
        if path_ is None and not is_pkg:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, self, origin=path_ or os.path.join(self.root, *fullname.split(".")),
                                              is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.join(self.root, *fullname.split("."))]
        spec.has_location = path_ is not None
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        path_, _ = self._locate(module.__name__)
        if path_ is None:
            return
        module.__file__ = path_
        with open(path_, "rb") as fh:
            code = marshal.loads(fh.read()[16:])          # 16-byte pyc header (magic, flags, mtime, size)
        exec(code, module.__dict__)


def is_source_tree():
    """True when the reference is imported from its sources, False for the bytecode build in oracle/_ref/."""
    return available() and os.path.exists(os.path.join(REF_ROOT, "models", "saliency_decoder", "sal_unet.py"))


def _install_stubs():
    import torch
    import torch.nn as nn

    if "mmcv" not in sys.modules:
        mmcv = types.ModuleType("mmcv")
        mmcv_utils = types.ModuleType("mmcv.utils")
        mmcv_reg = types.ModuleType("mmcv.utils.registry")

        class Registry:
            def __init__(self, name):
                self.name = name
                self._mods = {}

            def register_module(self, name=None, force=False, module=None):
                # mmcv semantics: usable as ``@register_module()`` or ``register_module(name=..., module=cls)``
                def _reg(cls):
                    key = name or cls.__name__
                    if key in self._mods and not force and self._mods[key] is not cls:
                        raise KeyError("%s is already registered in %s" % (key, self.name))
                    self._mods[key] = cls
                    return cls
                if module is not None:
                    return _reg(module)
                return _reg

            def build(self, cfg):
                cfg = dict(cfg)
                t = cfg.pop("type")
                cls = self._mods[t] if isinstance(t, str) else t
                return cls(**cfg)

        def build_from_cfg(cfg, registry, default_args=None):
            return registry.build(cfg)

        class Config(dict):
            @staticmethod
            def fromfile(path):
                raise NotImplementedError("stub")

        def get_logger(name, log_file=None, log_level=None):
            import logging
            return logging.getLogger(name)

        mmcv_utils.Registry = Registry
        mmcv_utils.get_logger = get_logger
        mmcv_reg.build_from_cfg = build_from_cfg
        mmcv_utils.registry = mmcv_reg
        mmcv.utils = mmcv_utils
        mmcv.Config = Config
        sys.modules["mmcv"] = mmcv
        sys.modules["mmcv.utils"] = mmcv_utils
        sys.modules["mmcv.utils.registry"] = mmcv_reg

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if not self.training or self.drop_prob == 0.0:
                    return x
                keep = 1 - self.drop_prob
                shape = (x.shape[0],) + (1,) * (x.ndim - 1)
                mask = x.new_empty(shape).bernoulli_(keep)
                return x * mask / keep

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        timm_layers.DropPath = DropPath
        timm_layers.trunc_normal_ = trunc_normal_
        timm_layers.to_2tuple = to_2tuple
        timm_models.layers = timm_layers
        timm.models = timm_models
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = timm_models
        sys.modules["timm.models.layers"] = timm_layers

    for name in ("soundfile", "resampy"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    if "skimage" not in sys.modules:
        skimage = types.ModuleType("skimage")
        sk_t = types.ModuleType("skimage.transform")
        sk_e = types.ModuleType("skimage.exposure")

        def resize(*a, **k):
            raise NotImplementedError("skimage stub: feed same-shape ndarrays")

        def img_as_float(x):
            import numpy as np
            return np.asarray(x, dtype=np.float64)

        sk_t.resize = resize
        skimage.transform = sk_t
        skimage.exposure = sk_e
        skimage.img_as_float = img_as_float
        sys.modules["skimage"] = skimage
        sys.modules["skimage.transform"] = sk_t
        sys.modules["skimage.exposure"] = sk_e


_loaded = {}


def load():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if is_source_tree():
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
    elif not any(isinstance(f, _BytecodeFinder) for f in sys.meta_path):
        ver = open(os.path.join(REF_ROOT, "PYTHON_VERSION")).read().strip() if os.path.exists(os.path.join(REF_ROOT, "PYTHON_VERSION")) else None
        if ver and ver != "%d.%d" % sys.version_info[:2]:
            raise RuntimeError("oracle/_ref was built for Python %s, this is %d.%d" % ((ver,) + tuple(sys.version_info[:2])))
        sys.meta_path.insert(0, _BytecodeFinder(REF_ROOT))     # ahead of site-packages, like sys.path[0] for the sources
    from models.saliency_decoder.sal_unet import SalUNet
    from models.dpm_solver.sampler import NoiseScheduleVP, model_wrapper, DPM_Solver
    from models.diffusion_decoder.diffusion_utils import get_beta_schedule, to_torch
    import metrics.metrics as ref_metrics

    ns = types.SimpleNamespace(
        SalUNet=SalUNet, NoiseScheduleVP=NoiseScheduleVP, model_wrapper=model_wrapper,
        DPM_Solver=DPM_Solver, get_beta_schedule=get_beta_schedule, to_torch=to_torch,
        metrics=ref_metrics)
    _loaded["ns"] = ns
    return ns


def decoder_kwargs():
    """decoder_net kwargs of cfgs/audio_visual.py:50-82 (identical in cfgs/visual.py:33-70)."""
    return dict(
        image_based=True, img_size=(224, 384), frames_len=1, tasks=["futr"],
        in_index=[0, 1, 2, 3], idx_to_planes={0: 96, 1: 192, 2: 384, 3: 768},
        mid_num_stages=4, temporal_size=9, temporal_list=[5, 5, 5, 5], keep_max_len=5,
        exclude_layers=[], futr_num_stages=0, ori_embed_dim=768, down_embed_dim=96,
        patch_size=[0, 3, 3, 3], patch_stride=[0, 1, 1, 1], patch_padding=[0, 2, 2, 2],
        up_channel=[768, 384, 192, 96], num_heads=[2, 2, 2, 2], mlp_ratio=[2.0] * 4,
        drop_path_rate=[0.15] * 4, qkv_bias=[True] * 4, kv_proj_method=["avg"] * 4,
        kernel_kv=[2, 4, 8, 16], padding_kv=[0] * 4, stride_kv=[2, 4, 8, 16],
        q_proj_method=["dw_bn"] * 4, kernel_q=[3] * 4, padding_q=[1] * 4, stride_q=[1] * 4)


def build_salunet():
    ns = load()
    m = ns.SalUNet(**decoder_kwargs())
    m.eval()
    return m
