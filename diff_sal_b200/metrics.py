"""Saliency metrics on the GPU: the reference's offline CC / SIM / NSS / AUC-Judd
(metrics/metrics.py:7-64,178-252, normalisation of metrics/utils.py:11-52) computed by ``dsb_metrics`` on device tensors,
so predicted maps can be scored without a host round trip.  Function names and argument order follow the reference's
``metrics.metrics`` module.  No CPU fallback: CPU tensors raise.
"""
import ctypes

import torch

from . import _lib
from .engine import DsbError, _bind, _stream


def saliency_metrics(pred, density, fixations, jitter=None):
    """pred / density / fixations: CUDA tensors [B, ...] of equal shape -> dict of fp64 CUDA tensors [B] with keys
    CC, SIM (vs ``density``), NSS, AUC_J (vs ``fixations`` > 0.5).  ``jitter`` (fp64, same shape) is the reference's
    ``np.random.rand(...) * 1e-7`` AUC-J tie breaker (metrics.py:44-45); None = no jitter."""
    if not (pred.is_cuda and density.is_cuda and fixations.is_cuda):
        raise DsbError("diff_sal_b200.metrics runs on the GPU only")
    if pred.shape != density.shape or pred.shape != fixations.shape:
        raise DsbError("pred / density / fixations must have the same shape")
    lib = _bind(_lib.lib())
    lib.dsb_metrics.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.dsb_metrics.restype = ctypes.c_int
    B = pred.shape[0]
    n = pred[0].numel()
    p = pred.to(torch.float32).contiguous()
    g = density.to(device=p.device, dtype=torch.float32).contiguous()
    f = fixations.to(device=p.device, dtype=torch.float32).contiguous()
    j = None if jitter is None else jitter.to(device=p.device, dtype=torch.float64).contiguous()
    out = torch.empty((B, 4), dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib.dsb_metrics(_lib.ptr(p), _lib.ptr(g), _lib.ptr(f), _lib.ptr(j), B, n, _lib.ptr(out), _stream())
    if rc != 0:
        raise DsbError("dsb_metrics failed (%d)" % rc)
    return {"CC": out[:, 0], "SIM": out[:, 1], "NSS": out[:, 2], "AUC_J": out[:, 3]}


def _one(x):
    return x if x.dim() > 2 else x.unsqueeze(0)


def CC(saliency_map1, saliency_map2):
    """metrics/metrics.py:202-224."""
    a, b = _one(saliency_map1), _one(saliency_map2)
    return saliency_metrics(a, b, torch.zeros_like(a))["CC"]


def SIM(saliency_map1, saliency_map2):
    """metrics/metrics.py:227-252."""
    a, b = _one(saliency_map1), _one(saliency_map2)
    return saliency_metrics(a, b, torch.zeros_like(a))["SIM"]


def NSS(saliency_map, fixation_map):
    """metrics/metrics.py:178-199."""
    a, f = _one(saliency_map), _one(fixation_map)
    return saliency_metrics(a, a, f)["NSS"]


def AUC_Judd(saliency_map, fixation_map, jitter=None):
    """metrics/metrics.py:7-64; ``jitter``: the rand*1e-7 array the reference draws from numpy's global RNG, or None."""
    a, f = _one(saliency_map), _one(fixation_map)
    return saliency_metrics(a, a, f, jitter=None if jitter is None else _one(jitter))["AUC_J"]
