"""Validation losses of the reference's sampling loop on the GPU.

``DiffusionTrainer.test`` / ``test_av_data`` score every sampled batch with
``get_kl_cc_sim_loss_wo_weight(config, pred_map, gt)`` (models/sal_losses.py:207-233; call sites
diffusion_trainer.py:741,797,868), i.e. ``kldiv2`` / ``cc_s2`` / ``similarity2`` / ``nss2`` (:14-176) as eager torch
reductions.  Here one kernel launch (``dsb_val_losses``) computes the four per-clip values for the whole batch; function
names, arguments and the returned dict follow the reference module.  CUDA tensors only -- there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .engine import DsbError, _bind, _stream


def per_clip_losses(s_map, gt):
    """s_map, gt: CUDA tensors [B, ...] of equal shape -> fp64 CUDA tensor [B, 4] = kl, cc, sim, nss of every clip."""
    if not (s_map.is_cuda and gt.is_cuda):
        raise DsbError("diff_sal_b200.sal_losses runs on the GPU only")
    assert s_map.size() == gt.size()
    lib = _bind(_lib.lib())
    lib.dsb_val_losses.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.dsb_val_losses.restype = ctypes.c_int
    B = s_map.shape[0]
    p = s_map.detach().to(torch.float32).contiguous()
    g = gt.detach().to(device=p.device, dtype=torch.float32).contiguous()
    out = torch.empty((B, 4), dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib.dsb_val_losses(_lib.ptr(p), _lib.ptr(g), B, p[0].numel(), _lib.ptr(out), _stream())
    if rc != 0:
        raise DsbError("dsb_val_losses failed (%d)" % rc)
    return out


def kldiv2(s_map, gt):
    """models/sal_losses.py:101-127."""
    return per_clip_losses(s_map, gt)[:, 0].mean().float()


def cc_s2(s_map, gt):
    """models/sal_losses.py:65-98."""
    return per_clip_losses(s_map, gt)[:, 1].mean().float()


def similarity2(s_map, gt):
    """models/sal_losses.py:150-176."""
    return per_clip_losses(s_map, gt)[:, 2].mean().float()


def nss2(s_map, gt):
    """models/sal_losses.py:14-35."""
    return per_clip_losses(s_map, gt)[:, 3].mean().float()


def get_kl_cc_sim_loss_wo_weight(config, pred_map, gt):
    """models/sal_losses.py:207-233: {"total": nss + cc + sim, "main": kl (0 unless config.loss.loss_kl), "cc", "sim", "nss"}."""
    m = per_clip_losses(pred_map, gt).mean(dim=0).float()
    kl = m[0] if config.loss.loss_kl else torch.tensor(0.0)
    return {"total": m[3] + m[1] + m[2], "main": kl, "cc": m[1], "sim": m[2], "nss": m[3]}
