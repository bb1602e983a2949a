"""ORACLE (test infrastructure, not product code): the reference's validation losses, restated.

kldiv2 / cc_s2 / similarity2 / nss2 and get_kl_cc_sim_loss_wo_weight of /root/reference/models/sal_losses.py:14-176,207-233
as plain torch (fp32, like the reference).  Pinned bit-exactly against the imported reference in
tests/test_oracle_vs_reference.py.
"""
import torch

EPS = 2.2204e-16


def _flat(x):
    return x.reshape(x.shape[0], -1)


def nss2(s_map, gt):
    """sal_losses.py:14-35."""
    s, g = _flat(s_map), _flat(gt)
    s = (s - s.mean(1, keepdim=True)) / (s.std(1, keepdim=True) + EPS)
    return torch.mean((s * g).sum(1) / g.sum(1))


def cc_s2(s_map, gt):
    """sal_losses.py:65-98."""
    s, g = _flat(s_map), _flat(gt)
    s = (s - s.mean(1, keepdim=True)) / s.std(1, keepdim=True)
    g = (g - g.mean(1, keepdim=True)) / g.std(1, keepdim=True)
    ab, aa, bb = (s * g).sum(1), (s * s).sum(1), (g * g).sum(1)
    return torch.mean(ab / torch.sqrt(aa * bb))


def kldiv2(s_map, gt):
    """sal_losses.py:101-127."""
    s, g = _flat(s_map), _flat(gt)
    s = s / (s.sum(1, keepdim=True) * 1.0)
    g = g / (g.sum(1, keepdim=True) * 1.0)
    eps = torch.tensor(EPS)
    return torch.mean(torch.sum(g * torch.log(eps + g / (s + eps)), 1))


def similarity2(s_map, gt):
    """sal_losses.py:130-176."""
    def norm(x):
        x = _flat(x)
        lo, hi = x.min(1, keepdim=True)[0], x.max(1, keepdim=True)[0]
        return (x - lo) / (hi - lo * 1.0)
    s, g = norm(s_map), norm(gt)
    s = s / (s.sum(1, keepdim=True) * 1.0)
    g = g / (g.sum(1, keepdim=True) * 1.0)
    return torch.mean(torch.sum(torch.min(s, g), 1))


def get_kl_cc_sim_loss_wo_weight(loss_kl, pred_map, gt):
    """sal_losses.py:207-233."""
    kl = kldiv2(pred_map, gt) if loss_kl else torch.tensor(0.0)
    cc, sim, nss = cc_s2(pred_map, gt), similarity2(pred_map, gt), nss2(pred_map, gt)
    return {"total": nss + cc + sim, "main": kl, "cc": cc, "sim": sim, "nss": nss}
