"""GPU: dsb_metrics (CC / SIM / NSS / AUC-Judd on device tensors) against the numpy oracle restatement
of the reference's metrics (oracle/metrics.py, itself pinned to metrics/metrics.py in tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(i, shape=(224, 384)):
    from oracle import metrics as M
    rng = np.random.RandomState(100 + i)
    dens, fix = M.synthetic_ground_truth(i, shape)
    pred = (rng.rand(*shape) ** 3 + (0.2 + 0.3 * i) * dens / dens.max()).astype(np.float32)
    return pred, dens.astype(np.float32), fix.astype(np.float32)


def test_metrics_match_oracle():
    from diff_sal_b200 import metrics as G
    from oracle import metrics as M
    cases = [_case(i) for i in range(3)]
    pred = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    dens = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    fix = torch.from_numpy(np.stack([c[2] for c in cases])).cuda()
    got = {k: v.cpu().numpy() for k, v in G.saliency_metrics(pred, dens, fix).items()}
    for i, (p, d, f) in enumerate(cases):
        assert abs(got["CC"][i] - M.cc(p, d)) <= 1e-9
        assert abs(got["SIM"][i] - M.sim(p, d)) <= 1e-9
        assert abs(got["NSS"][i] - M.nss(p, f)) <= 1e-9 * max(1.0, abs(M.nss(p, f)))
        assert abs(got["AUC_J"][i] - M.auc_judd(p, f, jitter=False)) <= 1e-12


def test_auc_judd_with_reference_jitter():
    """The reference draws rand*1e-7 from numpy's global RNG before AUC-J (metrics.py:44-45); feeding the same numbers
    reproduces its value on maps with heavy ties."""
    from diff_sal_b200 import metrics as G
    from oracle import metrics as M
    p, d, f = _case(1)
    p = np.round(p * 20) / 20                                    # many exact ties
    np.random.seed(7)
    want = M.auc_judd(p, f, jitter=True)
    np.random.seed(7)
    jit = np.random.rand(*p.shape) * 1e-7
    got = G.AUC_Judd(torch.from_numpy(p).cuda(), torch.from_numpy(f).cuda(), jitter=torch.from_numpy(jit).cuda())
    assert abs(got.item() - want) <= 1e-9


def test_reference_named_functions_and_errors():
    from diff_sal_b200 import metrics as G
    from diff_sal_b200.engine import DsbError
    from oracle import metrics as M
    p, d, f = _case(2, (112, 192))
    pt, dt, ft = [torch.from_numpy(a).cuda() for a in (p, d, f)]
    assert abs(G.CC(pt, dt).item() - M.cc(p, d)) <= 1e-9
    assert abs(G.SIM(pt, dt).item() - M.sim(p, d)) <= 1e-9
    assert abs(G.NSS(pt, ft).item() - M.nss(p, f)) <= 1e-9
    assert np.isnan(G.AUC_Judd(pt, torch.zeros_like(ft)).item())       # no fixation: NaN like the reference
    with pytest.raises(DsbError):
        G.CC(pt.cpu(), dt.cpu())


def test_validation_losses_on_device():
    """dsb_val_losses (models/sal_losses.py:14-176,207-233) against the fp32 oracle restatement, batch of maps from the
    sampler's value range plus a sparse ground truth."""
    import types
    from diff_sal_b200 import sal_losses as S
    from oracle import sal_losses as O
    g = torch.Generator().manual_seed(21)
    pred = torch.rand(4, 1, 224, 384, generator=g) * 0.3 + 0.1
    gt = torch.rand(4, 1, 224, 384, generator=g) ** 6
    per = S.per_clip_losses(pred.cuda(), gt.cuda()).cpu()
    for b in range(4):
        want = [O.kldiv2(pred[b:b + 1], gt[b:b + 1]), O.cc_s2(pred[b:b + 1], gt[b:b + 1]), O.similarity2(pred[b:b + 1], gt[b:b + 1]),
                O.nss2(pred[b:b + 1], gt[b:b + 1])]
        for k in range(4):
            assert abs(per[b, k].item() - want[k].item()) <= 2e-5 * max(1.0, abs(want[k].item())), (b, k)
    for kl in (True, False):
        cfg = types.SimpleNamespace(loss=types.SimpleNamespace(loss_kl=kl))
        got, want = S.get_kl_cc_sim_loss_wo_weight(cfg, pred.cuda(), gt.cuda()), O.get_kl_cc_sim_loss_wo_weight(kl, pred, gt)
        assert set(got) == set(want)
        for k in got:
            assert abs(float(got[k]) - float(want[k])) <= 2e-5 * max(1.0, abs(float(want[k]))), k
    from diff_sal_b200.engine import DsbError
    with pytest.raises(DsbError):
        S.per_clip_losses(pred, gt)                       # CPU tensors: no fallback
