"""ORACLE (test infrastructure): numpy restatement of the reference's offline saliency
metrics, used as the secondary parity instrument (CC / NSS / SIM / AUC-J within 0.5 %).

  AUC_Judd  /root/reference/metrics/metrics.py:7-64   (jitter from the GLOBAL numpy RNG,
            rand*1e-7 added before range-normalisation -> seed numpy before each call)
  NSS       metrics.py:178-199       CC  metrics.py:202-224       SIM  metrics.py:227-252
  normalize /root/reference/metrics/utils.py:11-52 (axis=None branch)

Inputs must be same-shape float ndarrays (the reference's skimage resize branch is out of
scope).  Pinned against the imported reference in tests/test_oracle_vs_reference.py.
"""
import numpy as np


def _norm(x, method):
    x = np.asarray(x)
    if method == "standard":
        return (x - np.mean(x)) / np.std(x)
    if method == "range":
        return (x - np.min(x)) / (np.max(x) - np.min(x))
    if method == "sum":
        return x / float(np.sum(x))
    raise ValueError(method)


def auc_judd(sal, fix, jitter=True):
    sal = np.array(sal, dtype=np.float64)          # private copy: the reference jitters in place
    fix = np.asarray(fix) > 0.5
    if not np.any(fix):
        return np.nan
    if jitter:
        sal = sal + np.random.rand(*sal.shape) * 1e-7
    sal = _norm(sal, "range")
    S = sal.ravel()
    F = fix.ravel()
    s_fix = S[F]
    n_fix = len(s_fix)
    n_pix = len(S)
    thr = np.sort(s_fix)[::-1]
    # count of S >= thr[k] for every threshold in one pass (the reference loops)
    s_sorted = np.sort(S)
    above = n_pix - np.searchsorted(s_sorted, thr, side="left")
    tp = np.zeros(n_fix + 2)
    fp = np.zeros(n_fix + 2)
    tp[-1] = 1
    fp[-1] = 1
    k = np.arange(n_fix)
    tp[1:-1] = (k + 1) / float(n_fix)
    fp[1:-1] = (above - k - 1) / float(n_pix - n_fix)
    return float(np.trapezoid(tp, fp))


def nss(sal, fix):
    s = _norm(np.asarray(sal, dtype=np.float64), "standard")
    return float(np.mean(s[np.asarray(fix) > 0.5]))


def cc(a, b):
    a = _norm(np.asarray(a, dtype=np.float64), "standard")
    b = _norm(np.asarray(b, dtype=np.float64), "standard")
    return float(np.corrcoef(a.ravel(), b.ravel())[0, 1])


def sim(a, b):
    a = _norm(_norm(np.asarray(a, dtype=np.float64), "range"), "sum")
    b = _norm(_norm(np.asarray(b, dtype=np.float64), "range"), "sum")
    return float(np.sum(np.minimum(a, b)))


def synthetic_ground_truth(index, hw=(224, 384), n_fix=30, seed=777):
    """Seeded GT for clip ``index``: density = sum of 3 Gaussians (sigma 20 px), fixation
    map = n_fix pixels drawn from that density."""
    rng = np.random.RandomState(seed + index)
    H, W = hw
    yy, xx = np.mgrid[0:H, 0:W]
    dens = np.zeros(hw, dtype=np.float64)
    for _ in range(3):
        cy, cx = rng.uniform(0.2 * H, 0.8 * H), rng.uniform(0.2 * W, 0.8 * W)
        dens += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 20.0 ** 2))
    p = dens.ravel() / dens.sum()
    idx = rng.choice(H * W, size=n_fix, replace=False, p=p)
    fix = np.zeros(H * W, dtype=np.float64)
    fix[idx] = 1.0
    return dens, fix.reshape(hw)


def ground_truth_from_map(ref_map, index, n_fix=30, seed=777):
    """GT correlated with a reference prediction, so that the metrics are well away from zero and a relative
    tolerance is meaningful even with random-init weights: density = minmax(ref)^2 plus 5 % seeded noise,
    fixations = n_fix pixels drawn with probability ~ minmax(ref)^4."""
    rng = np.random.RandomState(seed + index)
    r = _norm(np.asarray(ref_map, dtype=np.float64), "range")
    dens = r ** 2 + 0.05 * rng.rand(*r.shape)
    p = (r ** 4).ravel()
    idx = rng.choice(r.size, size=n_fix, replace=False, p=p / p.sum())
    fix = np.zeros(r.size, dtype=np.float64)
    fix[idx] = 1.0
    return dens, fix.reshape(r.shape)


def all_metrics(pred, index, seed=0, gt=None):
    """CC/SIM against the density, NSS/AUC-J against the fixations, numpy RNG pinned."""
    dens, fix = synthetic_ground_truth(index, pred.shape) if gt is None else gt
    np.random.seed(seed)
    return {"CC": cc(pred, dens), "SIM": sim(pred, dens), "NSS": nss(pred, fix), "AUC_J": auc_judd(pred, fix)}
