"""Host side of the sampling loop: schedules, step coefficients and the reference-compatible sampler API.

All per-step scalar math (the reference does ~30 tiny device ops per step for it, sampler.py:114-167,1255-1294) runs
here on the host in float64 from the reference's fp32 tables; every tensor update is one launch of the fused axpy
kernel (dsb_sampler_update), or the whole loop is compiled to a program of EVAL / AXPY ops and run by dsb_sample
(optionally as one CUDA graph).

API mirrored from /root/reference (same names, argument meaning, error behaviour):
  get_beta_schedule                         models/diffusion_decoder/diffusion_utils.py:5-45
  NoiseScheduleVP / model_wrapper / DPM_Solver.sample(method="multistep")
                                            models/dpm_solver/sampler.py:6-167,170-334,336-1247
  DiffusionSampler.sample_ddim / sample_image   diffusion_trainer.py:439-480,545-640
  compute_alpha / generalized_steps         util/denoising.py:3-36
"""
import math

import numpy as np
import torch

# ============================================================================================ schedules


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """float64 numpy betas, diffusion_utils.py:5-45."""
    n = num_diffusion_timesteps
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        betas = np.linspace(-6, 6, n)
        betas = 1 / (np.exp(-betas) + 1) * (beta_end - beta_start) + beta_start
    elif beta_schedule == "cosine":
        step = n + 1
        s = 0.008
        x = np.linspace(0, step, step)
        ac = np.cos(((x / step) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def to_torch(a):
    return torch.tensor(a, dtype=torch.float32)


def _as_fp32_cpu(betas):
    if isinstance(betas, torch.Tensor):
        return betas.detach().to(device="cpu", dtype=torch.float32)
    return torch.tensor(np.asarray(betas), dtype=torch.float32)


class DdimTables:
    """diffusion_trainer.py:47-76: fp32 cumprod of fp32 (1 - beta), kept as host float64 scalars."""

    def __init__(self, betas):
        b = _as_fp32_cpu(betas)
        ah = (1.0 - b).cumprod(dim=0)
        self.num_timesteps = b.shape[0]
        self.alphas_hat = ah.double().numpy()
        self.sqrt_alphas_hat = torch.sqrt(ah).double().numpy()
        self.sqrt_recip_alphas_hat = torch.sqrt(1.0 / ah).double().numpy()
        self.sqrt_recipm1_alphas_hat = torch.sqrt(1.0 / ah - 1).double().numpy()
        # posterior q(x_{t-1} | x_t, x_0) terms exactly as written at diffusion_trainer.py:55,66-74 (fp32 torch math)
        prev = torch.cat([torch.ones(1), ah[:-1]], dim=0)
        pv = b * (1.0 - prev) / (1.0 - ah)
        self.posterior_log_variance_clipped = torch.log(torch.maximum(pv, torch.tensor(1e-20))).double().numpy()
        self.posterior_mean_coef1 = (b * torch.sqrt(ah) / (1.0 - ah)).double().numpy()
        self.posterior_mean_coef2 = ((1.0 - prev) * torch.sqrt(1.0 - b) / (1.0 - ah)).double().numpy()


def _scalar(t):
    if isinstance(t, torch.Tensor):
        return float(t.reshape(-1)[0].item())
    if isinstance(t, np.ndarray):
        return float(t.reshape(-1)[0])
    return float(t)


class NoiseScheduleVP:
    """Forward-SDE wrapper with the reference's discrete-time semantics (sampler.py:6-167).

    The log-alpha table is built exactly like the reference (fp32 cumsum, clip where lambda < -5.1, which
    truncates the cosine schedule to total_N = 996); evaluation is piecewise-linear interpolation with the
    outermost segments extended (interpolate_fn, sampler.py:1255-1294), done on the host in float64.
    Methods accept python floats or 1-element tensors and return python floats.
    """

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20.0, dtype=torch.float32):
        if schedule not in ["discrete", "linear"]:
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'discrete' or 'linear'".format(schedule))
        self.schedule = schedule
        self.T = 1.0
        if schedule == "discrete":
            if betas is not None:
                log_alphas = 0.5 * torch.log(1 - _as_fp32_cpu(betas)).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(_as_fp32_cpu(alphas_cumprod))
            log_alphas = self.numerical_clip_alpha(log_alphas)
            self.total_N = int(log_alphas.shape[0])
            self.log_alpha_array = log_alphas.to(dtype).double().numpy()
            self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].to(dtype).double().numpy()
        else:
            self.total_N = 1000
            self.beta_0 = continuous_beta_0
            self.beta_1 = continuous_beta_1

    @staticmethod
    def numerical_clip_alpha(log_alphas, clipped_lambda=-5.1):
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(clipped_lambda)).item())
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        return log_alphas

    @staticmethod
    def _interp(x, xp, yp):
        k = len(xp)
        idx = int(np.searchsorted(xp, x, side="left"))
        i0 = 0 if idx == 0 else (k - 2 if idx == k else idx - 1)
        return yp[i0] + (x - xp[i0]) * (yp[i0 + 1] - yp[i0]) / (xp[i0 + 1] - xp[i0])

    def marginal_log_mean_coeff(self, t):
        t = _scalar(t)
        if self.schedule == "discrete":
            return float(self._interp(t, self.t_array, self.log_alpha_array))
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def marginal_alpha(self, t):
        return math.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return math.sqrt(1.0 - math.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        la = self.marginal_log_mean_coeff(t)
        return la - 0.5 * math.log(1.0 - math.exp(2.0 * la))

    def inverse_lambda(self, lamb):
        lamb = _scalar(lamb)
        if self.schedule == "linear":
            tmp = 2.0 * (self.beta_1 - self.beta_0) * np.logaddexp(-2.0 * lamb, 0.0)
            delta = self.beta_0 ** 2 + tmp
            return float(tmp / (math.sqrt(delta) + self.beta_0) / (self.beta_1 - self.beta_0))
        log_alpha = -0.5 * float(np.logaddexp(0.0, -2.0 * lamb))
        return float(self._interp(log_alpha, self.log_alpha_array[::-1], self.t_array[::-1]))

    def model_time(self, t_continuous):
        """model_wrapper.get_model_input_time (sampler.py:271-280): note the literal 1000."""
        if self.schedule == "discrete":
            return (_scalar(t_continuous) - 1.0 / self.total_N) * 1000.0
        return _scalar(t_continuous)


# ============================================================================================ elementwise updates


def _axpy(coefs, tensors, noise=None, noise_coef=0.0, out=None):
    """out = sum_k coefs[k] * tensors[k] + noise_coef * noise through the fused CUDA kernel (no CPU path)."""
    import ctypes
    from . import _lib
    from .engine import _bind, _stream
    t0 = tensors[0]
    if not t0.is_cuda:
        raise RuntimeError("diff_sal_b200 sampler updates run on the GPU only (got a %s tensor)" % t0.device)
    lib = _bind(_lib.lib())
    ts = [t.to(dtype=torch.float32).contiguous() for t in tensors]
    if out is None:
        out = torch.empty_like(ts[0])
    n = ts[0].numel()
    if n % 4:
        raise RuntimeError("sampler update needs a multiple of 4 elements")
    while len(ts) > 4:          # fold long sums pairwise (not reached by the reference's solvers)
        raise RuntimeError("at most 4 terms per update")
    cs = (ctypes.c_float * len(ts))(*[float(c) for c in coefs])
    ps = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    with torch.cuda.device(t0.device):
        rc = lib.dsb_sampler_update(None, cs, ps, len(ts), _lib.ptr(noise), float(noise_coef), _lib.ptr(out), n, _stream())
    if rc != 0:
        raise RuntimeError("dsb_sampler_update failed (%d)" % rc)
    return out


# ============================================================================================ model_wrapper


class _WrappedModel:
    """Callable returned by model_wrapper; carries what DPM_Solver needs to compile the loop."""

    def __init__(self, model, noise_schedule, model_type, model_kwargs):
        self.model = model
        self.noise_schedule = noise_schedule
        self.model_type = model_type
        self.model_kwargs = model_kwargs

    def raw(self, x, t_continuous, img=None):
        ns = self.noise_schedule
        t_in = torch.full((x.shape[0],), ns.model_time(t_continuous), dtype=torch.float32, device=x.device)
        return self.model(x, t_in, img, **self.model_kwargs)

    def mix(self, t_continuous):
        """noise prediction = cx * x + cr * raw_output  (sampler.py:287-298)."""
        ns = self.noise_schedule
        if self.model_type == "noise":
            return 0.0, 1.0
        a, s = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
        if self.model_type == "x_start":
            return 1.0 / s, -a / s
        if self.model_type == "v":
            return s, a
        if self.model_type == "score":
            return 0.0, -s
        raise ValueError(self.model_type)

    def __call__(self, x, t_continuous, img=None):
        out = self.raw(x, t_continuous, img)
        cx, cr = self.mix(t_continuous)
        if cx == 0.0 and cr == 1.0:
            return out
        return _axpy([cx, cr], [x, out])


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond", condition=None,
                  unconditional_condition=None, guidance_scale=1.0, classifier_fn=None, classifier_kwargs={}):
    """sampler.py:170-334.  ``model(x, t_input, img, **model_kwargs)`` is the denoiser; only the unconditional
    guidance path is reachable from DiffSal (diffusion_trainer.py:598-608); the classifier variants fail loudly."""
    assert model_type in ["noise", "x_start", "v", "score"]
    assert guidance_type in ["uncond", "classifier", "classifier-free"]
    if guidance_type != "uncond":
        raise NotImplementedError("guidance_type=%r is outside the DiffSal hot path (uncond only)" % guidance_type)
    return _WrappedModel(model, noise_schedule, model_type, dict(model_kwargs))


# ============================================================================================ step coefficients


def dpm_time_steps(ns, skip_type, t_T, t_0, N):
    """DPM_Solver.get_time_steps (sampler.py:454-481)."""
    if skip_type == "logSNR":
        lam_T, lam_0 = ns.marginal_lambda(t_T), ns.marginal_lambda(t_0)
        # the reference round-trips both ends through fp32 (.cpu().item() of fp32 tensors) and linspaces in fp32
        ls = torch.linspace(float(np.float32(lam_T)), float(np.float32(lam_0)), N + 1).double().numpy()
        return [ns.inverse_lambda(l) for l in ls]
    if skip_type == "time_uniform":
        return [float(v) for v in torch.linspace(t_T, t_0, N + 1).double().numpy()]
    if skip_type == "time_quadratic":
        return [float(v) for v in torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2).double().numpy()]
    raise ValueError("Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))


def multistep_update_coefs(ns, algorithm_type, t_prev, t, order, solver_type="dpmsolver"):
    """x_t = c_x * x + sum_k c_m[k] * m_k, with m_0 the newest model value (t_prev[-1]).
    Expands dpm_solver_first_update / multistep second / third update (sampler.py:548-593,797-905)."""
    pp = algorithm_type == "dpmsolver++"
    lam_t, lam_0 = ns.marginal_lambda(t), ns.marginal_lambda(t_prev[-1])
    h = lam_t - lam_0
    la_0, la_t = ns.marginal_log_mean_coeff(t_prev[-1]), ns.marginal_log_mean_coeff(t)
    sig_0, sig_t = ns.marginal_std(t_prev[-1]), ns.marginal_std(t)
    alpha_t = math.exp(la_t)
    phi_1 = math.expm1(-h) if pp else math.expm1(h)
    c_x = (sig_t / sig_0) if pp else math.exp(la_t - la_0)
    b1 = (alpha_t * phi_1) if pp else (sig_t * phi_1)          # x_t = c_x x - b1 m0 ...
    if order == 1:
        return c_x, [-b1]
    if order == 2:
        r0 = (lam_0 - ns.marginal_lambda(t_prev[-2])) / h
        if solver_type == "dpmsolver":
            d = -0.5 * b1                                      # ... - 0.5 b1 D1
        elif solver_type == "taylor":
            d = alpha_t * (phi_1 / h + 1.0) if pp else -sig_t * (phi_1 / h - 1.0)
        else:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        # D1 = (m0 - m1) / r0
        return c_x, [-b1 + d / r0, -d / r0]
    if order == 3:
        lam_1, lam_2 = ns.marginal_lambda(t_prev[-2]), ns.marginal_lambda(t_prev[-3])
        r0, r1 = (lam_0 - lam_1) / h, (lam_1 - lam_2) / h
        if pp:
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            e1, e2 = alpha_t * phi_2, -alpha_t * phi_3
        else:
            phi_2 = phi_1 / h - 1.0
            phi_3 = phi_2 / h - 0.5
            e1, e2 = -sig_t * phi_2, -sig_t * phi_3
        # D1_0 = (m0-m1)/r0, D1_1 = (m1-m2)/r1, D1 = D1_0 + r0/(r0+r1) (D1_0-D1_1), D2 = (D1_0-D1_1)/(r0+r1)
        g = r0 / (r0 + r1)
        a0 = e1 * (1.0 + g) + e2 / (r0 + r1)                   # coefficient of D1_0
        a1 = -e1 * g - e2 / (r0 + r1)                          # coefficient of D1_1
        return c_x, [-b1 + a0 / r0, -a0 / r0 + a1 / r1, -a1 / r1]
    raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))


# program buffer ids (include/diffsal_b200.h): 0 = x, 1 = raw network output, 2.. = model-value history
_X, _RAW, _HIST = 0, 1, 2


def quantile_rank(ratio, n):
    """torch.quantile's 'linear' rule: rank = fp32(ratio) * (n - 1) evaluated in fp32 -> (floor(rank), frac(rank))."""
    rank = np.float32(np.float32(ratio) * np.float32(n - 1))
    k = int(np.floor(rank))
    return k, float(np.float32(rank - np.float32(k)))


def build_dpm_program(ns, steps, order=2, algorithm_type="dpmsolver", model_type="x_start", skip_type="logSNR",
                      lower_order_final=False, denoise_to_zero=True, solver_type="dpmsolver", t_start=None, t_end=None,
                      thresholding=None, map_elems=224 * 384):
    """DPM_Solver.sample(method='multistep') (sampler.py:1171-1247) as EVAL/AXPY ops; returns (ops, model_times).
    ``thresholding=(ratio, max_val)`` appends DPM_Solver.dynamic_thresholding_fn (sampler.py:417-426) to every data
    prediction, as ``correcting_x0_fn="dynamic_thresholding"`` does (sampler.py:410-411,441-442)."""
    assert steps >= order
    t_0 = 1.0 / ns.total_N if t_end is None else t_end
    t_T = ns.T if t_start is None else t_start
    assert t_0 > 0 and t_T > 0
    pp = algorithm_type == "dpmsolver++"
    wm = _WrappedModel(None, ns, model_type, {})
    ts = dpm_time_steps(ns, skip_type, t_T, t_0, steps)
    ops, times = [], []

    def model_value(t, slot):
        """m = model_fn(x, t) into history slot; m = cx*x + cr*raw."""
        times.append(ns.model_time(t))
        ops.append(("eval", ns.model_time(t)))
        cx, cr = wm.mix(t)                                   # noise prediction
        if pp:                                               # data prediction x0 = (x - sigma*noise)/alpha
            a, s = ns.marginal_alpha(t), ns.marginal_std(t)
            cx, cr = (1.0 - s * cx) / a, -s * cr / a
            if model_type == "x_start":
                cx, cr = 0.0, 1.0                            # the two conversions cancel exactly
        terms = [(_RAW, cr)] if cx == 0.0 else [(_X, cx), (_RAW, cr)]
        ops.append(("axpy", slot, terms, 0.0, -1))
        if pp and thresholding is not None:
            ops.append(("thresh", slot) + quantile_rank(thresholding[0], map_elems) + (float(thresholding[1]),))

    hist = []                                                # [(t, slot)], oldest first
    free = [_HIST + k for k in range(order)]

    def push(t):
        slot = free.pop(0) if free else hist.pop(0)[1]
        if len(hist) >= order:
            hist.pop(0)
        model_value(t, slot)
        hist.append((t, slot))

    def update(t, k):
        tp = [h_[0] for h_ in hist]
        c_x, c_m = multistep_update_coefs(ns, algorithm_type, tp, t, k, solver_type)
        terms = [(_X, c_x)] + [(hist[-1 - j][1], c_m[j]) for j in range(k)]
        ops.append(("axpy", _X, terms, 0.0, -1))

    push(ts[0])
    for step in range(1, order):
        update(ts[step], step)
        push(ts[step])
    for step in range(order, steps + 1):
        k = min(order, steps + 1 - step) if (lower_order_final and steps < 10) else order
        update(ts[step], k)
        if step < steps:
            if len(hist) == order:
                free.append(hist.pop(0)[1])
            push(ts[step])
    if denoise_to_zero:
        times.append(ns.model_time(t_0))
        ops.append(("eval", ns.model_time(t_0)))
        cx, cr = wm.mix(t_0)
        a, s = ns.marginal_alpha(t_0), ns.marginal_std(t_0)
        cx, cr = (1.0 - s * cx) / a, -s * cr / a
        if model_type == "x_start":
            cx, cr = 0.0, 1.0
        terms = [(_RAW, cr)] if cx == 0.0 else [(_X, cx), (_RAW, cr)]
        ops.append(("axpy", _X, terms, 0.0, -1))
        if thresholding is not None:                         # denoise_to_zero_fn = data_prediction_fn (sampler.py:542-546)
            ops.append(("thresh", _X) + quantile_rank(thresholding[0], map_elems) + (float(thresholding[1]),))
    return ops, times


def singlestep_orders(steps, order):
    """get_orders_and_timesteps_for_singlestep_solver (sampler.py:514-536): (orders, K)."""
    if order == 3:
        K = steps // 3 + 1
        if steps % 3 == 0:
            return [3] * (K - 2) + [2, 1], K
        if steps % 3 == 1:
            return [3] * (K - 1) + [1], K
        return [3] * (K - 1) + [2], K
    if order == 2:
        if steps % 2 == 0:
            return [2] * (steps // 2), steps // 2
        return [2] * (steps // 2) + [1], steps // 2 + 1
    if order == 1:
        return [1] * steps, 1
    raise ValueError("'order' must be '1' or '2' or '3'.")


class _SinglestepBuilder:
    """Emits dpm_solver_first_update / singlestep_dpm_solver_second_update / _third_update (sampler.py:548-795) as
    EVAL/AXPY ops.  Buffers: x = 0, model_s / model_s1 / model_s2 = 2 / 3 / 4, the intermediate iterate = 5.  Unlike the
    reference (which calls ``self.model_fn(x, s)`` without ``img`` there, sampler.py:573,585,630,...), every evaluation runs
    the conditioned network: SURVEY 8f row N4."""
    MS, MS1, MS2, XI = 2, 3, 4, 5

    def __init__(self, ns, algorithm_type, model_type, solver_type, thresholding=None, map_elems=224 * 384):
        if solver_type not in ("dpmsolver", "taylor"):
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        self.ns, self.pp, self.model_type, self.st = ns, algorithm_type == "dpmsolver++", model_type, solver_type
        self.wm = _WrappedModel(None, ns, model_type, {})
        self.thr, self.map_elems = thresholding, map_elems
        self.ops, self.times = [], []

    def model_value(self, t, slot, xbuf):
        """buf[slot] = model_fn(buf[xbuf], t): noise prediction (dpmsolver) or data prediction (dpmsolver++)."""
        ns = self.ns
        self.times.append(ns.model_time(t))
        self.ops.append(("eval", ns.model_time(t), xbuf))
        cx, cr = self.wm.mix(t)
        if self.pp:
            a, s_ = ns.marginal_alpha(t), ns.marginal_std(t)
            cx, cr = (1.0 - s_ * cx) / a, -s_ * cr / a
            if self.model_type == "x_start":
                cx, cr = 0.0, 1.0
        terms = [(_RAW, cr)] if cx == 0.0 else [(xbuf, cx), (_RAW, cr)]
        self.ops.append(("axpy", slot, terms, 0.0, -1))
        if self.pp and self.thr is not None:
            self.ops.append(("thresh", slot) + quantile_rank(self.thr[0], self.map_elems) + (float(self.thr[1]),))

    def _base(self, s, t):
        ns = self.ns
        h = ns.marginal_lambda(t) - ns.marginal_lambda(s)
        return ns, h

    def _lin(self, s, u, h_frac_phi):
        """(coefficient of x, coefficient of model_s) of the first-order part from s to u."""
        ns = self.ns
        if self.pp:
            return ns.marginal_std(u) / ns.marginal_std(s), -ns.marginal_alpha(u) * math.expm1(-h_frac_phi)
        return math.exp(ns.marginal_log_mean_coeff(u) - ns.marginal_log_mean_coeff(s)), -ns.marginal_std(u) * math.expm1(h_frac_phi)

    def first(self, s, t, have_model_s=False, dst=_X):
        ns, h = self._base(s, t)
        if not have_model_s:
            self.model_value(s, self.MS, _X)
        cx, cm = self._lin(s, t, h)
        self.ops.append(("axpy", dst, [(_X, cx), (self.MS, cm)], 0.0, -1))

    def second(self, s, t, r1=None, have_model_s=False, dst=_X):
        ns, h = self._base(s, t)
        r1 = 0.5 if r1 is None else r1
        s1 = ns.inverse_lambda(ns.marginal_lambda(s) + r1 * h)
        if not have_model_s:
            self.model_value(s, self.MS, _X)
        cx1, cm1 = self._lin(s, s1, r1 * h)
        self.ops.append(("axpy", self.XI, [(_X, cx1), (self.MS, cm1)], 0.0, -1))
        self.model_value(s1, self.MS1, self.XI)
        cx, cm = self._lin(s, t, h)                          # cm = -(alpha_t | sigma_t) * phi_1
        if self.st == "dpmsolver":
            d = (0.5 / r1) * cm                              # ... - (0.5 / r1) (a phi_1) (m1 - m0)
        elif self.pp:
            d = (1.0 / r1) * (ns.marginal_alpha(t) * (math.expm1(-h) / h + 1.0))
        else:
            d = -(1.0 / r1) * (ns.marginal_std(t) * (math.expm1(h) / h - 1.0))
        self.ops.append(("axpy", dst, [(_X, cx), (self.MS, cm - d), (self.MS1, d)], 0.0, -1))

    def third(self, s, t, r1=None, r2=None, have_model_s=False, have_model_s1=False, dst=_X):
        ns, h = self._base(s, t)
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        lam_s = ns.marginal_lambda(s)
        s1, s2 = ns.inverse_lambda(lam_s + r1 * h), ns.inverse_lambda(lam_s + r2 * h)
        if not have_model_s:
            self.model_value(s, self.MS, _X)
        if not have_model_s1:
            cx1, cm1 = self._lin(s, s1, r1 * h)
            self.ops.append(("axpy", self.XI, [(_X, cx1), (self.MS, cm1)], 0.0, -1))
            self.model_value(s1, self.MS1, self.XI)
        cx2, cm2 = self._lin(s, s2, r2 * h)
        if self.pp:
            phi_1 = math.expm1(-h)
            phi_22 = math.expm1(-r2 * h) / (r2 * h) + 1.0
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            a2, at = ns.marginal_alpha(s2), ns.marginal_alpha(t)
            d2 = r2 / r1 * (a2 * phi_22)                     # + d2 (m1 - m0)
            e1, e2, e3 = (1.0 / r2) * (at * phi_2), at * phi_2, -at * phi_3
        else:
            phi_1 = math.expm1(h)
            phi_22 = math.expm1(r2 * h) / (r2 * h) - 1.0
            phi_2 = phi_1 / h - 1.0
            phi_3 = phi_2 / h - 0.5
            g2, gt = ns.marginal_std(s2), ns.marginal_std(t)
            d2 = -r2 / r1 * (g2 * phi_22)
            e1, e2, e3 = -(1.0 / r2) * (gt * phi_2), -gt * phi_2, -gt * phi_3
        self.ops.append(("axpy", self.XI, [(_X, cx2), (self.MS, cm2 - d2), (self.MS1, d2)], 0.0, -1))
        self.model_value(s2, self.MS2, self.XI)
        cx, cm = self._lin(s, t, h)
        if self.st == "dpmsolver":
            self.ops.append(("axpy", dst, [(_X, cx), (self.MS, cm - e1), (self.MS2, e1)], 0.0, -1))
        else:
            # D1_0 = (m1 - m0) / r1, D1_1 = (m2 - m0) / r2, D1 = (r2 D1_0 - r1 D1_1) / (r2 - r1), D2 = 2 (D1_1 - D1_0) / (r2 - r1)
            k0 = (e2 * r2 - 2.0 * e3) / (r2 - r1) / r1       # coefficient of (m1 - m0)
            k1 = (-e2 * r1 + 2.0 * e3) / (r2 - r1) / r2      # coefficient of (m2 - m0)
            self.ops.append(("axpy", dst, [(_X, cx), (self.MS, cm - k0 - k1), (self.MS1, k0), (self.MS2, k1)], 0.0, -1))

    def denoise_to_zero(self, t_0):
        ns = self.ns
        self.times.append(ns.model_time(t_0))
        self.ops.append(("eval", ns.model_time(t_0), _X))
        cx, cr = self.wm.mix(t_0)
        a, s_ = ns.marginal_alpha(t_0), ns.marginal_std(t_0)
        cx, cr = (1.0 - s_ * cx) / a, -s_ * cr / a
        if self.model_type == "x_start":
            cx, cr = 0.0, 1.0
        terms = [(_RAW, cr)] if cx == 0.0 else [(_X, cx), (_RAW, cr)]
        self.ops.append(("axpy", _X, terms, 0.0, -1))
        if self.thr is not None:
            self.ops.append(("thresh", _X) + quantile_rank(self.thr[0], self.map_elems) + (float(self.thr[1]),))


def build_dpm_singlestep_program(ns, steps, order=2, algorithm_type="dpmsolver", model_type="x_start", skip_type="logSNR",
                                 method="singlestep", denoise_to_zero=True, solver_type="dpmsolver", t_start=None, t_end=None,
                                 thresholding=None, map_elems=224 * 384):
    """DPM_Solver.sample(method='singlestep' | 'singlestep_fixed') (sampler.py:1216-1239); returns (ops, model_times)."""
    t_0 = 1.0 / ns.total_N if t_end is None else t_end
    t_T = ns.T if t_start is None else t_start
    assert t_0 > 0 and t_T > 0
    if method == "singlestep":
        orders, K = singlestep_orders(steps, order)
        if skip_type == "logSNR":
            outer = dpm_time_steps(ns, skip_type, t_T, t_0, K)
        else:
            full = dpm_time_steps(ns, skip_type, t_T, t_0, steps)
            idx = [0]
            for o in orders:
                idx.append(idx[-1] + o)
            outer = [full[i] for i in idx]
    elif method == "singlestep_fixed":
        K = steps // order
        orders = [order] * K
        outer = dpm_time_steps(ns, skip_type, t_T, t_0, K)
    else:
        raise ValueError("Got wrong method {}".format(method))
    b = _SinglestepBuilder(ns, algorithm_type, model_type, solver_type, thresholding, map_elems)
    for step, o in enumerate(orders):
        s_, t_ = outer[step], outer[step + 1]            # IndexError for order 1 + logSNR, exactly like the reference
        # the reference round-trips s, t through fp32 (.item() of fp32 tensors) before it spaces the inner steps
        inner = dpm_time_steps(ns, skip_type, float(np.float32(s_)), float(np.float32(t_)), o)
        lam = [ns.marginal_lambda(v) for v in inner]
        h = lam[-1] - lam[0]
        r1 = None if o <= 1 else (lam[1] - lam[0]) / h
        r2 = None if o <= 2 else (lam[2] - lam[0]) / h
        if o == 1:
            b.first(s_, t_)
        elif o == 2:
            b.second(s_, t_, r1)
        else:
            b.third(s_, t_, r1, r2)
    if denoise_to_zero:
        b.denoise_to_zero(t_0)
    return b.ops, b.times


def _adaptive_error(x_lower, x_higher, x_prev, atol, rtol):
    """max over the batch of dpm_solver_adaptive's error norm (sampler.py:996-999), computed on the device."""
    import ctypes
    from . import _lib
    from .engine import _bind, _stream
    lib = _bind(_lib.lib())
    B = x_lower.shape[0]
    out = torch.empty(B, dtype=torch.float32, device=x_lower.device)
    with torch.cuda.device(x_lower.device):
        rc = lib.dsb_sampler_adaptive_error(_lib.ptr(x_lower), _lib.ptr(x_higher), _lib.ptr(x_prev), B, x_lower.numel() // B,
                                            float(atol), float(rtol), _lib.ptr(out), _stream())
    if rc != 0:
        raise RuntimeError("dsb_sampler_adaptive_error failed (%d)" % rc)
    return float(out.max().item())


def sample_dpm_adaptive(ns, x, model, order=2, algorithm_type="dpmsolver", model_type="x_start", t_T=None, t_0=None,
                        h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5, solver_type="dpmsolver"):
    """DPM_Solver.dpm_solver_adaptive (sampler.py:958-1009).  The step-size control is data dependent, so this is a host
    loop: per trial step one small program (lower- and higher-order update sharing their evaluations) through the fused
    update / denoiser kernels, the error norm on the device, one scalar read back.  ``model(x, t[B]) -> raw output``.
    Returns (x_0, nfe)."""
    if order not in (2, 3):
        raise ValueError("For adaptive step size solver, order must be 2 or 3, got {}".format(order))
    t_T = ns.T if t_T is None else t_T
    t_0 = 1.0 / ns.total_N if t_0 is None else t_0
    s = t_T
    lam_s, lam_0 = ns.marginal_lambda(s), ns.marginal_lambda(t_0)
    h = h_init
    x = x.to(dtype=torch.float32).contiguous().clone()
    x_prev = x
    nfe = 0
    LO = 6                                                    # buffer of the lower-order iterate
    while abs(s - t_0) > t_err:
        t = ns.inverse_lambda(lam_s + h)
        b = _SinglestepBuilder(ns, algorithm_type, model_type, solver_type)
        if order == 2:
            b.first(s, t, dst=LO)
            b.second(s, t, 0.5, have_model_s=True, dst=7)
        else:
            b.second(s, t, 1.0 / 3.0, dst=LO)
            b.third(s, t, 1.0 / 3.0, 2.0 / 3.0, have_model_s=True, have_model_s1=True, dst=7)
        bufs = run_program_generic(b.ops, x, model, return_buffers=True)
        x_lower, x_higher = bufs[LO], bufs[7]
        E = _adaptive_error(x_lower, x_higher, x_prev, atol, rtol)
        if E <= 1.0:
            x, s, x_prev = x_higher, t, x_lower
            lam_s = ns.marginal_lambda(s)
        h = min(theta * h * float(np.float32(float(np.float32(E)) ** (-1.0 / order))), lam_0 - lam_s)
        nfe += order
    return x, nfe


def build_ddim_program(tables, timesteps, eta=0.0, training_target="x0"):
    """DiffusionTrainer.sample_ddim (diffusion_trainer.py:439-480) as EVAL/AXPY ops.
    Returns (ops, n_noise_slabs); noise slab k belongs to the k-th non-final step."""
    skip = tables.num_timesteps // timesteps
    seq = list(range(0, tables.num_timesteps, skip))
    seq_next = [-1] + seq[:-1]
    ops, n_noise = [], 0
    for time, time_next in zip(reversed(seq), reversed(seq_next)):
        ops.append(("eval", float(time)))
        alpha = tables.alphas_hat[time]
        if training_target == "x0":
            a, b = tables.sqrt_recip_alphas_hat[time], tables.sqrt_recipm1_alphas_hat[time]
            xs_x, xs_r = 0.0, 1.0                           # x_start = raw
            pn_x, pn_r = a / b, -1.0 / b                    # pred_noise = (a x - raw) / b
        else:
            pn_x, pn_r = 0.0, 1.0
            xs_x, xs_r = 1.0 / math.sqrt(alpha), -math.sqrt(1 - alpha) / math.sqrt(alpha)
        if time_next < 0:
            terms = [(_RAW, xs_r)] if xs_x == 0.0 else [(_X, xs_x), (_RAW, xs_r)]
            ops.append(("axpy", _X, terms, 0.0, -1))
            continue
        alpha_next = tables.alphas_hat[time_next]
        c1 = eta * math.sqrt((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha))
        c2 = math.sqrt((1 - alpha_next) - c1 ** 2)
        sa = tables.sqrt_alphas_hat[time_next]
        terms = [(_X, sa * xs_x + c2 * pn_x), (_RAW, sa * xs_r + c2 * pn_r)]
        if c1 != 0.0:
            ops.append(("axpy", _X, terms, c1, n_noise))
            n_noise += 1
        else:
            ops.append(("axpy", _X, terms, 0.0, -1))
    return ops, n_noise


def build_ddpm_program(tables, timesteps, training_target="x0"):
    """DiffusionTrainer.sample_ddpm / p_sample (diffusion_trainer.py:488-540) as EVAL/AXPY ops: per step
    x <- coef1[t] * x_recon + coef2[t] * x + exp(0.5 * logvar[t]) * z  (z only for t > 0; the reference's
    ``x_recon.clamp(-1, 1)`` at :507 is not in-place and has no effect).  Returns (ops, n_noise_slabs)."""
    skip = tables.num_timesteps // timesteps
    ops, n_noise = [], 0
    for time in reversed(range(0, tables.num_timesteps, skip)):
        ops.append(("eval", float(time)))
        c1, c2 = tables.posterior_mean_coef1[time], tables.posterior_mean_coef2[time]
        if training_target == "x0":
            xr_x, xr_r = 0.0, 1.0
        else:
            xr_x, xr_r = tables.sqrt_recip_alphas_hat[time], -tables.sqrt_recipm1_alphas_hat[time]
        terms = [(_X, c2 + c1 * xr_x), (_RAW, c1 * xr_r)]
        if time > 0:
            ops.append(("axpy", _X, terms, math.exp(0.5 * tables.posterior_log_variance_clipped[time]), n_noise))
            n_noise += 1
        else:
            ops.append(("axpy", _X, terms, 0.0, -1))
    return ops, n_noise


def _correct(kind, buf, args):
    """'clamp' / 'thresh' program ops outside dsb_sample (generic denoiser path): same CUDA kernels, one-op program."""
    import ctypes
    from . import _lib
    from .engine import _bind, _stream
    if not buf.is_cuda:
        raise RuntimeError("diff_sal_b200 sampler correctors run on the GPU only")
    lib = _bind(_lib.lib())
    buf = buf.to(dtype=torch.float32).contiguous()
    B = buf.shape[0]
    with torch.cuda.device(buf.device):
        if kind == "clamp":
            rc = lib.dsb_sampler_clamp(_lib.ptr(buf), buf.numel(), float(args[0]), float(args[1]), _stream())
        else:
            rc = lib.dsb_sampler_dynamic_threshold(_lib.ptr(buf), B, buf.numel() // B, int(args[0]), float(args[1]),
                                                   float(args[2]), _stream())
    if rc != 0:
        raise RuntimeError("sampler corrector %r failed (%d)" % (kind, rc))
    return buf


def run_program_generic(ops, x, model, noise=None, return_buffers=False):
    """Executes a sampler program with an arbitrary denoiser callable ``model(x, t[B]) -> tensor`` (one fused
    update launch per AXPY).  The SalUNetB200 fast path runs the same ops inside dsb_sample instead."""
    bufs = {_X: x}
    for op in ops:
        if op[0] == "eval":
            t = torch.full((x.shape[0],), op[1], dtype=torch.float32, device=x.device)
            bufs[_RAW] = model(bufs[op[2] if len(op) > 2 else _X], t)
        elif op[0] in ("clamp", "thresh"):
            bufs[op[1]] = _correct(op[0], bufs[op[1]], op[2:])
        else:
            _, dst, terms, ncoef, nidx = op
            nz = noise[nidx] if (nidx >= 0 and noise is not None) else None
            bufs[dst] = _axpy([c for _, c in terms], [bufs[s] for s, _ in terms], nz, ncoef)
    return bufs if return_buffers else bufs[_X]


# ============================================================================================ DPM_Solver


class DPM_Solver:
    """sampler.py:336-1247: multistep, singlestep / singlestep_fixed (compiled to EVAL/AXPY programs) and adaptive (host
    loop).  ``model_fn`` is what model_wrapper returned."""

    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        assert algorithm_type in ["dpmsolver", "dpmsolver++"]
        if correcting_xt_fn is not None or correcting_x0_fn not in (None, "dynamic_thresholding"):
            raise NotImplementedError("only correcting_x0_fn in (None, 'dynamic_thresholding') runs in the fused loop; "
                                      "Python callables per step are outside the hot path")
        self.thresholding = (float(dynamic_thresholding_ratio), float(thresholding_max_val)) \
            if correcting_x0_fn == "dynamic_thresholding" else None
        if not isinstance(model_fn, _WrappedModel):
            raise TypeError("model_fn must come from diff_sal_b200.sampler.model_wrapper")
        self.wrapped = model_fn
        self.noise_schedule = noise_schedule
        self.algorithm_type = algorithm_type

    def sample(self, x, img=None, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform",
               method="multistep", lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver",
               atol=0.0078, rtol=0.05, return_intermediate=False, use_graph=True):
        if return_intermediate:
            raise NotImplementedError("return_intermediate is not supported by the fused loop")
        ns = self.noise_schedule
        model = self.wrapped.model
        kwargs = self.wrapped.model_kwargs
        if method == "adaptive":
            # data-dependent step sizes: host loop around the same kernels (sampler.py:958-1009,1171-1172).  Unlike the
            # reference, whose adaptive branch never forwards ``img`` to the network, the conditioning is applied.
            if self.thresholding is not None:
                raise NotImplementedError("dynamic thresholding with the adaptive solver is outside the hot path")
            xs, _ = sample_dpm_adaptive(ns, x, lambda x_, t_: model(x_, t_, img, **kwargs), order, self.algorithm_type,
                                        self.wrapped.model_type, t_start, t_end, atol=atol, rtol=rtol, solver_type=solver_type)
            if denoise_to_zero:
                b = _SinglestepBuilder(ns, self.algorithm_type, self.wrapped.model_type, solver_type)
                b.denoise_to_zero(1.0 / ns.total_N if t_end is None else t_end)
                xs = run_program_generic(b.ops, xs, lambda x_, t_: model(x_, t_, img, **kwargs))
            return xs
        if method == "multistep":
            ops, _ = build_dpm_program(ns, steps, order, self.algorithm_type, self.wrapped.model_type, skip_type,
                                       lower_order_final, denoise_to_zero, solver_type, t_start, t_end,
                                       thresholding=self.thresholding, map_elems=x[0].numel())
        elif method in ("singlestep", "singlestep_fixed"):
            ops, _ = build_dpm_singlestep_program(ns, steps, order, self.algorithm_type, self.wrapped.model_type, skip_type,
                                                  method, denoise_to_zero, solver_type, t_start, t_end,
                                                  thresholding=self.thresholding, map_elems=x[0].numel())
        else:
            raise ValueError("Got wrong method {}".format(method))
        fused = getattr(model, "_dsb_fused_sample", None)
        if fused is not None:
            return fused(ops, x, img, kwargs, use_graph=use_graph)
        return run_program_generic(ops, x.clone(), lambda x_, t_: model(x_, t_, img, **kwargs))


# ============================================================================================ DDIM driver


class DiffusionSampler:
    """The sampling half of the reference's DiffusionTrainer (diffusion_trainer.py:29-76,434-640) around a
    SalUNetB200 decoder: same config keys (config.sampling.*, config.training.training_target), same
    ``sample_ddim(x, img, audio_cond)`` / ``sample_image(x, img, audio)`` call shapes.  ``img`` is the MViT feature
    list and ``audio`` the [B,512,9,7,12] audio feature tensor (the encoders are outside the hot path)."""

    def __init__(self, decoder_net, config, betas=None):
        self.decoder_net = decoder_net
        self.config = config
        self.training_target = config.training.training_target
        assert self.training_target in ["x0", "noise"]
        if betas is None:
            betas = get_beta_schedule(beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
                                      beta_end=config.diffusion.beta_end,
                                      num_diffusion_timesteps=config.diffusion.num_diffusion_timesteps)
        self.betas = to_torch(betas)
        self.tables = DdimTables(self.betas)
        self.num_timesteps = self.tables.num_timesteps

    @torch.no_grad()
    def sample_ddim(self, x, img=None, audio_cond=None, use_graph=True):
        eta = self.config.sampling.eta
        ops, n_noise = build_ddim_program(self.tables, self.config.sampling.timesteps, eta, self.training_target)
        noise = torch.randn((n_noise,) + tuple(x.shape), device=x.device) if n_noise else None
        return self.decoder_net._dsb_fused_sample(ops, x, img, {"audio_feat_list": audio_cond}, noise=noise,
                                                  use_graph=use_graph)

    @torch.no_grad()
    def sample_ddpm(self, x, img=None, audio_cond=None, use_graph=True, noise=None):
        """Ancestral sampling (diffusion_trainer.py:522-540).  The reference's own method re-runs the encoders and feeds
        an undefined feature list (:529-531); here ``img`` / ``audio_cond`` are the decoder's conditioning as in
        ``sample_ddim``.  ``noise`` [n_steps-1, B, 1, H, W] may be given for reproducibility (default: randn)."""
        ops, n_noise = build_ddpm_program(self.tables, self.config.sampling.timesteps, self.training_target)
        if noise is None and n_noise:
            noise = torch.randn((n_noise,) + tuple(x.shape), device=x.device)
        return self.decoder_net._dsb_fused_sample(ops, x, img, {"audio_feat_list": audio_cond}, noise=noise,
                                                  use_graph=use_graph)

    @torch.no_grad()
    def sample_image(self, x, img=None, audio=None, base_samples=None, use_graph=True):
        st = self.config.sampling.sample_type
        if st == "ddim":
            return self.sample_ddim(x, img, audio, use_graph=use_graph)
        if st == "ddpm":
            return self.sample_ddpm(x, img, audio, use_graph=use_graph)
        if st in ["dpmsolver", "dpmsolver++"]:
            ns = NoiseScheduleVP(schedule="discrete", betas=self.betas)
            mtype = "x_start" if self.training_target == "x0" else "noise"
            mf = model_wrapper(self.decoder_net, ns, model_type=mtype, model_kwargs={"audio_feat_list": audio},
                               guidance_type="uncond")
            s = self.config.sampling
            solver = DPM_Solver(mf, ns, algorithm_type=st,
                                correcting_x0_fn="dynamic_thresholding" if getattr(s, "thresholding", False) else None)
            return solver.sample(x, img, steps=(s.timesteps - 1 if s.denoise else s.timesteps), order=s.dpm_solver_order,
                                 skip_type=s.skip_type, method=s.dpm_solver_method, lower_order_final=s.lower_order_final,
                                 denoise_to_zero=s.denoise, solver_type=s.dpm_solver_type, atol=s.dpm_solver_atol,
                                 rtol=s.dpm_solver_rtol, use_graph=use_graph)
        raise NotImplementedError(st)


# ============================================================================================ output side


def _postprocess(x, want_u8):
    import ctypes
    from . import _lib
    from .engine import _bind, _stream
    if not x.is_cuda:
        raise RuntimeError("diff_sal_b200 post-processing runs on the GPU only")
    lib = _bind(_lib.lib())
    lib.dsb_postprocess.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    x = x.to(dtype=torch.float32).contiguous()
    B = x.shape[0]
    n = x.numel() // B
    clamped = torch.empty_like(x)
    u8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None
    with torch.cuda.device(x.device):
        rc = lib.dsb_postprocess(_lib.ptr(x), B, n, _lib.ptr(clamped), _lib.ptr(u8), _stream())
    if rc != 0:
        raise RuntimeError("dsb_postprocess failed (%d)" % rc)
    return clamped, u8


def inverse_data_transform(config, X):
    """datasets/__init__.py:26-35 for the shipped data config (no image_mean, no logit transform, not rescaled,
    cfgs/diffusion.yml:1-8): clamp to [0, 1].  Other data configs are outside the hot path and raise."""
    d = getattr(config, "data", None)
    if hasattr(config, "image_mean") or (d is not None and (getattr(d, "logit_transform", False) or getattr(d, "rescaled", False))):
        raise NotImplementedError("only the shipped data config (no mean / logit / rescale) is on the hot path")
    return _postprocess(X, False)[0]


def normalize_data(maps):
    """util/utils.py:11-16 per map: (x - min) * 255 / (max - min) clipped to [0, 255] as uint8 (of the clamped map,
    which is what DiffusionTrainer.save_img feeds it, diffusion_trainer.py:898-935).  maps: [B,1,H,W] CUDA tensor."""
    return _postprocess(maps, True)[1]


# ============================================================================================ util/denoising.py


def compute_alpha(beta, t):
    """util/denoising.py:3-6."""
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def generalized_steps(x, seq, model, b, img=None, **kwargs):
    """util/denoising.py:9-36 (DDIM repo loop, eps-parameterised ``model(data, t)``).  The per-step .to('cuda') /
    .to('cpu') round trips of the reference are not reproduced; xs / x0_preds stay on x's device."""
    with torch.no_grad():
        n = x.size(0)
        seq = list(seq)
        seq_next = [-1] + seq[:-1]
        bcpu = _as_fp32_cpu(b)
        ah = torch.cat([torch.ones(1), (1 - bcpu).cumprod(dim=0)]).double().numpy()   # index t+1
        eta = kwargs.get("eta", 0)
        xs, x0_preds = [x], []
        for i, j in zip(reversed(seq), reversed(seq_next)):
            t = (torch.ones(n) * i).to(x.device)
            at, at_next = float(ah[i + 1]), float(ah[j + 1])
            xt = xs[-1]
            et = model({"img": img, "input": xt}, t)
            x0_t = _axpy([1.0 / math.sqrt(at), -math.sqrt(1 - at) / math.sqrt(at)], [xt, et])
            x0_preds.append(x0_t)
            c1 = eta * math.sqrt((1 - at / at_next) * (1 - at_next) / (1 - at))
            c2 = math.sqrt((1 - at_next) - c1 ** 2)
            nz = torch.randn_like(x) if c1 != 0.0 else None
            xs.append(_axpy([math.sqrt(at_next), c2], [x0_t, et], nz, c1))
    return xs, x0_preds


def ddpm_steps(x, seq, model, b, **kwargs):
    """util/denoising.py:39-67 (DDIM-repo ancestral loop, eps-parameterised ``model(x, t_float)``): per step
    x0 = clamp(sqrt(1/a_t) x - sqrt(1/a_t - 1) e, -1, 1); mean = (sqrt(a_{t-1}) beta_t x0 + sqrt(1 - beta_t)
    (1 - a_{t-1}) x) / (1 - a_t); x <- mean + [t != 0] * sqrt(beta_t) * z.  As in generalized_steps the reference's
    per-step host round trips are not reproduced.  ``noise`` (list of tensors, one per step) may be passed for
    reproducibility."""
    with torch.no_grad():
        n = x.size(0)
        seq = list(seq)
        seq_next = [-1] + seq[:-1]
        bcpu = _as_fp32_cpu(b)
        ah = torch.cat([torch.ones(1), (1 - bcpu).cumprod(dim=0)]).double().numpy()   # index t+1
        noise = kwargs.get("noise")
        xs, x0_preds = [x], []
        for step, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            t = (torch.ones(n) * i).to(x.device)
            at, atm1 = float(ah[i + 1]), float(ah[j + 1])
            beta_t = 1 - at / atm1
            xt = xs[-1]
            e = model(xt, t.float())
            x0 = _correct("clamp", _axpy([math.sqrt(1.0 / at), -math.sqrt(1.0 / at - 1)], [xt, e]), (-1.0, 1.0))
            x0_preds.append(x0)
            z = noise[step] if noise is not None else torch.randn_like(x)
            mask = 0.0 if i == 0 else 1.0
            xs.append(_axpy([math.sqrt(atm1) * beta_t / (1.0 - at), math.sqrt(1 - beta_t) * (1 - atm1) / (1.0 - at)],
                            [x0, xt], z, mask * math.sqrt(beta_t)))
    return xs, x0_preds
