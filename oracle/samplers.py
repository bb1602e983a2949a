"""ORACLE (test infrastructure, not product code): fp32 restatement of the reference's
sampling loops around a denoiser callable ``model(x, t) -> prediction``.

Restates
  get_beta_schedule('cosine') + tables   models/diffusion_decoder/diffusion_utils.py:34-41,
                                         diffusion_trainer.py:47-76
  DiffusionTrainer.sample_ddim           diffusion_trainer.py:439-480 (+ :434-437)
  NoiseScheduleVP (discrete)             models/dpm_solver/sampler.py:6-167
  interpolate_fn                         sampler.py:1255-1294
  model_wrapper (uncond; noise/x_start)  sampler.py:170-334
  DPM_Solver.sample, multistep path      sampler.py:1048-1247, updates :548-593, :797-905,
                                         data_prediction_fn :434-443

Scalars are kept as 1-element fp32 torch tensors and combined in the reference's
expression order so that the restatement tracks the reference to fp32 round-off.
Pinned against the imported reference in tests/test_oracle_vs_reference.py.
"""
import math

import numpy as np
import torch


# ----------------------------------------------------------------------------- schedule
def cosine_betas(n=1000):
    """diffusion_utils.py:34-41 (float64 numpy)."""
    step = n + 1
    s = 0.008
    x = np.linspace(0, step, step)
    ac = np.cos(((x / step) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


def betas_fp32(n=1000):
    return torch.tensor(cosine_betas(n), dtype=torch.float32)


class DdimTables:
    """diffusion_trainer.py:47-76: fp32 cumprod of fp32 (1 - beta)."""

    def __init__(self, betas=None):
        betas = betas_fp32() if betas is None else betas
        self.betas = betas
        self.alphas_hat = (1.0 - betas).cumprod(dim=0)
        self.sqrt_alphas_hat = torch.sqrt(self.alphas_hat)
        self.sqrt_recip_alphas_hat = torch.sqrt(1.0 / self.alphas_hat)
        self.sqrt_recipm1_alphas_hat = torch.sqrt(1.0 / self.alphas_hat - 1)
        self.num_timesteps = betas.shape[0]
        # posterior q(x_{t-1} | x_t, x_0) terms (diffusion_trainer.py:55,66-74)
        alphas = 1.0 - betas
        prev = torch.cat([torch.ones(1), self.alphas_hat[:-1]], dim=0)
        self.posterior_variance = betas * (1.0 - prev) / (1.0 - self.alphas_hat)
        self.posterior_log_variance_clipped = torch.log(torch.maximum(self.posterior_variance, torch.tensor(1e-20)))
        self.posterior_mean_coef1 = betas * torch.sqrt(self.alphas_hat) / (1.0 - self.alphas_hat)
        self.posterior_mean_coef2 = (1.0 - prev) * torch.sqrt(alphas) / (1.0 - self.alphas_hat)


def sample_ddim(model, x, timesteps, eta=0.0, training_target="x0", tables=None, noise_fn=None):
    """diffusion_trainer.py:439-480.  ``model(x, t_int64[B])``.  ``noise_fn(x)`` supplies the
    eta-noise (default randn_like; with eta == 0 it is multiplied by c1 == 0)."""
    tb = DdimTables() if tables is None else tables
    skip = tb.num_timesteps // timesteps
    seq = list(range(0, tb.num_timesteps, skip))
    seq_next = [-1] + seq[:-1]
    n = x.shape[0]
    for time, time_next in zip(reversed(seq), reversed(seq_next)):
        t = torch.full((n,), time, dtype=torch.int64)
        alpha = tb.alphas_hat[time]
        alpha_next = tb.alphas_hat[time_next]
        if training_target == "x0":
            x_start = model(x, t)
            pred_noise = (tb.sqrt_recip_alphas_hat[time] * x - x_start) / tb.sqrt_recipm1_alphas_hat[time]
        else:
            pred_noise = model(x, t)
            x_start = (x - pred_noise * (1 - alpha).sqrt()) / alpha.sqrt()
        if time_next < 0:
            x = x_start
            continue
        c1 = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c2 = ((1 - alpha_next) - c1 ** 2).sqrt()
        z = torch.randn_like(x) if noise_fn is None else noise_fn(x)
        x = tb.sqrt_alphas_hat[time_next] * x_start + c1 * z + c2 * pred_noise
    return x


def sample_ddpm(model, x, timesteps, training_target="x0", tables=None, noise_fn=None):
    """DiffusionTrainer.sample_ddpm / p_sample / p_mean_variance / q_posterior (diffusion_trainer.py:488-540) around
    ``model(x, t_int64[B])``.  The reference's ``x_recon.clamp(-1.0, 1.0)`` (:507) is not in-place and therefore has
    no effect; ``posterior_mean_coef1`` is ``betas * sqrt(alphas_hat) / (1 - alphas_hat)`` exactly as written at :71
    (not the textbook sqrt(alphas_hat_prev)).  ``noise_fn(x)`` supplies the ancestral noise (default randn_like,
    drawn only for t > 0 as in :518-520)."""
    tb = DdimTables() if tables is None else tables
    skip = tb.num_timesteps // timesteps
    n = x.shape[0]
    for time in reversed(range(0, tb.num_timesteps, skip)):
        t = torch.full((n,), time, dtype=torch.int64)
        if training_target == "x0":
            x_recon = model(x, t)
        else:
            x_recon = tb.sqrt_recip_alphas_hat[time] * x - tb.sqrt_recipm1_alphas_hat[time] * model(x, t)
        mean = tb.posterior_mean_coef1[time] * x_recon + tb.posterior_mean_coef2[time] * x
        if time > 0:
            z = torch.randn_like(x) if noise_fn is None else noise_fn(x)
        else:
            z = torch.zeros_like(x)
        x = mean + z * (0.5 * tb.posterior_log_variance_clipped[time]).exp()
    return x


def ddpm_steps(x, seq, model, b, noise_fn=None):
    """util/denoising.py:39-67 (``model(x, t_float[B])`` predicts eps).  The reference's ``.to('cuda')`` /
    ``.to('cpu')`` hops are device plumbing and are dropped; arithmetic and call order are the reference's."""
    n = x.size(0)
    seq = list(seq)
    seq_next = [-1] + seq[:-1]
    xs, x0_preds = [x], []

    def compute_alpha(beta, t):                                     # util/denoising.py:3-6
        beta = torch.cat([torch.zeros(1), beta], dim=0)
        return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)

    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = torch.ones(n) * i
        next_t = torch.ones(n) * j
        at = compute_alpha(b, t.long())
        atm1 = compute_alpha(b, next_t.long())
        beta_t = 1 - at / atm1
        x = xs[-1]
        e = model(x, t.float())
        x0_from_e = (1.0 / at).sqrt() * x - (1.0 / at - 1).sqrt() * e
        x0_from_e = torch.clamp(x0_from_e, -1, 1)
        x0_preds.append(x0_from_e)
        mean = ((atm1.sqrt() * beta_t) * x0_from_e + ((1 - beta_t).sqrt() * (1 - atm1)) * x) / (1.0 - at)
        noise = torch.randn_like(x) if noise_fn is None else noise_fn(x)
        mask = (1 - (t == 0).float()).view(-1, 1, 1, 1)
        xs.append(mean + mask * torch.exp(0.5 * beta_t.log()) * noise)
    return xs, x0_preds


def dynamic_thresholding(x0, ratio=0.995, max_val=1.0):
    """DPM_Solver.dynamic_thresholding_fn (sampler.py:417-426)."""
    s = torch.quantile(torch.abs(x0).reshape((x0.shape[0], -1)), ratio, dim=1)
    s = torch.maximum(s, max_val * torch.ones_like(s)).reshape((-1,) + (1,) * (x0.dim() - 1))
    return torch.clamp(x0, -s, s) / s


# ----------------------------------------------------------------------------- DPM-solver
def _interp(x, xp, yp):
    """sampler.py:1255-1294 for a 1-element x and ascending 1-D xp: piecewise linear, the
    outermost segments extended beyond the ends."""
    K = xp.shape[0]
    idx = int(torch.searchsorted(xp, x.reshape(())).item())   # number of keypoints < x
    if idx == 0:
        i0 = 0
    elif idx == K:
        i0 = K - 2
    else:
        i0 = idx - 1
    x0, x1, y0, y1 = xp[i0], xp[i0 + 1], yp[i0], yp[i0 + 1]
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0)


class NoiseScheduleDiscrete:
    """sampler.py:71-82,114-167 with schedule='discrete', betas given."""

    def __init__(self, betas):
        log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(-5.1)).item())
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        self.T = 1.0
        self.log_alpha_array = log_alphas.float()
        self.total_N = log_alphas.shape[0]
        self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].float()

    def log_alpha(self, t):
        return _interp(t, self.t_array, self.log_alpha_array)

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def sigma(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))

    def inverse_lambda(self, lamb):
        log_alpha = -0.5 * torch.logaddexp(torch.zeros(()), -2.0 * lamb)
        return _interp(log_alpha, torch.flip(self.log_alpha_array, [0]), torch.flip(self.t_array, [0]))


def dpm_time_steps(ns, steps, skip_type="logSNR", t_T=None, t_0=None):
    """sampler.py:454-481 (t_T / t_0: python floats, default the whole range)."""
    t_T = ns.T if t_T is None else t_T
    t_0 = 1.0 / ns.total_N if t_0 is None else t_0
    if skip_type == "logSNR":
        lam_T = ns.lam(torch.tensor(t_T))
        lam_0 = ns.lam(torch.tensor(t_0))
        ls = torch.linspace(lam_T.item(), lam_0.item(), steps + 1)
        return [ns.inverse_lambda(l) for l in ls]
    if skip_type == "time_uniform":
        return list(torch.linspace(t_T, t_0, steps + 1))
    if skip_type == "time_quadratic":
        return list(torch.linspace(t_T ** 0.5, t_0 ** 0.5, steps + 1).pow(2))
    raise ValueError(skip_type)


def sample_dpm(model, x, betas=None, steps=9, order=2, algorithm_type="dpmsolver",
               model_type="x_start", skip_type="logSNR", lower_order_final=False,
               denoise_to_zero=True, solver_type="dpmsolver", return_model_times=False,
               correcting_x0_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
    """DPM_Solver.sample(method='multistep') driven through model_wrapper(uncond).
    ``model(x, t_float[B])`` is the raw network; model_type says what it predicts."""
    betas = betas_fp32() if betas is None else betas
    ns = NoiseScheduleDiscrete(betas)
    n = x.shape[0]
    model_times = []

    def noise_pred(x, t):
        t_in = (t - 1.0 / ns.total_N) * 1000.0
        model_times.append(float(t_in))
        out = model(x, t_in.reshape(1).expand(n))
        if model_type == "noise":
            return out
        if model_type == "x_start":
            return (x - ns.alpha(t) * out) / ns.sigma(t)
        raise ValueError(model_type)

    def data_pred(x, t):
        noise = noise_pred(x, t)
        x0 = (x - ns.sigma(t) * noise) / ns.alpha(t)
        if correcting_x0_fn == "dynamic_thresholding":       # sampler.py:410-411,434-443
            x0 = dynamic_thresholding(x0, dynamic_thresholding_ratio, thresholding_max_val)
        elif correcting_x0_fn is not None:
            x0 = correcting_x0_fn(x0, t)
        return x0

    def model_fn(x, t):
        return data_pred(x, t) if algorithm_type == "dpmsolver++" else noise_pred(x, t)

    def update(x, ms, ts, t, k):
        lam_t, lam_0 = ns.lam(t), ns.lam(ts[-1])
        h = lam_t - lam_0
        la_0, la_t = ns.log_alpha(ts[-1]), ns.log_alpha(t)
        sig_0, sig_t = ns.sigma(ts[-1]), ns.sigma(t)
        alpha_t = torch.exp(la_t)
        pp = algorithm_type == "dpmsolver++"
        phi_1 = torch.expm1(-h) if pp else torch.expm1(h)
        if k == 1:
            if pp:
                return sig_t / sig_0 * x - alpha_t * phi_1 * ms[-1]
            return torch.exp(la_t - la_0) * x - (sig_t * phi_1) * ms[-1]
        if k == 2:
            h_0 = lam_0 - ns.lam(ts[-2])
            r0 = h_0 / h
            D1 = (1.0 / r0) * (ms[-1] - ms[-2])
            if pp:
                if solver_type == "dpmsolver":
                    return (sig_t / sig_0) * x - (alpha_t * phi_1) * ms[-1] - 0.5 * (alpha_t * phi_1) * D1
                return (sig_t / sig_0) * x - (alpha_t * phi_1) * ms[-1] + (alpha_t * (phi_1 / h + 1.0)) * D1
            if solver_type == "dpmsolver":
                return torch.exp(la_t - la_0) * x - (sig_t * phi_1) * ms[-1] - 0.5 * (sig_t * phi_1) * D1
            return torch.exp(la_t - la_0) * x - (sig_t * phi_1) * ms[-1] - (sig_t * (phi_1 / h - 1.0)) * D1
        if k == 3:
            lam_1, lam_2 = ns.lam(ts[-2]), ns.lam(ts[-3])
            h_1, h_0 = lam_1 - lam_2, lam_0 - lam_1
            r0, r1 = h_0 / h, h_1 / h
            D1_0 = (1.0 / r0) * (ms[-1] - ms[-2])
            D1_1 = (1.0 / r1) * (ms[-2] - ms[-3])
            D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
            D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
            if pp:
                phi_2 = phi_1 / h + 1.0
                phi_3 = phi_2 / h - 0.5
                return (sig_t / sig_0) * x - (alpha_t * phi_1) * ms[-1] + (alpha_t * phi_2) * D1 - (alpha_t * phi_3) * D2
            phi_2 = phi_1 / h - 1.0
            phi_3 = phi_2 / h - 0.5
            return torch.exp(la_t - la_0) * x - (sig_t * phi_1) * ms[-1] - (sig_t * phi_2) * D1 - (sig_t * phi_3) * D2
        raise ValueError(k)

    assert steps >= order
    ts_all = dpm_time_steps(ns, steps, skip_type)
    t_prev = [ts_all[0]]
    m_prev = [model_fn(x, ts_all[0])]
    for step in range(1, order):
        t = ts_all[step]
        x = update(x, m_prev, t_prev, t, step)
        t_prev.append(t)
        m_prev.append(model_fn(x, t))
    for step in range(order, steps + 1):
        t = ts_all[step]
        k = min(order, steps + 1 - step) if (lower_order_final and steps < 10) else order
        x = update(x, m_prev, t_prev, t, k)
        t_prev = t_prev[1:] + [t]
        if step < steps:
            m_prev = m_prev[1:] + [model_fn(x, t)]
        else:
            m_prev = m_prev[1:] + [None]
    if denoise_to_zero:
        x = data_pred(x, torch.ones(()) * (1.0 / ns.total_N))
    if return_model_times:
        return x, model_times
    return x


def _dpm_model_fns(model, ns, n, algorithm_type, model_type, model_times):
    """model_wrapper(uncond) + DPM_Solver.{noise_prediction_fn, data_prediction_fn, model_fn} (sampler.py:282-298,428-452)."""
    def noise_pred(x, t):
        t_in = (t - 1.0 / ns.total_N) * 1000.0
        model_times.append(float(t_in))
        out = model(x, t_in.reshape(1).expand(n))
        if model_type == "noise":
            return out
        if model_type == "x_start":
            return (x - ns.alpha(t) * out) / ns.sigma(t)
        raise ValueError(model_type)

    def data_pred(x, t):
        return (x - ns.sigma(t) * noise_pred(x, t)) / ns.alpha(t)

    model_fn = data_pred if algorithm_type == "dpmsolver++" else noise_pred
    return noise_pred, data_pred, model_fn


class _Singlestep:
    """dpm_solver_first_update / singlestep_dpm_solver_second_update / _third_update (sampler.py:548-795).  The reference
    calls ``self.model_fn(x, s)`` WITHOUT the conditioning there (:573,585,630,635,...); ``model_fn`` here is a closure that
    already carries it, which is the only deviation ("singlestep/adaptive paths with conditioning")."""

    def __init__(self, ns, model_fn, algorithm_type, solver_type):
        self.ns, self.model_fn, self.pp, self.solver_type = ns, model_fn, algorithm_type == "dpmsolver++", solver_type

    def first(self, x, s, t, model_s=None, return_intermediate=False):
        ns = self.ns
        h = ns.lam(t) - ns.lam(s)
        la_s, la_t = ns.log_alpha(s), ns.log_alpha(t)
        sig_s, sig_t = ns.sigma(s), ns.sigma(t)
        alpha_t = torch.exp(la_t)
        if model_s is None:
            model_s = self.model_fn(x, s)
        if self.pp:
            x_t = sig_t / sig_s * x - alpha_t * torch.expm1(-h) * model_s
        else:
            x_t = torch.exp(la_t - la_s) * x - (sig_t * torch.expm1(h)) * model_s
        return (x_t, {"model_s": model_s}) if return_intermediate else x_t

    def second(self, x, s, t, r1=0.5, model_s=None, return_intermediate=False):
        ns, st = self.ns, self.solver_type
        if r1 is None:
            r1 = 0.5
        lam_s, lam_t = ns.lam(s), ns.lam(t)
        h = lam_t - lam_s
        s1 = ns.inverse_lambda(lam_s + r1 * h)
        la_s, la_s1, la_t = ns.log_alpha(s), ns.log_alpha(s1), ns.log_alpha(t)
        sig_s, sig_s1, sig_t = ns.sigma(s), ns.sigma(s1), ns.sigma(t)
        alpha_s1, alpha_t = torch.exp(la_s1), torch.exp(la_t)
        if model_s is None:
            model_s = self.model_fn(x, s)
        if self.pp:
            phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)
            x_s1 = (sig_s1 / sig_s) * x - (alpha_s1 * phi_11) * model_s
            model_s1 = self.model_fn(x_s1, s1)
            if st == "dpmsolver":
                x_t = (sig_t / sig_s) * x - (alpha_t * phi_1) * model_s - (0.5 / r1) * (alpha_t * phi_1) * (model_s1 - model_s)
            else:
                x_t = (sig_t / sig_s) * x - (alpha_t * phi_1) * model_s + (1.0 / r1) * (alpha_t * (phi_1 / h + 1.0)) * (model_s1 - model_s)
        else:
            phi_11, phi_1 = torch.expm1(r1 * h), torch.expm1(h)
            x_s1 = torch.exp(la_s1 - la_s) * x - (sig_s1 * phi_11) * model_s
            model_s1 = self.model_fn(x_s1, s1)
            if st == "dpmsolver":
                x_t = torch.exp(la_t - la_s) * x - (sig_t * phi_1) * model_s - (0.5 / r1) * (sig_t * phi_1) * (model_s1 - model_s)
            else:
                x_t = torch.exp(la_t - la_s) * x - (sig_t * phi_1) * model_s - (1.0 / r1) * (sig_t * (phi_1 / h - 1.0)) * (model_s1 - model_s)
        return (x_t, {"model_s": model_s, "model_s1": model_s1}) if return_intermediate else x_t

    def third(self, x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0, model_s=None, model_s1=None, return_intermediate=False):
        ns, st = self.ns, self.solver_type
        if r1 is None:
            r1 = 1.0 / 3.0
        if r2 is None:
            r2 = 2.0 / 3.0
        lam_s, lam_t = ns.lam(s), ns.lam(t)
        h = lam_t - lam_s
        s1, s2 = ns.inverse_lambda(lam_s + r1 * h), ns.inverse_lambda(lam_s + r2 * h)
        la_s, la_s1, la_s2, la_t = ns.log_alpha(s), ns.log_alpha(s1), ns.log_alpha(s2), ns.log_alpha(t)
        sig_s, sig_s1, sig_s2, sig_t = ns.sigma(s), ns.sigma(s1), ns.sigma(s2), ns.sigma(t)
        alpha_s1, alpha_s2, alpha_t = torch.exp(la_s1), torch.exp(la_s2), torch.exp(la_t)
        if model_s is None:
            model_s = self.model_fn(x, s)
        if self.pp:
            phi_11, phi_12, phi_1 = torch.expm1(-r1 * h), torch.expm1(-r2 * h), torch.expm1(-h)
            phi_22 = torch.expm1(-r2 * h) / (r2 * h) + 1.0
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            if model_s1 is None:
                x_s1 = (sig_s1 / sig_s) * x - (alpha_s1 * phi_11) * model_s
                model_s1 = self.model_fn(x_s1, s1)
            x_s2 = (sig_s2 / sig_s) * x - (alpha_s2 * phi_12) * model_s + r2 / r1 * (alpha_s2 * phi_22) * (model_s1 - model_s)
            model_s2 = self.model_fn(x_s2, s2)
            if st == "dpmsolver":
                x_t = (sig_t / sig_s) * x - (alpha_t * phi_1) * model_s + (1.0 / r2) * (alpha_t * phi_2) * (model_s2 - model_s)
            else:
                D1_0 = (1.0 / r1) * (model_s1 - model_s)
                D1_1 = (1.0 / r2) * (model_s2 - model_s)
                D1 = (r2 * D1_0 - r1 * D1_1) / (r2 - r1)
                D2 = 2.0 * (D1_1 - D1_0) / (r2 - r1)
                x_t = (sig_t / sig_s) * x - (alpha_t * phi_1) * model_s + (alpha_t * phi_2) * D1 - (alpha_t * phi_3) * D2
        else:
            phi_11, phi_12, phi_1 = torch.expm1(r1 * h), torch.expm1(r2 * h), torch.expm1(h)
            phi_22 = torch.expm1(r2 * h) / (r2 * h) - 1.0
            phi_2 = phi_1 / h - 1.0
            phi_3 = phi_2 / h - 0.5
            if model_s1 is None:
                x_s1 = torch.exp(la_s1 - la_s) * x - (sig_s1 * phi_11) * model_s
                model_s1 = self.model_fn(x_s1, s1)
            x_s2 = torch.exp(la_s2 - la_s) * x - (sig_s2 * phi_12) * model_s - r2 / r1 * (sig_s2 * phi_22) * (model_s1 - model_s)
            model_s2 = self.model_fn(x_s2, s2)
            if st == "dpmsolver":
                x_t = torch.exp(la_t - la_s) * x - (sig_t * phi_1) * model_s - (1.0 / r2) * (sig_t * phi_2) * (model_s2 - model_s)
            else:
                D1_0 = (1.0 / r1) * (model_s1 - model_s)
                D1_1 = (1.0 / r2) * (model_s2 - model_s)
                D1 = (r2 * D1_0 - r1 * D1_1) / (r2 - r1)
                D2 = 2.0 * (D1_1 - D1_0) / (r2 - r1)
                x_t = torch.exp(la_t - la_s) * x - (sig_t * phi_1) * model_s - (sig_t * phi_2) * D1 - (sig_t * phi_3) * D2
        if return_intermediate:
            return x_t, {"model_s": model_s, "model_s1": model_s1, "model_s2": model_s2}
        return x_t


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


def sample_dpm_singlestep(model, x, betas=None, steps=9, order=2, algorithm_type="dpmsolver", model_type="x_start",
                          skip_type="logSNR", method="singlestep", denoise_to_zero=True, solver_type="dpmsolver",
                          return_model_times=False):
    """DPM_Solver.sample(method='singlestep' | 'singlestep_fixed') (sampler.py:1216-1239)."""
    betas = betas_fp32() if betas is None else betas
    ns = NoiseScheduleDiscrete(betas)
    model_times = []
    _, data_pred, model_fn = _dpm_model_fns(model, ns, x.shape[0], algorithm_type, model_type, model_times)
    ss = _Singlestep(ns, model_fn, algorithm_type, solver_type)
    if method == "singlestep":
        orders, K = singlestep_orders(steps, order)
        if skip_type == "logSNR":
            outer = dpm_time_steps(ns, K, skip_type)
        else:
            full = dpm_time_steps(ns, steps, skip_type)
            idx = [0]
            for o in orders:
                idx.append(idx[-1] + o)
            outer = [full[i] for i in idx]
    elif method == "singlestep_fixed":
        K = steps // order
        orders = [order] * K
        outer = dpm_time_steps(ns, K, skip_type)
    else:
        raise ValueError(method)
    for step, o in enumerate(orders):
        s_, t_ = outer[step], outer[step + 1]
        inner = dpm_time_steps(ns, o, skip_type, t_T=float(s_), t_0=float(t_))
        lam_inner = [ns.lam(torch.as_tensor(v, dtype=torch.float32)) for v in inner]
        h = lam_inner[-1] - lam_inner[0]
        r1 = None if o <= 1 else (lam_inner[1] - lam_inner[0]) / h
        r2 = None if o <= 2 else (lam_inner[2] - lam_inner[0]) / h
        if o == 1:
            x = ss.first(x, s_, t_)
        elif o == 2:
            x = ss.second(x, s_, t_, r1=r1)
        else:
            x = ss.third(x, s_, t_, r1=r1, r2=r2)
    if denoise_to_zero:
        x = data_pred(x, torch.ones(()) * (1.0 / ns.total_N))
    return (x, model_times) if return_model_times else x


def sample_dpm_adaptive(model, x, betas=None, order=2, algorithm_type="dpmsolver", model_type="x_start", h_init=0.05,
                        atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5, solver_type="dpmsolver", denoise_to_zero=False,
                        return_model_times=False):
    """DPM_Solver.sample(method='adaptive') -> dpm_solver_adaptive (sampler.py:958-1009,1171-1172)."""
    betas = betas_fp32() if betas is None else betas
    ns = NoiseScheduleDiscrete(betas)
    model_times = []
    _, data_pred, model_fn = _dpm_model_fns(model, ns, x.shape[0], algorithm_type, model_type, model_times)
    ss = _Singlestep(ns, model_fn, algorithm_type, solver_type)
    t_T, t_0 = ns.T, 1.0 / ns.total_N
    s = t_T * torch.ones((1,))
    lam_s = ns.lam(s)
    lam_0 = ns.lam(t_0 * torch.ones_like(s))
    h = h_init * torch.ones_like(s)
    x_prev = x
    nfe = 0
    if order == 2:
        r1 = 0.5
        lower = lambda x_, s_, t_: ss.first(x_, s_, t_, return_intermediate=True)
        higher = lambda x_, s_, t_, **kw: ss.second(x_, s_, t_, r1=r1, **kw)
    elif order == 3:
        r1, r2 = 1.0 / 3.0, 2.0 / 3.0
        lower = lambda x_, s_, t_: ss.second(x_, s_, t_, r1=r1, return_intermediate=True)
        higher = lambda x_, s_, t_, **kw: ss.third(x_, s_, t_, r1=r1, r2=r2, **kw)
    else:
        raise ValueError("For adaptive step size solver, order must be 2 or 3, got {}".format(order))
    while torch.abs(s - t_0).mean() > t_err:
        t = ns.inverse_lambda(lam_s + h)
        x_lower, kw = lower(x, s, t)
        x_higher = higher(x, s, t, **kw)
        delta = torch.max(torch.ones_like(x) * atol, rtol * torch.max(torch.abs(x_lower), torch.abs(x_prev)))
        norm_fn = lambda v: torch.sqrt(torch.square(v.reshape((v.shape[0], -1))).mean(dim=-1, keepdim=True))
        E = norm_fn((x_higher - x_lower) / delta).max()
        if torch.all(E <= 1.0):
            x = x_higher
            s = t
            x_prev = x_lower
            lam_s = ns.lam(s)
        h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lam_0 - lam_s)
        nfe += order
    if denoise_to_zero:
        x = data_pred(x, torch.ones(()) * t_0)
    return (x, model_times, nfe) if return_model_times else x


# ----------------------------------------------------------------------------- post-processing
def inverse_data_transform(x):
    """datasets/__init__.py:26-35 with cfgs/diffusion.yml:1-8 (no rescale, no logit)."""
    return torch.clamp(x, 0.0, 1.0)


def minmax_map(x):
    """Per-clip min-max normalisation to [0,1] (util/utils.py:11-16 before the uint8 cast;
    metrics/utils.py:11-13) -- the space the 1e-2 max-abs tolerance is stated in."""
    flat = x.reshape(x.shape[0], -1)
    lo = flat.min(dim=1, keepdim=True).values
    hi = flat.max(dim=1, keepdim=True).values
    return ((flat - lo) / (hi - lo)).reshape(x.shape)
