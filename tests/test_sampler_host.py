"""CPU: host-side sampler logic of the product (schedules, step coefficients, sampler programs) against the
oracle, plus the C-ABI export check.  No compute call touches the GPU here."""
import ctypes
import subprocess
import os
import re

import numpy as np
import pytest
import torch

from diff_sal_b200 import sampler as S
from oracle import samplers as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy(x, t):
    return torch.tanh(0.7 * x + 0.001 * t.float()[:, None, None, None]) * 0.5 + 0.1 * torch.roll(x, 1, -1)


def interpret(ops, x, net, noise=None):
    """Test-side reference interpreter of a sampler program (what dsb_sample does on the GPU)."""
    bufs = {0: x.double()}
    for op in ops:
        if op[0] == "eval":
            t = torch.full((x.shape[0],), op[1], dtype=torch.float32)
            bufs[1] = net(bufs[op[2] if len(op) > 2 else 0].float(), t).double()
        elif op[0] == "clamp":
            bufs[op[1]] = bufs[op[1]].clamp(op[2], op[3])
        elif op[0] == "thresh":
            _, dst, k, w, mx = op
            v = bufs[dst].float()
            srt = v.abs().reshape(v.shape[0], -1).sort(dim=1).values
            a, b = srt[:, k], srt[:, min(k + 1, srt.shape[1] - 1)]
            q = torch.lerp(a, b, torch.tensor(w))
            sc = torch.maximum(q, torch.tensor(mx)).reshape((-1,) + (1,) * (v.dim() - 1))
            bufs[dst] = (torch.clamp(v, -sc, sc) / sc).double()
        else:
            _, dst, terms, ncoef, nidx = op
            acc = sum(c * bufs[s] for s, c in terms)
            if nidx >= 0:
                acc = acc + ncoef * noise[nidx].double()
            bufs[dst] = acc
    return bufs[0].float()


def test_betas_match_oracle():
    b = S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    assert np.array_equal(b, O.cosine_betas(1000))


def test_noise_schedule_matches_oracle():
    betas = O.betas_fp32()
    ns, ref = S.NoiseScheduleVP("discrete", betas=betas), O.NoiseScheduleDiscrete(betas)
    assert ns.total_N == ref.total_N == 996           # cosine schedule clipped at lambda = -5.1
    for t in [1.0, 0.73, 0.5, 0.11, 1.0 / 996, 0.0123]:
        tt = torch.tensor(t)
        assert abs(ns.marginal_log_mean_coeff(t) - ref.log_alpha(tt).item()) < 2e-6 * max(1, abs(ref.log_alpha(tt).item()))
        assert abs(ns.marginal_std(t) - ref.sigma(tt).item()) < 1e-5
        lam = ref.lam(tt)
        assert abs(ns.marginal_lambda(t) - lam.item()) < 2e-4
        assert abs(ns.inverse_lambda(lam.item()) - ref.inverse_lambda(lam).item()) < 1e-5


def test_model_times_10_nfe():
    """model-time inputs for 10 NFE (steps=9 + denoise_to_zero)."""
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    _, times = S.build_dpm_program(ns, 9, 2, "dpmsolver", "x_start")
    want = [999.0, 990.6, 965.0, 886.9, 673.6, 329.0, 110.9, 31.0, 6.0, 0.0]
    assert len(times) == 10
    for a, b in zip(times, want):
        assert abs(a - b) < 0.06, (times, want)


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("mtype", ["x_start", "noise"])
@pytest.mark.parametrize("order,steps,lof", [(2, 9, False), (1, 4, False), (3, 7, False), (2, 5, True), (3, 6, True),
                                              (1, 1, False), (2, 2, False), (2, 24, False)])
def test_dpm_program_matches_oracle(algo, mtype, order, steps, lof):
    betas = O.betas_fp32()
    x = torch.randn(1, 1, 8, 8, generator=torch.Generator().manual_seed(3))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    ops, _ = S.build_dpm_program(ns, steps, order, algo, mtype, "logSNR", lof, True)
    got = interpret(ops, x, _toy)
    ref = O.sample_dpm(_toy, x, betas, steps=steps, order=order, algorithm_type=algo, model_type=mtype,
                       lower_order_final=lof)
    tol = 2e-4 * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() < tol


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("mtype", ["x_start", "noise"])
@pytest.mark.parametrize("method,order,steps,skip,solver_type", [
    ("singlestep", 1, 4, "time_uniform", "dpmsolver"), ("singlestep", 2, 6, "logSNR", "dpmsolver"),
    ("singlestep", 2, 7, "logSNR", "taylor"), ("singlestep", 3, 9, "logSNR", "dpmsolver"),
    ("singlestep", 3, 10, "time_uniform", "dpmsolver"), ("singlestep", 3, 11, "logSNR", "taylor"),
    ("singlestep", 2, 5, "time_quadratic", "dpmsolver"), ("singlestep_fixed", 3, 9, "logSNR", "dpmsolver")])
def test_dpm_singlestep_program_matches_oracle(algo, mtype, method, order, steps, skip, solver_type):
    """the singlestep solvers as EVAL/AXPY programs (intermediate evaluations read buffer 5)."""
    betas = O.betas_fp32()
    x = torch.randn(1, 1, 8, 8, generator=torch.Generator().manual_seed(13))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    ops, times = S.build_dpm_singlestep_program(ns, steps, order, algo, mtype, skip, method, True, solver_type)
    ref, rtimes = O.sample_dpm_singlestep(_toy, x, betas, steps=steps, order=order, algorithm_type=algo, model_type=mtype,
                                          skip_type=skip, method=method, denoise_to_zero=True, solver_type=solver_type,
                                          return_model_times=True)
    assert len(times) == len(rtimes)
    for a, b in zip(times, rtimes):
        assert abs(a - b) < 0.06
    tol = 2e-4 * max(1.0, ref.abs().max().item())
    assert (interpret(ops, x, _toy) - ref).abs().max().item() < tol


@pytest.mark.parametrize("solver_type", ["dpmsolver", "taylor"])
def test_dpm_program_taylor(solver_type):
    betas = O.betas_fp32()
    x = torch.randn(1, 1, 8, 8, generator=torch.Generator().manual_seed(4))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    for algo in ("dpmsolver", "dpmsolver++"):
        ops, _ = S.build_dpm_program(ns, 6, 2, algo, "x_start", solver_type=solver_type)
        ref = O.sample_dpm(_toy, x, betas, steps=6, order=2, algorithm_type=algo, model_type="x_start",
                           solver_type=solver_type)
        assert (interpret(ops, x, _toy) - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("target", ["x0", "noise"])
@pytest.mark.parametrize("Sn,eta", [(1, 0.0), (5, 0.0), (10, 0.5), (25, 0.0)])
def test_ddim_program_matches_oracle(target, Sn, eta):
    tables = S.DdimTables(O.betas_fp32())
    x = torch.randn(2, 1, 8, 8, generator=torch.Generator().manual_seed(5))
    ops, n_noise = S.build_ddim_program(tables, Sn, eta, target)
    g = torch.Generator().manual_seed(6)
    noise = [torch.randn(x.shape, generator=g) for _ in range(max(n_noise, Sn))]
    it = iter(noise)
    ref = O.sample_ddim(_toy, x, Sn, eta=eta, training_target=target, noise_fn=lambda x_: next(it))
    got = interpret(ops, x, _toy, noise)
    assert (got - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_unsupported_paths_fail_loudly():
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    with pytest.raises(NotImplementedError):
        S.model_wrapper(lambda *a: None, ns, guidance_type="classifier")
    mf = S.model_wrapper(lambda *a: None, ns, model_type="noise")
    with pytest.raises(ValueError):
        S.DPM_Solver(mf, ns).sample(torch.zeros(1, 1, 8, 8), method="no_such_method")
    with pytest.raises(NotImplementedError):
        S.DPM_Solver(mf, ns).sample(torch.zeros(1, 1, 8, 8), method="singlestep", return_intermediate=True)
    with pytest.raises(IndexError):
        # the reference's singlestep order 1 with logSNR spacing indexes past its K = 1 outer steps (sampler.py:536-540,1224)
        S.build_dpm_singlestep_program(ns, 4, 1, skip_type="logSNR")
    with pytest.raises(AssertionError):
        S.build_dpm_program(ns, 1, 2)           # steps >= order (sampler.py:1174)
    with pytest.raises(RuntimeError):
        S._axpy([1.0], [torch.zeros(4)])        # CPU tensors: no fallback


def test_c_abi_exports_every_declared_symbol():
    from diff_sal_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from diff_sal_b200 import build
        build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "diffsal_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dsb_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libdiffsal_b200.so does not export %s" % n
    assert not any(n.startswith("dsb_test_") for n in names)
    # the per-kernel test entries live in their own library and are not exported by the product library
    tlib = ctypes.CDLL(_lib.TEST_LIB_PATH)
    thdr = open(os.path.join(ROOT, "include", "diffsal_b200_test.h")).read()
    tnames = sorted(set(re.findall(r"\b(dsb_test_[a-z_0-9]+)\s*\(", thdr)))
    assert len(tnames) >= 15
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for n in tnames:
        assert hasattr(tlib, n), "libdiffsal_b200_test.so does not export %s" % n
        assert (" T " + n) not in out, "product library exports the test entry %s" % n


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diff_sal_b200.engine import DsbError, Engine
    with pytest.raises(DsbError):
        Engine(max_batch=1)


@pytest.mark.parametrize("target", ["x0", "noise"])
@pytest.mark.parametrize("Sn", [1, 4, 10])
def test_ddpm_program_matches_oracle(target, Sn):
    """Ancestral sampling (diffusion_trainer.py:488-540) as a program vs the oracle loop, same noise."""
    betas = O.betas_fp32()
    tables = S.DdimTables(betas)
    ops, n_noise = S.build_ddpm_program(tables, Sn, target)
    assert n_noise == Sn - 1
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1, 8, 12, generator=g)
    noise = torch.randn(max(n_noise, 1), 2, 1, 8, 12, generator=g)
    it = iter(noise)
    ref = O.sample_ddpm(lambda x_, t_: _toy(x_, t_), x, Sn, target, noise_fn=lambda x_: next(it))
    got = interpret(ops, x, lambda x_, t_: _toy(x_, t_), noise)
    assert (got - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("algo", ["dpmsolver++", "dpmsolver"])
def test_dynamic_thresholding_program_matches_oracle(algo):
    """correcting_x0_fn='dynamic_thresholding' (sampler.py:410-426,441-442): program ops vs the oracle loop."""
    betas = O.betas_fp32()
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    g = torch.Generator().manual_seed(5)
    x = 2.5 * torch.randn(2, 1, 16, 24, generator=g)
    net = lambda x_, t_: 1.7 * _toy(x_, t_)
    ops, _ = S.build_dpm_program(ns, 5, 2, algo, "x_start", "logSNR", False, True, thresholding=(0.995, 1.0),
                                 map_elems=16 * 24)
    assert sum(op[0] == "thresh" for op in ops) == (6 if algo == "dpmsolver++" else 1)
    ref = O.sample_dpm(net, x, betas, steps=5, order=2, algorithm_type=algo, model_type="x_start",
                       correcting_x0_fn="dynamic_thresholding")
    got = interpret(ops, x, net)
    assert (got - ref).abs().max().item() < 5e-5


def test_quantile_rank_is_torch_rule():
    for n, p in [(86016, 0.995), (384, 0.995), (1000, 0.5), (11, 0.9)]:
        k, w = S.quantile_rank(p, n)
        v = torch.randn(3, n, generator=torch.Generator().manual_seed(n)).abs()
        srt = v.sort(dim=1).values
        want = torch.quantile(v, p, dim=1)
        got = torch.lerp(srt[:, k], srt[:, min(k + 1, n - 1)], torch.tensor(w))
        assert torch.equal(got, want), (n, p)
