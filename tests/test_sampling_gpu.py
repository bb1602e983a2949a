"""GPU: the whole sampling loop (DDIM and DPM-solver, through the reference-compatible API and dsb_sample) against
the golden fixtures generated from the unmodified reference and against the fp32 oracle.

north_star tolerance: per-pixel max-abs <= 1e-2 on min-max-normalised maps; CC / NSS / SIM / AUC-J within 0.5 %."""
import os
import types

import numpy as np
import pytest
import torch

from diff_sal_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-2


def gold(name):
    return torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])


def minmax(x):
    flat = x.reshape(x.shape[0], -1)
    lo = flat.min(dim=1, keepdim=True).values
    hi = flat.max(dim=1, keepdim=True).values
    return ((flat - lo) / (hi - lo)).reshape(x.shape)


def config(sample_type="ddim", timesteps=5, eta=0.0, order=2, target="x0"):
    ns = types.SimpleNamespace
    return ns(training=ns(training_target=target),
              diffusion=ns(beta_schedule="cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
              sampling=ns(sample_type=sample_type, timesteps=timesteps, eta=eta, skip_type="logSNR", dpm_solver_order=order,
                          denoise=True, dpm_solver_method="multistep", dpm_solver_type="dpmsolver", dpm_solver_atol=0.0078,
                          dpm_solver_rtol=0.05, lower_order_final=False, thresholding=False))


@pytest.fixture(scope="module")
def nets():
    from diff_sal_b200.salunet import SalUNetB200
    cache = {}

    def get(kind, audio):
        if (kind, audio) not in cache:
            m = SalUNetB200(max_batch=2, audio_visual=audio, image_based=True, img_size=(224, 384), mid_num_stages=4)
            m.load_state_dict(synth.make_state_dict(kind))
            cache[(kind, audio)] = m
        return cache[(kind, audio)]
    yield get
    for m in cache.values():
        m.engine.close()


def inputs(B, audio):
    x, feats, aud = synth.make_inputs(B, audio=audio)
    return x.cuda(), [f.cuda() for f in feats], (aud.cuda() if aud is not None else None)


@pytest.mark.parametrize("S", [1, 5])
def test_ddim_against_reference_golden(nets, S):
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, aud = inputs(1, True)
    smp = DiffusionSampler(nets("wide", True), config("ddim", S))
    y = smp.sample_image(x, feats, aud).cpu()
    assert (minmax(y) - minmax(gold("ddim%d_wide_av" % S))).abs().max().item() <= TOL


@pytest.mark.parametrize("name,algo,mtype", [("dpm_wide_av_xstart_o2_s4", "dpmsolver", "x_start"),
                                              ("dpmpp_wide_av_xstart_o2_s4", "dpmsolver++", "x_start"),
                                              ("dpm_wide_av_noise_o2_s4", "dpmsolver", "noise")])
def test_dpm_solver_against_reference_golden(nets, name, algo, mtype):
    from diff_sal_b200.sampler import DPM_Solver, NoiseScheduleVP, get_beta_schedule, model_wrapper, to_torch
    x, feats, aud = inputs(1, True)
    betas = to_torch(get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ns = NoiseScheduleVP(schedule="discrete", betas=betas)
    mf = model_wrapper(nets("wide", True), ns, model_type=mtype, model_kwargs={"audio_feat_list": aud}, guidance_type="uncond")
    y = DPM_Solver(mf, ns, algorithm_type=algo).sample(x, feats, steps=4, order=2, skip_type="logSNR", method="multistep",
                                                     lower_order_final=False, denoise_to_zero=True).cpu()
    assert (minmax(y) - minmax(gold(name))).abs().max().item() <= TOL


def test_config1_visual_only_dpm_solver(nets):
    """BASELINE config 1 (visual-only, batch 1, reference init, DPM-solver multistep-2, 10 NFE)."""
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, _ = inputs(1, False)
    smp = DiffusionSampler(nets("ref_init", False), config("dpmsolver", 10))
    y = smp.sample_image(x, feats, None).cpu()
    assert (minmax(y) - minmax(gold("cfg1_dpm_refinit_vis_xstart_o2_s9"))).abs().max().item() <= TOL


def test_graph_replay_equals_eager_and_is_repeatable(nets):
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, aud = inputs(2, True)
    smp = DiffusionSampler(nets("wide", True), config("ddim", 3))
    eager = smp.sample_image(x, feats, aud, use_graph=False).clone()
    g1 = smp.sample_image(x, feats, aud, use_graph=True).clone()     # capture + launch
    g2 = smp.sample_image(x, feats, aud, use_graph=True).clone()     # cached replay
    torch.cuda.synchronize()
    assert torch.equal(eager, g1) and torch.equal(g1, g2)


def test_batch_invariance_per_clip(nets):
    """A clip's map does not depend on what else is in the batch (what makes 1-vs-N-GPU sharding bit-identical)."""
    from diff_sal_b200.sampler import DiffusionSampler
    smp = DiffusionSampler(nets("wide", True), config("ddim", 2))
    x, feats, aud = inputs(2, True)
    both = smp.sample_image(x, feats, aud).clone()
    one = smp.sample_image(x[1:], [f[1:].contiguous() for f in feats], aud[1:].contiguous()).clone()
    torch.cuda.synchronize()
    assert torch.equal(both[1:], one)


def test_generic_callable_path_matches_fused(nets):
    """DPM_Solver driven by an arbitrary callable (one dsb_denoise + one fused update per op) equals dsb_sample."""
    from diff_sal_b200 import sampler as S
    net = nets("wide", True)
    x, feats, aud = inputs(1, True)
    ns = S.NoiseScheduleVP("discrete", betas=S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02,
                                                                              num_diffusion_timesteps=1000)))
    ops, _ = S.build_dpm_program(ns, 3, 2, "dpmsolver++", "noise")
    fused = net._dsb_fused_sample(ops, x, feats, {"audio_feat_list": aud}).clone()
    generic = S.run_program_generic(ops, x.clone(), lambda x_, t_: net(x_, t_, feats, aud))
    torch.cuda.synchronize()
    assert torch.equal(fused, generic)


def test_generalized_steps_api(nets):
    """util/denoising.py generalized_steps: model(data, t) eps-parameterised, returns (xs, x0_preds)."""
    from diff_sal_b200 import sampler as S
    from oracle import salunet, samplers as O
    net = nets("wide", True)
    x, feats, aud = inputs(1, True)
    b = O.betas_fp32()
    seq = range(0, 1000, 500)
    xs, x0s = S.generalized_steps(x, seq, lambda data, t: net(data["input"], t, data["img"], aud), b.cuda(), img=feats)
    assert len(xs) == 3 and len(x0s) == 2
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(1, audio=True)
    ref = O.sample_ddim(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, 2, training_target="noise")
    got = xs[-1].cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_metrics_within_half_percent(nets):
    """CC / NSS / SIM / AUC-J of the CUDA map vs the fp32 oracle map on synthetic ground truth."""
    from diff_sal_b200.sampler import DiffusionSampler
    from oracle import metrics, salunet, samplers as O
    x, feats, aud = inputs(1, True)
    y = DiffusionSampler(nets("wide", True), config("ddim", 2)).sample_image(x, feats, aud).cpu()
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(1, audio=True)
    ref = O.sample_ddim(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, 2)
    ref_map = O.inverse_data_transform(ref)[0, 0].double().numpy()
    gt = metrics.ground_truth_from_map(ref_map, 0)      # GT correlated with the reference prediction
    a = metrics.all_metrics(O.inverse_data_transform(y)[0, 0].double().numpy(), 0, gt=gt)
    b = metrics.all_metrics(ref_map, 0, gt=gt)
    print("metrics cuda", a, "oracle", b)
    for k in a:
        assert abs(b[k]) > 0.1
        assert abs(a[k] - b[k]) <= 5e-3 * abs(b[k]), (k, a[k], b[k])


def test_output_postprocessing_matches_reference_formulas():
    """inverse_data_transform (clamp) and normalize_data (min-max -> uint8) on the GPU vs the numpy formulas of
    datasets/__init__.py:35 and util/utils.py:11-16."""
    from diff_sal_b200 import sampler as S
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(3, 1, 224, 384, generator=g) * 1.4 - 0.2).cuda()
    cfg = types.SimpleNamespace(data=types.SimpleNamespace(logit_transform=False, rescaled=False))
    c = S.inverse_data_transform(cfg, x)
    assert torch.equal(c, torch.clamp(x, 0.0, 1.0))
    u8 = S.normalize_data(x).cpu().numpy()
    for i in range(3):
        d = torch.clamp(x[i], 0.0, 1.0).cpu().numpy()
        ref = np.clip((d - d.min()) * (255.0 / (d.max() - d.min())), 0, 255).astype(np.uint8)
        diff = np.abs(u8[i].astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


# ------------------------------------------------------------------------------------------ SURVEY 8f row N4
def test_ddpm_ancestral_matches_oracle(nets):
    """DiffusionTrainer.sample_ddpm / p_sample (diffusion_trainer.py:488-540) with fixed ancestral noise."""
    from diff_sal_b200.sampler import DiffusionSampler
    from oracle import salunet, samplers as O
    x, feats, aud = inputs(1, True)
    # timesteps 3 -> seq = range(0, 1000, 333) = [0, 333, 666, 999]: 4 evaluations, noise on the 3 with t > 0
    noise = torch.randn(3, 1, 1, 224, 384, generator=torch.Generator().manual_seed(11))
    smp = DiffusionSampler(nets("wide", True), config("ddpm", 3))
    y = smp.sample_ddpm(x, feats, aud, noise=noise.cuda()).cpu()
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(1, audio=True)
    it = iter(noise)
    ref = O.sample_ddpm(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, 3, "x0", noise_fn=lambda x_: next(it))
    assert (minmax(y) - minmax(ref)).abs().max().item() <= TOL
    # sample_image dispatch (sample_type 'ddpm') draws its own noise: same shape, finite, graph == eager is covered above
    y2 = smp.sample_image(x, feats, aud)
    assert y2.shape == x.shape and torch.isfinite(y2).all()


def test_dynamic_thresholding_kernel_is_torch_quantile():
    """dsb_sampler_dynamic_threshold vs DPM_Solver.dynamic_thresholding_fn (sampler.py:417-426) on random maps."""
    from diff_sal_b200 import sampler as S
    from oracle import samplers as O
    g = torch.Generator().manual_seed(3)
    for n, scale, mx in [(224 * 384, 3.0, 1.0), (224 * 384, 0.2, 1.0), (4096, 5.0, 0.5)]:
        x = scale * torch.randn(3, 1, n // 64, 64, generator=g)
        x[1, 0, 0, :8] = x[1, 0, 1, :8]                       # a few exact ties
        k, w = S.quantile_rank(0.995, n)
        got = S._correct("thresh", x.cuda().clone(), (k, w, mx)).cpu()
        ref = O.dynamic_thresholding(x, 0.995, mx)
        assert (got - ref).abs().max().item() <= 1e-6


def test_dpm_solver_pp_dynamic_thresholding(nets):
    """correcting_x0_fn='dynamic_thresholding' in the fused loop vs the oracle loop (max_val 0.2 so that the
    99.5 % quantile of the saliency map, not the floor, sets the scale)."""
    from diff_sal_b200 import sampler as S
    from oracle import salunet, samplers as O
    net = nets("wide", True)
    x, feats, aud = inputs(1, True)
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    mf = S.model_wrapper(net, ns, model_type="x_start", model_kwargs={"audio_feat_list": aud}, guidance_type="uncond")
    y = S.DPM_Solver(mf, ns, algorithm_type="dpmsolver++", correcting_x0_fn="dynamic_thresholding",
                     thresholding_max_val=0.2).sample(x, feats, steps=2, order=2, skip_type="logSNR", method="multistep",
                                                      lower_order_final=False, denoise_to_zero=True).cpu()
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(1, audio=True)
    ref = O.sample_dpm(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, steps=2, order=2, algorithm_type="dpmsolver++",
                       model_type="x_start", correcting_x0_fn="dynamic_thresholding", thresholding_max_val=0.2)
    assert ref.max().item() > 0.99                                  # the scale was active
    # every data prediction is divided by s = max(q_99.5, 0.2) ~ 0.3 before it re-enters the solver, so the
    # denoiser's bf16-operand error (2e-3 ... 5e-3 on the un-thresholded loops above) is multiplied by ~3.3 per step:
    # the stated 1e-2 applies to the reference's shipped configuration (thresholding off), this variant gets 3x
    assert (minmax(y) - minmax(ref)).abs().max().item() <= 3 * TOL
    # shipped setting (thresholding_max_val = 1.0): saliency maps live in (0, 1), s = 1 and the op is a clamp only
    y1 = S.DPM_Solver(mf, ns, algorithm_type="dpmsolver++", correcting_x0_fn="dynamic_thresholding").sample(
        x, feats, steps=2, order=2, skip_type="logSNR", method="multistep", lower_order_final=False, denoise_to_zero=True).cpu()
    ref1 = O.sample_dpm(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, steps=2, order=2, algorithm_type="dpmsolver++",
                        model_type="x_start", correcting_x0_fn="dynamic_thresholding")
    assert (minmax(y1) - minmax(ref1)).abs().max().item() <= TOL


def test_ddpm_steps_api():
    """util/denoising.py:39-67 around an arbitrary eps-model (toy network on the GPU), fixed noise."""
    from diff_sal_b200 import sampler as S
    from oracle import samplers as O
    toy = lambda x_, t_: torch.tanh(0.7 * x_ + 0.001 * t_.float()[:, None, None, None]) * 0.5 + 0.1 * torch.roll(x_, 1, -1)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 1, 16, 16, generator=g)
    noise = [torch.randn(2, 1, 16, 16, generator=g) for _ in range(4)]
    b = O.betas_fp32()
    it = iter(noise)
    rxs, rx0 = O.ddpm_steps(x, range(0, 1000, 250), toy, b, noise_fn=lambda x_: next(it))
    xs, x0s = S.ddpm_steps(x.cuda(), range(0, 1000, 250), toy, b, noise=[z.cuda() for z in noise])
    assert len(xs) == 5 and len(x0s) == 4
    for a, r in zip(xs + x0s, rxs + rx0):
        assert (a.cpu() - r).abs().max().item() <= 1e-4 * max(1.0, r.abs().max().item())


@pytest.mark.parametrize("algo,order,steps", [("dpmsolver", 2, 4), ("dpmsolver++", 3, 6)])
def test_dpm_singlestep_with_conditioning(nets, algo, order, steps):
    """DPM_Solver.sample(method='singlestep') (sampler.py:573-795,1216-1239) in the fused loop, with the
    conditioning forwarded to every (also the intermediate) network evaluation, against the fp32 oracle."""
    from diff_sal_b200 import sampler as S
    from oracle import salunet, samplers as O
    net = nets("wide", True)
    x, feats, aud = inputs(1, True)
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    mf = S.model_wrapper(net, ns, model_type="x_start", model_kwargs={"audio_feat_list": aud}, guidance_type="uncond")
    y = S.DPM_Solver(mf, ns, algorithm_type=algo).sample(x, feats, steps=steps, order=order, skip_type="logSNR",
                                                        method="singlestep", denoise_to_zero=True).cpu()
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(1, audio=True)
    ref = O.sample_dpm_singlestep(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, steps=steps, order=order,
                                  algorithm_type=algo, model_type="x_start", skip_type="logSNR", method="singlestep")
    assert (minmax(y) - minmax(ref)).abs().max().item() <= TOL
    # graph replay == eager for programs with intermediate evaluations
    y2 = S.DPM_Solver(mf, ns, algorithm_type=algo).sample(x, feats, steps=steps, order=order, skip_type="logSNR",
                                                         method="singlestep", denoise_to_zero=True, use_graph=False).cpu()
    assert torch.equal(y, y2)


@pytest.mark.parametrize("algo,order", [("dpmsolver", 2), ("dpmsolver++", 3)])
def test_dpm_adaptive_on_toy_net(algo, order):
    """DPM_Solver.sample(method='adaptive') (sampler.py:958-1009): host-driven step-size control over the fused update
    kernels and the device error norm; same accepted / rejected steps (NFE) and the same iterate as the fp32 oracle."""
    from diff_sal_b200 import sampler as S
    from oracle import samplers as O
    toy = lambda x_, t_: torch.tanh(0.7 * x_ + 0.001 * t_.float()[:, None, None, None]) * 0.5 + 0.1 * torch.roll(x_, 1, -1)
    x = torch.randn(2, 1, 16, 24, generator=torch.Generator().manual_seed(6))
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    mf = S.model_wrapper(lambda x_, t_, img, **kw: toy(x_, t_), ns, model_type="noise", model_kwargs={}, guidance_type="uncond")
    y = S.DPM_Solver(mf, ns, algorithm_type=algo).sample(x.cuda(), None, order=order, method="adaptive",
                                                        denoise_to_zero=False).cpu()
    ref, times, nfe = O.sample_dpm_adaptive(toy, x, order=order, algorithm_type=algo, model_type="noise", return_model_times=True)
    _, got_nfe = S.sample_dpm_adaptive(ns, x.cuda(), lambda x_, t_: toy(x_, t_), order, algo, "noise")
    assert got_nfe == nfe
    assert (y - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_dpm_adaptive_with_the_conditioned_denoiser(nets):
    """The adaptive solver around SalUNetB200 (the reference's own adaptive branch cannot run a conditioned SalUNet: it
    never passes ``img``): finite, in range, and -- with an x0 network -- close to the multistep solution."""
    from diff_sal_b200 import sampler as S
    from oracle import samplers as O
    net = nets("wide", True)
    x, feats, aud = inputs(1, True)
    ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
    mf = S.model_wrapper(net, ns, model_type="x_start", model_kwargs={"audio_feat_list": aud}, guidance_type="uncond")
    solver = S.DPM_Solver(mf, ns, algorithm_type="dpmsolver++")
    ya = solver.sample(x, feats, order=2, method="adaptive", denoise_to_zero=True, atol=0.05, rtol=0.2).cpu()
    ym = solver.sample(x, feats, steps=9, order=2, skip_type="logSNR", method="multistep", denoise_to_zero=True).cpu()
    assert torch.isfinite(ya).all() and ya.min().item() >= 0.0 and ya.max().item() <= 1.0
    assert (minmax(ya) - minmax(ym)).abs().max().item() <= 5e-2
