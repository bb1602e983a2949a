"""GPU parity on the configuration bench.py measures (BASELINE.json configs[1]) and on longer loops.

Audio-visual, batch 8, DPM-solver multistep order 2, logSNR steps, steps=9 + denoise_to_zero (10 NFE), x0-parameterised
SalUNet -- against golden maps produced by the UNMODIFIED reference (tests/golden/make_golden_cfg2.py runs
models/dpm_solver/sampler.py:1048-1247 around the reference SalUNet on the same seeded inputs / weights).

north_star tolerance, checked PER CLIP: max-abs <= 1e-2 on min-max-normalised maps; CC / NSS / SIM / AUC-J
(metrics/metrics.py) within 0.5 % relative of the reference map's, on synthetic ground truth.
"""
import os
import types

import numpy as np
import pytest
import torch

from diff_sal_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-2          # per-pixel max-abs on min-max-normalised maps (north_star)
MTOL = 5e-3         # relative tolerance of CC / NSS / SIM / AUC-J (north_star)


def gold(name):
    return torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])


def minmax(x):
    flat = x.reshape(x.shape[0], -1)
    lo = flat.min(dim=1, keepdim=True).values
    hi = flat.max(dim=1, keepdim=True).values
    return ((flat - lo) / (hi - lo)).reshape(x.shape)


def config(sample_type, timesteps):
    ns = types.SimpleNamespace
    return ns(training=ns(training_target="x0"),
              diffusion=ns(beta_schedule="cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
              sampling=ns(sample_type=sample_type, timesteps=timesteps, eta=0.0, skip_type="logSNR", dpm_solver_order=2,
                          denoise=True, dpm_solver_method="multistep", dpm_solver_type="dpmsolver", dpm_solver_atol=0.0078,
                          dpm_solver_rtol=0.05, lower_order_final=False, thresholding=False))


@pytest.fixture(scope="module")
def nets():
    from diff_sal_b200.salunet import SalUNetB200
    cache = {}

    def get(kind):
        if kind not in cache:
            m = SalUNetB200(max_batch=8, audio_visual=True)
            m.load_state_dict(synth.make_state_dict(kind))
            cache[kind] = m
        return cache[kind]
    yield get
    for m in cache.values():
        m.engine.close()


def inputs(B):
    x, feats, aud = synth.make_inputs(B, audio=True)
    return x.cuda(), [f.cuda() for f in feats], aud.cuda()


def per_clip_checks(y, ref, tol=TOL, mtol=MTOL, label=""):
    """y, ref: [B,1,H,W] CPU.  Returns the worst per-clip min-max error and the worst relative metric deviation."""
    from oracle import metrics, samplers as O
    err = (minmax(y) - minmax(ref)).abs().reshape(y.shape[0], -1).max(dim=1).values
    worst_m = 0.0
    for b in range(y.shape[0]):
        ref_map = O.inverse_data_transform(ref[b:b + 1])[0, 0].double().numpy()
        got_map = O.inverse_data_transform(y[b:b + 1])[0, 0].double().numpy()
        gt = metrics.ground_truth_from_map(ref_map, b)          # GT correlated with the reference prediction
        a = metrics.all_metrics(got_map, b, gt=gt)
        r = metrics.all_metrics(ref_map, b, gt=gt)
        for k in a:
            assert abs(r[k]) > 0.05, (k, r[k])
            dev = abs(a[k] - r[k]) / abs(r[k])
            worst_m = max(worst_m, dev)
            assert dev <= mtol, "%s clip %d %s: %.6f vs reference %.6f (%.3f %%)" % (label, b, k, a[k], r[k], 100 * dev)
    print("%s per-clip min-max err %s ; worst metric deviation %.4f %%" % (label, ["%.2e" % e for e in err.tolist()], 100 * worst_m))
    assert err.max().item() <= tol, "%s: per-clip min-max errors %s" % (label, err.tolist())
    return err.max().item(), worst_m


@pytest.mark.parametrize("kind", ["wide", "refinit"])
def test_config2_av_dpm10_b8(nets, kind):
    """The bench configuration, both weight sets, every clip of the batch of 8, against the reference's maps."""
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, aud = inputs(8)
    smp = DiffusionSampler(nets("wide" if kind == "wide" else "ref_init"), config("dpmsolver", 10))
    y = smp.sample_image(x, feats, aud).cpu()
    per_clip_checks(y, gold("cfg2_dpm_%s_av_b8_s9" % kind), label="cfg2/" + kind)


def test_config2_against_oracle_loop(nets):
    """Same loop, clips 0-1, against the fp32 CPU oracle executed here (independent of the committed fixture)."""
    from diff_sal_b200.sampler import DiffusionSampler
    from oracle import salunet, samplers as O
    x, feats, aud = inputs(8)
    y = DiffusionSampler(nets("wide"), config("dpmsolver", 10)).sample_image(x, feats, aud).cpu()
    sd = synth.make_state_dict("wide")
    xc, fc, ac = synth.make_inputs(2, audio=True)
    ref = O.sample_dpm(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, steps=9, order=2, algorithm_type="dpmsolver",
                       model_type="x_start")
    per_clip_checks(y[:2], ref, label="cfg2/oracle")
    # the oracle itself reproduces the reference fixture (also checked on CPU in test_oracle_golden.py)
    assert (ref - gold("cfg2_dpm_wide_av_b8_s9")[:2]).abs().max().item() <= 1e-4


def test_config2_engine_path_bench_uses(nets):
    """bench.py drives Engine.set_condition / Engine.sample with the program of sampler.build_dpm_program directly:
    same maps as the DiffusionSampler API (bitwise), so the fixture above covers what is timed."""
    from diff_sal_b200 import sampler as S
    from diff_sal_b200.sampler import DiffusionSampler
    net = nets("wide")
    x, feats, aud = inputs(8)
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ops, times = S.build_dpm_program(S.NoiseScheduleVP("discrete", betas=betas), 9, 2, "dpmsolver", "x_start", "logSNR", False, True)
    assert len(times) == 10
    net.engine.set_condition(feats, aud)
    a = net.engine.sample(ops, x.clone(), use_graph=True).clone()
    net._weights_changed()                                   # drop the module's condition cache (the engine was driven directly)
    b = DiffusionSampler(net, config("dpmsolver", 10)).sample_image(x, feats, aud)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert (minmax(a.cpu()) - minmax(gold("cfg2_dpm_wide_av_b8_s9"))).abs().max().item() <= TOL


def test_dpm_25_nfe(nets):
    """25 NFE (steps=24 + denoise_to_zero), batch 2: error must not grow with the loop length."""
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, aud = inputs(2)
    y = DiffusionSampler(nets("wide"), config("dpmsolver", 25)).sample_image(x, feats, aud).cpu()
    per_clip_checks(y, gold("dpm_wide_av_b2_s24"), label="dpm25")


@pytest.mark.parametrize("S", [10, 25])
def test_ddim_long_loops(nets, S):
    """The trainer's default sampler (sample_ddim, diffusion_trainer.py:439-480) at 10 and 25 steps, batch 2."""
    from diff_sal_b200.sampler import DiffusionSampler
    x, feats, aud = inputs(2)
    y = DiffusionSampler(nets("wide"), config("ddim", S)).sample_image(x, feats, aud).cpu()
    per_clip_checks(y, gold("ddim_wide_av_b2_s%d" % S), label="ddim%d" % S)


def test_config2_noise_parameterisation_true_batch(nets):
    """model_type="noise" -- what the trainer literally passes (diffusion_trainer.py:601) -- is the one DPM-solver mode
    the unmodified reference runs as a true batch of 8 (its x_start conversion, sampler.py:290-292, does not broadcast
    beyond batch 1).  With an x0-trained sigmoid network read as noise the iterates grow to +-800, i.e. the loop amplifies
    every perturbation; the map is still compared per clip after min-max normalisation, with the tolerance widened to
    what that amplification leaves (documented, not the north_star bar, which is stated for the x_start loop)."""
    from diff_sal_b200 import sampler as S
    net = nets("wide")
    x, feats, aud = inputs(8)
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    mf = S.model_wrapper(net, ns, model_type="noise", model_kwargs={"audio_feat_list": aud}, guidance_type="uncond")
    y = S.DPM_Solver(mf, ns, algorithm_type="dpmsolver").sample(x, feats, steps=9, order=2, skip_type="logSNR",
                                                               method="multistep", lower_order_final=False,
                                                               denoise_to_zero=True).cpu()
    ref = gold("cfg2_dpm_wide_av_b8_s9_noise")
    err = (minmax(y) - minmax(ref)).abs().reshape(8, -1).max(dim=1).values
    print("cfg2/noise per-clip min-max err", ["%.2e" % e for e in err.tolist()])
    assert err.max().item() <= 3 * TOL


def test_fused_postprocess_returns_uint8(nets):
    """DSB_OP_POSTPROCESS at the end of the program: the loop hands back inverse_data_transform'ed
    maps and normalize_data's uint8 maps without a separate pass over host memory."""
    from diff_sal_b200 import sampler as S
    net = nets("wide")
    x, feats, aud = inputs(2)
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ops, _ = S.build_dpm_program(S.NoiseScheduleVP("discrete", betas=betas), 3, 2, "dpmsolver", "x_start", "logSNR", False, True)
    net.engine.set_condition(feats, aud)
    plain = net.engine.sample(ops, x.clone(), use_graph=True).clone()
    u8 = torch.empty((2, 1, 224, 384), dtype=torch.uint8, device="cuda")
    post = net.engine.sample(ops + [("post", 0)], x.clone(), use_graph=True, out_u8=u8).clone()
    torch.cuda.synchronize()
    net._weights_changed()
    assert torch.equal(post, plain.clamp(0.0, 1.0))
    assert torch.equal(u8, S.normalize_data(plain))
