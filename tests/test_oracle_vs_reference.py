"""CPU, build container only: the oracle restatement against the imported, unmodified
reference (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import types

import numpy as np
import pytest
import torch

from diff_sal_b200 import synth
from oracle import metrics, ref_loader, salunet, samplers

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _toy(x, t):
    return torch.tanh(0.7 * x + 0.001 * t.float()[:, None, None, None]) * 0.5 + 0.1 * torch.roll(x, 1, -1)


def test_state_dict_spec_matches_reference():
    ref = ref_loader.build_salunet().state_dict()
    spec = synth.state_dict_spec()
    assert [k for k, _ in spec] == list(ref.keys())
    for k, s in spec:
        assert tuple(ref[k].shape) == tuple(s), k


def test_betas_bit_exact():
    ns = ref_loader.load()
    rb = ns.to_torch(ns.get_beta_schedule(beta_schedule="cosine", beta_start=1e-4, beta_end=0.02,
                                          num_diffusion_timesteps=1000))
    assert torch.equal(rb, samplers.betas_fp32())


@pytest.mark.parametrize("kind", ["wide", "ref_init"])
@pytest.mark.parametrize("audio", [True, False])
def test_forward_batch2(kind, audio):
    m = ref_loader.build_salunet()
    sd = synth.make_state_dict(kind)
    m.load_state_dict(sd, strict=True)
    x, feats, aud = synth.make_inputs(2, audio=audio)
    t = torch.tensor([37, 812])
    with torch.no_grad():
        ref = m(x, t, [f.clone() for f in feats], aud)
    assert (salunet.forward(sd, x, t, feats, aud) - ref).abs().max().item() < 5e-6


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("mtype", ["x_start", "noise"])
@pytest.mark.parametrize("order,steps,lof", [(2, 9, False), (1, 4, False), (3, 7, False), (2, 5, True),
                                              (3, 6, True), (1, 1, False), (2, 24, False)])
def test_dpm_solver_bit_exact_on_toy_net(algo, mtype, order, steps, lof):
    ns = ref_loader.load()
    betas = samplers.betas_fp32()
    # the reference's x_start wrapper only broadcasts for batch 1 (sampler.py:290-292)
    x = torch.randn(1 if mtype == "x_start" else 2, 1, 8, 8, generator=torch.Generator().manual_seed(1))
    nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
    mf = ns.model_wrapper(lambda x_, t_, img, **kw: _toy(x_, t_), nsv, model_type=mtype, model_kwargs={},
                          guidance_type="uncond")
    ref = ns.DPM_Solver(mf, nsv, algorithm_type=algo).sample(
        x, None, steps=steps, order=order, skip_type="logSNR", method="multistep",
        lower_order_final=lof, denoise_to_zero=True)
    mine = samplers.sample_dpm(_toy, x, betas, steps=steps, order=order, algorithm_type=algo,
                               model_type=mtype, lower_order_final=lof)
    assert torch.equal(ref, mine)


@pytest.mark.parametrize("target", ["x0", "noise"])
@pytest.mark.parametrize("S,eta", [(1, 0.0), (5, 0.0), (10, 0.5), (25, 0.0)])
def test_ddim_bit_exact_on_toy_net(target, S, eta):
    ref_loader.load()
    import diffusion_trainer as dt

    class Dec(torch.nn.Module):
        def forward(self, x, t, img, audio=None):
            return _toy(x, t)

    tb = samplers.DdimTables()
    tr = dt.DiffusionTrainer.__new__(dt.DiffusionTrainer)
    tr.device = torch.device("cpu")
    tr.num_timesteps = 1000
    tr.training_target = target
    for k in ("alphas_hat", "sqrt_alphas_hat", "sqrt_recip_alphas_hat", "sqrt_recipm1_alphas_hat"):
        setattr(tr, k, getattr(tb, k))
    tr.config = types.SimpleNamespace(sampling=types.SimpleNamespace(timesteps=S, eta=eta))
    tr.model = types.SimpleNamespace(module=types.SimpleNamespace(decoder_net=Dec()))
    x = torch.randn(2, 1, 8, 8, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(5)
    ref = tr.sample_ddim(x, [torch.zeros(1)], None)
    torch.manual_seed(5)
    mine = samplers.sample_ddim(_toy, x, S, eta=eta, training_target=target)
    assert torch.equal(ref, mine)


def test_metrics_match_reference():
    rm = ref_loader.load().metrics
    rng = np.random.RandomState(0)
    for i in range(2):
        dens, fix = metrics.synthetic_ground_truth(i)
        pred = rng.rand(224, 384) ** 3 + 0.3 * dens
        np.random.seed(0)
        a = rm.AUC_Judd(pred.copy(), fix)
        np.random.seed(0)
        assert abs(a - metrics.auc_judd(pred, fix)) < 1e-12
        assert abs(rm.NSS(pred, fix) - metrics.nss(pred, fix)) < 1e-12
        assert abs(rm.CC(pred, dens) - metrics.cc(pred, dens)) < 1e-12
        assert abs(rm.SIM(pred, dens) - metrics.sim(pred, dens)) < 1e-12


def _ref_audio_attn(depth=1):
    ref_loader.load()
    from models.audio_attention import AudioAttnNet
    return AudioAttnNet(depth=depth, heads=2, dim=512, mlp_dim=256, patch_dim=512, num_patches=16, height=7, width=12,
                        pool="cls", dim_head=64, dropout=0.0, emb_dropout=0.0).eval()


def test_audio_attn_spec_matches_reference():
    ref = _ref_audio_attn().state_dict()
    spec = synth.audio_attn_state_dict_spec()
    assert [k for k, _ in spec] == list(ref.keys())
    for k, s in spec:
        assert tuple(ref[k].shape) == tuple(s), k


@pytest.mark.parametrize("depth,batch", [(1, 2), (2, 1)])
def test_audio_attention_matches_reference(depth, batch):
    from oracle import audio_attention
    net = _ref_audio_attn(depth)
    sd = synth.make_audio_attn_state_dict(seed=3, depth=depth)
    net.load_state_dict(sd, strict=True)
    _, _, aud = synth.make_inputs(batch, audio=True, seed=77)
    with torch.no_grad():
        ref = net(aud.clone())
    assert (audio_attention.forward(sd, aud) - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("target", ["x0", "noise"])
@pytest.mark.parametrize("S", [1, 4, 10])
def test_ddpm_bit_exact_on_toy_net(target, S):
    """Ancestral sampling: the reference's p_sample / p_mean_variance / q_posterior (diffusion_trainer.py:488-520)
    driven over the same reversed(range(0, T, T // S)) sequence as sample_ddpm (:533-538), whose own encoder calls
    (:529-531) reference modules the decoder-only harness does not have."""
    ref_loader.load()
    import diffusion_trainer as dt

    class Dec(torch.nn.Module):
        def forward(self, x, t, img, audio=None):
            return _toy(x, t)

    tb = samplers.DdimTables()
    tr = dt.DiffusionTrainer.__new__(dt.DiffusionTrainer)
    tr.device = torch.device("cpu")
    tr.num_timesteps = 1000
    tr.training_target = target
    for k in ("sqrt_recip_alphas_hat", "sqrt_recipm1_alphas_hat", "posterior_mean_coef1", "posterior_mean_coef2",
              "posterior_variance", "posterior_log_variance_clipped"):
        setattr(tr, k, getattr(tb, k))
    tr.model = types.SimpleNamespace(module=types.SimpleNamespace(decoder_net=Dec()))
    x = torch.randn(2, 1, 8, 8, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(5)
    ref = x
    for t in reversed(range(0, 1000, 1000 // S)):
        ref = tr.p_sample(ref, t, [torch.zeros(1)])
    torch.manual_seed(5)
    mine = samplers.sample_ddpm(_toy, x, S, training_target=target)
    assert torch.equal(ref, mine)


def test_posterior_tables_match_reference_formulas():
    """diffusion_trainer.py:47-74 evaluated literally on the oracle's betas."""
    tb = samplers.DdimTables()
    betas = tb.betas
    alphas = 1.0 - betas
    ah = alphas.cumprod(dim=0)
    prev = torch.cat([torch.ones(1), ah[:-1]], dim=0)
    pv = betas * (1.0 - prev) / (1.0 - ah)
    assert torch.equal(tb.posterior_log_variance_clipped, torch.log(torch.maximum(pv, torch.tensor(1e-20))))
    assert torch.equal(tb.posterior_mean_coef1, betas * torch.sqrt(ah) / (1.0 - ah))
    assert torch.equal(tb.posterior_mean_coef2, (1.0 - prev) * torch.sqrt(alphas) / (1.0 - ah))


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
def test_dynamic_thresholding_bit_exact_on_toy_net(algo):
    ns = ref_loader.load()
    betas = samplers.betas_fp32()
    x = 2.5 * torch.randn(1, 1, 16, 24, generator=torch.Generator().manual_seed(1))
    net = lambda x_, t_: 1.7 * _toy(x_, t_)
    nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
    mf = ns.model_wrapper(lambda x_, t_, img, **kw: net(x_, t_), nsv, model_type="x_start", model_kwargs={},
                          guidance_type="uncond")
    ref = ns.DPM_Solver(mf, nsv, algorithm_type=algo, correcting_x0_fn="dynamic_thresholding").sample(
        x, None, steps=5, order=2, skip_type="logSNR", method="multistep", lower_order_final=False, denoise_to_zero=True)
    mine = samplers.sample_dpm(net, x, betas, steps=5, order=2, algorithm_type=algo, model_type="x_start",
                               correcting_x0_fn="dynamic_thresholding")
    assert torch.equal(ref, mine)


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("method,order,steps,skip,solver_type", [
    ("singlestep", 1, 4, "time_uniform", "dpmsolver"), ("singlestep", 2, 6, "logSNR", "dpmsolver"),
    ("singlestep", 2, 7, "logSNR", "taylor"), ("singlestep", 3, 9, "logSNR", "dpmsolver"),
    ("singlestep", 3, 10, "time_uniform", "dpmsolver"), ("singlestep", 3, 11, "logSNR", "taylor"),
    ("singlestep", 2, 5, "time_quadratic", "dpmsolver"), ("singlestep_fixed", 2, 6, "logSNR", "dpmsolver"),
    ("singlestep_fixed", 3, 9, "time_uniform", "dpmsolver")])
def test_dpm_singlestep_bit_exact_on_toy_net(algo, method, order, steps, skip, solver_type):
    """DPM_Solver.sample(method='singlestep' / 'singlestep_fixed') (sampler.py:573-795,1216-1239).
    The toy network ignores the conditioning, which the reference's singlestep branch does not forward."""
    ns = ref_loader.load()
    betas = samplers.betas_fp32()
    x = torch.randn(1, 1, 16, 24, generator=torch.Generator().manual_seed(5))
    nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
    for mtype in ("x_start", "noise"):
        mf = ns.model_wrapper(lambda x_, t_, img, **kw: _toy(x_, t_), nsv, model_type=mtype, model_kwargs={}, guidance_type="uncond")
        ref = ns.DPM_Solver(mf, nsv, algorithm_type=algo).sample(x, None, steps=steps, order=order, skip_type=skip, method=method,
                                                                  denoise_to_zero=True, solver_type=solver_type)
        mine = samplers.sample_dpm_singlestep(_toy, x, betas, steps=steps, order=order, algorithm_type=algo, model_type=mtype,
                                              skip_type=skip, method=method, denoise_to_zero=True, solver_type=solver_type)
        assert torch.equal(ref, mine), (mtype, (ref - mine).abs().max().item())


@pytest.mark.parametrize("algo", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("order,solver_type", [(2, "dpmsolver"), (3, "dpmsolver"), (3, "taylor")])
def test_dpm_adaptive_bit_exact_on_toy_net(algo, order, solver_type):
    """DPM_Solver.sample(method='adaptive') -> dpm_solver_adaptive (sampler.py:958-1009): same accepted / rejected steps,
    same iterate, bit for bit."""
    ns = ref_loader.load()
    betas = samplers.betas_fp32()
    x = torch.randn(2, 1, 16, 24, generator=torch.Generator().manual_seed(6))
    nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
    mf = ns.model_wrapper(lambda x_, t_, img, **kw: _toy(x_, t_), nsv, model_type="noise", model_kwargs={}, guidance_type="uncond")
    ref = ns.DPM_Solver(mf, nsv, algorithm_type=algo).sample(x, None, order=order, method="adaptive", denoise_to_zero=False,
                                                              solver_type=solver_type, atol=0.0078, rtol=0.05)
    mine, times, nfe = samplers.sample_dpm_adaptive(_toy, x, betas, order=order, algorithm_type=algo, model_type="noise",
                                                    solver_type=solver_type, return_model_times=True)
    assert torch.equal(ref, mine), (ref - mine).abs().max().item()
    assert nfe >= order and len(times) > order


def test_ddpm_steps_bit_exact_on_toy_net(monkeypatch):
    """util/denoising.py:39-67.  The reference hard-codes .to('cuda') / .to('cpu') hops (:48,55,64); they are made
    no-ops for this CPU run (device plumbing only, no arithmetic)."""
    ref_loader.load()
    import util.denoising as den
    orig_to = torch.Tensor.to

    def to(self, *args, **kwargs):
        if args and args[0] in ("cuda", "cpu"):
            return self
        return orig_to(self, *args, **kwargs)

    monkeypatch.setattr(torch.Tensor, "to", to)
    betas = samplers.betas_fp32()
    x = torch.randn(2, 1, 8, 8, generator=torch.Generator().manual_seed(4))
    seq = range(0, 1000, 250)
    torch.manual_seed(9)
    rxs, rx0 = den.ddpm_steps(x, seq, lambda x_, t_: _toy(x_, t_), betas)
    torch.manual_seed(9)
    mxs, mx0 = samplers.ddpm_steps(x, seq, lambda x_, t_: _toy(x_, t_), betas)
    assert len(rxs) == len(mxs) == 5 and len(rx0) == len(mx0) == 4
    for a, b in zip(rxs + rx0, mxs + mx0):
        assert torch.equal(a, b)


def test_vggish_matches_reference():
    from oracle import vggish
    ref_loader.load()
    from models.vggish import VGGish
    net = VGGish(pretrained=False).eval()
    ref_sd = net.state_dict()
    assert [k for k, _ in synth.vggish_state_dict_spec()] == list(ref_sd.keys())
    for k, s in synth.vggish_state_dict_spec():
        assert tuple(ref_sd[k].shape) == tuple(s), k
    sd = synth.make_vggish_state_dict(seed=2)
    net.load_state_dict(sd, strict=False)
    x = synth.make_audio_input(1, seed=9).view(-1, 1, 112, 192)[:3]
    with torch.no_grad():
        ref = net.forward_feat(x)
    assert (vggish.forward_feat(sd, x) - ref).abs().max().item() < 1e-4


def test_validation_losses_match_reference():
    """models/sal_losses.py kldiv2 / cc_s2 / similarity2 / nss2 behind get_kl_cc_sim_loss_wo_weight."""
    import types
    ref_loader.load()
    import models.sal_losses as ref
    from oracle import sal_losses as mine
    g = torch.Generator().manual_seed(12)
    pred = torch.rand(3, 1, 56, 96, generator=g) * 0.8 + 0.05
    gt = torch.rand(3, 1, 56, 96, generator=g) ** 4
    for name in ("kldiv2", "cc_s2", "similarity2", "nss2"):
        a, b = getattr(ref, name)(pred, gt), getattr(mine, name)(pred, gt)
        assert abs(a.item() - b.item()) <= 2e-6 * max(1.0, abs(a.item())), name
    for kl in (True, False):
        cfg = types.SimpleNamespace(loss=types.SimpleNamespace(loss_kl=kl, loss_cc=True, loss_sim=True, loss_nss=True))
        ra, mb = ref.get_kl_cc_sim_loss_wo_weight(cfg, pred, gt), mine.get_kl_cc_sim_loss_wo_weight(kl, pred, gt)
        assert set(ra) == set(mb)
        for k in ra:
            assert abs(float(ra[k]) - float(mb[k])) <= 2e-6 * max(1.0, abs(float(ra[k]))), k


def test_mvit_matches_reference():
    """SURVEY 8f row N2 (video half): MViTv2-S (models/mvit.py:796-1152) -- key list, shapes, and one clip through the
    oracle restatement against the unmodified reference on seeded 'wide' weights."""
    from oracle import mvit
    ref_loader.load()
    from models.mvit import MViT
    net = MViT(arch="small", pretrained=None, out_scales=[0, 1, 2, 3]).eval()
    ref_sd = net.state_dict()
    spec = synth.mvit_state_dict_spec()
    assert [k for k, _ in spec] == list(ref_sd.keys())
    for k, s in spec:
        assert tuple(ref_sd[k].shape) == tuple(s), k
    sd = synth.make_mvit_state_dict()
    net.load_state_dict(sd, strict=True)
    x = synth.make_video_input(1)
    with torch.no_grad():
        ref = net(x)
        raw4d = net(x.reshape(-1, 3, 224, 384))              # the loader's raw 4-D view (mvit.py:1110-1111)
    mine = mvit.forward(sd, x)
    assert [tuple(a.shape) for a in mine] == [(1, 768, 8, 7, 12), (1, 384, 8, 14, 24), (1, 192, 8, 28, 48), (1, 96, 8, 56, 96)]
    for a, b, c in zip(ref, mine, raw4d):
        assert (a - b).abs().max().item() <= 2e-4 * max(1.0, a.abs().max().item())
        assert torch.equal(a, c)
