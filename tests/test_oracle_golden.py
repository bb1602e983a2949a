"""CPU: the oracle restatement (oracle/) against the committed golden fixtures that were
generated from the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from diff_sal_b200 import synth
from oracle import salunet, samplers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 5e-6          # fp32 re-association only (restatement vs reference modules)


def gold(name):
    return torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])


@pytest.mark.parametrize("name,kind,audio,t", [
    ("step_wide_av_t500", "wide", True, [500]),
    ("step_wide_vis_t500", "wide", False, [500]),
    ("step_refinit_av_t37", "ref_init", True, [37]),
    ("step_wide_av_t886p9", "wide", True, [886.9]),
])
def test_single_evaluation(name, kind, audio, t):
    sd = synth.make_state_dict(kind)
    x, feats, aud = synth.make_inputs(1, audio=audio)
    y = salunet.forward(sd, x, torch.tensor(t), feats, aud)
    assert (y - gold(name)).abs().max().item() < TOL


def _net(kind, audio):
    sd = synth.make_state_dict(kind)
    x, feats, aud = synth.make_inputs(1, audio=audio)
    return x, (lambda x_, t_: salunet.forward(sd, x_, t_, feats, aud))


@pytest.mark.parametrize("S", [1, 5])
def test_ddim(S):
    x, net = _net("wide", True)
    y = samplers.sample_ddim(net, x, S, eta=0.0, training_target="x0")
    assert (y - gold("ddim%d_wide_av" % S)).abs().max().item() < TOL


@pytest.mark.parametrize("name,algo,mtype,tol", [
    ("dpm_wide_av_xstart_o2_s4", "dpmsolver", "x_start", TOL),
    ("dpmpp_wide_av_xstart_o2_s4", "dpmsolver++", "x_start", TOL),
    ("dpm_wide_av_noise_o2_s4", "dpmsolver", "noise", 2e-3),   # values reach +-670
])
def test_dpm_solver(name, algo, mtype, tol):
    x, net = _net("wide", True)
    y = samplers.sample_dpm(net, x, steps=4, order=2, algorithm_type=algo, model_type=mtype)
    assert (y - gold(name)).abs().max().item() < tol


def test_config1_visual_only_dpm():
    """BASELINE config 1: visual-only, batch 1, reference init, DPM-solver multistep-2, 10 NFE."""
    x, net = _net("ref_init", False)
    y = samplers.sample_dpm(net, x, steps=9, order=2, algorithm_type="dpmsolver", model_type="x_start")
    assert (y - gold("cfg1_dpm_refinit_vis_xstart_o2_s9")).abs().max().item() < TOL


def test_config2_bench_configuration():
    """BASELINE config 2 (what bench.py times): audio-visual, DPM-solver multistep-2, 10 NFE; clips 6-7 of the batch of 8
    as one oracle batch of 2 (the oracle is batch-capable; the reference's x_start conversion is not, see
    tests/golden/make_golden_cfg2.py) against the reference fixture."""
    sd = synth.make_state_dict("wide")
    x, feats, aud = synth.make_inputs(8, audio=True)
    x, feats, aud = x[6:], [f[6:] for f in feats], aud[6:]
    y = samplers.sample_dpm(lambda x_, t_: salunet.forward(sd, x_, t_, feats, aud), x, steps=9, order=2,
                            algorithm_type="dpmsolver", model_type="x_start")
    assert (y - gold("cfg2_dpm_wide_av_b8_s9")[6:]).abs().max().item() < 1e-5


def test_ddim_10_steps_batch2():
    x, feats, aud = synth.make_inputs(2, audio=True)
    sd = synth.make_state_dict("wide")
    y = samplers.sample_ddim(lambda x_, t_: salunet.forward(sd, x_, t_, feats, aud), x, 10, eta=0.0, training_target="x0")
    assert (y - gold("ddim_wide_av_b2_s10")).abs().max().item() < 1e-5


def test_visual_only_is_noise_invariant():
    """in visual-only eval mode the output does not depend on x_t or t (only frames
    0..4 reach ReduceTemp and the noise slice is frame 8)."""
    sd = synth.make_state_dict("wide")
    x, feats, _ = synth.make_inputs(1, audio=False)
    a = salunet.forward(sd, x, torch.tensor([500]), feats, None)
    b = salunet.forward(sd, torch.randn_like(x), torch.tensor([3]), feats, None)
    assert torch.equal(a, b)


def test_audio_attention():
    """the once-per-clip audio transformer (models/audio_attention.py) against the reference fixture."""
    from oracle import audio_attention
    sd = synth.make_audio_attn_state_dict()
    _, _, aud = synth.make_inputs(1, audio=True)
    y = audio_attention.forward(sd, aud)
    assert (y - gold("audio_attn_wide_b1")).abs().max().item() < 2e-5        # values reach +-6


def test_audio_attention_feeds_decoder():
    """diff_model.py:70-113: decoder conditioned on the transformer output, one evaluation at t = 500."""
    from oracle import audio_attention
    x, feats, aud = synth.make_inputs(1, audio=True)
    emb = audio_attention.forward(synth.make_audio_attn_state_dict(), aud)
    y = salunet.forward(synth.make_state_dict("wide"), x, torch.tensor([500]), feats, emb)
    assert (y - gold("step_wide_av_attn_t500")).abs().max().item() < TOL


def test_vggish_features():
    """SURVEY 8f row N2 (audio half): VGGish.forward_feat (models/vggish.py:87-103) against the reference fixture."""
    from oracle import vggish
    y = vggish.forward_feat(synth.make_vggish_state_dict(), synth.make_audio_input(1).view(-1, 1, 112, 192))
    assert y.shape == (9, 512, 7, 12)
    assert (y[:, ::4] - gold("vggish_feat_b1_c4")).abs().max().item() < 1e-4         # values reach ~30
