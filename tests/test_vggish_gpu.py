"""GPU parity of the VGGish feature stack (audio half) through the C ABI (dsb_vggish_*).

Six conv + ReLU layers with bf16 operands / bf16 inter-layer activations and fp32 accumulation: the output (values up to
~30) is compared with the fp32 oracle at relative RMS <= 1e-2 and max-abs <= 2 % of the output range.  Downstream, the
whole audio side (VGGish -> AudioAttnNet -> one decoder evaluation) must meet the path's stated tolerance (max-abs <= 1e-2
on the min-max-normalised map)."""
import os

import numpy as np
import pytest
import torch

from diff_sal_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _net(max_frames=27, seed=0):
    from diff_sal_b200.vggish import VGGishB200
    net = VGGishB200(pretrained=False, max_frames=max_frames)
    net.load_state_dict(synth.make_vggish_state_dict(seed=seed))
    return net


def _close(y, ref):
    d = y.double() - ref.double()
    assert (d.pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt()).item() <= 1e-2
    assert d.abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_golden_fixture_and_oracle():
    from oracle import vggish
    x = synth.make_audio_input(1).view(-1, 1, 112, 192)
    net = _net()
    y = net.forward_feat(x.cuda()).cpu()
    assert y.shape == (9, 512, 7, 12) and net.engine.last_launch_count == 9
    _close(y[:, ::4], torch.from_numpy(np.load(os.path.join(GOLD, "vggish_feat_b1_c4.npz"))["y"]))
    _close(y, vggish.forward_feat(synth.make_vggish_state_dict(), x))


def test_batch_invariant_and_repeatable():
    x = synth.make_audio_input(3, seed=5).view(-1, 1, 112, 192)
    net = _net()
    y = net.forward_feat(x.cuda())
    assert torch.equal(y, net.forward_feat(x.cuda()))
    assert torch.equal(y[9:18], net.forward_feat(x[9:18].cuda()))


def test_audio_side_feeds_decoder_within_tolerance():
    """forward_vggish (models/diff_model.py:70-81): VGGish -> '(b t) c h w -> b c t h w' -> AudioAttnNet -> decoder."""
    from diff_sal_b200.audio_attention import AudioAttnNetB200
    from diff_sal_b200.salunet import SalUNetB200
    from oracle import audio_attention, salunet, samplers, vggish
    a = synth.make_audio_input(1)
    x, feats, _ = synth.make_inputs(1, audio=True)
    # B200 chain
    fm = _net().forward_feat(a.view(-1, 1, 112, 192).cuda())                       # [9,512,7,12]
    fm = fm.reshape(1, 9, 512, 7, 12).permute(0, 2, 1, 3, 4).contiguous()          # rearrange '(b t) c h w -> b c t h w'
    att = AudioAttnNetB200(depth=1, heads=2, dim=512, mlp_dim=256, patch_dim=512, height=7, width=12, max_batch=1)
    att.load_state_dict(synth.make_audio_attn_state_dict())
    dec = SalUNetB200(max_batch=1, audio_visual=True)
    dec.load_state_dict(synth.make_state_dict("wide"))
    y = dec(x.cuda(), torch.tensor([500.0]), [f.cuda() for f in feats], att(fm)).cpu()
    # oracle chain
    rf = vggish.forward_feat(synth.make_vggish_state_dict(), a.view(-1, 1, 112, 192))
    rf = rf.reshape(1, 9, 512, 7, 12).permute(0, 2, 1, 3, 4).contiguous()
    ref = salunet.forward(synth.make_state_dict("wide"), x, torch.tensor([500]), feats,
                          audio_attention.forward(synth.make_audio_attn_state_dict(), rf))
    assert (samplers.minmax_map(y) - samplers.minmax_map(ref)).abs().max().item() <= 1e-2


def test_rejects_wrong_input_and_missing_weights():
    from diff_sal_b200.engine import DsbError
    from diff_sal_b200.vggish import VGGishB200
    with pytest.raises(DsbError):
        VGGishB200().forward_feat(torch.zeros(1, 1, 112, 192, device="cuda"))
    with pytest.raises(DsbError):
        _net().forward_feat(torch.zeros(1, 1, 96, 64, device="cuda"))
    sd = synth.make_vggish_state_dict()
    del sd["features.8.bias"]
    with pytest.raises(DsbError):
        VGGishB200().load_state_dict(sd)
