"""GPU: the MViTv2-S video encoder (video half; models/mvit.py:796-1152) through the C ABI against the
fp32 CPU oracle (oracle/mvit.py, itself pinned to the unmodified reference in tests/test_oracle_vs_reference.py), and a
decoder evaluation conditioned on its features."""
import pytest
import torch

from diff_sal_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    from diff_sal_b200.mvit import MViTB200
    m = MViTB200(arch="small", out_scales=[0, 1, 2, 3], max_batch=2)
    m.load_state_dict(synth.make_mvit_state_dict())
    yield m
    m.engine.close()


def rel_rms(a, b):
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_mvit_features_match_oracle(net):
    from oracle import mvit
    x = synth.make_video_input(1)
    got = [t.cpu() for t in net(x.cuda())]
    ref = mvit.forward(synth.make_mvit_state_dict(), x)
    assert [tuple(t.shape) for t in got] == [(1, 768, 8, 7, 12), (1, 384, 8, 14, 24), (1, 192, 8, 28, 48), (1, 96, 8, 56, 96)]
    for g, r in zip(got, ref):
        err = (g - r).abs().max().item()
        print("scale C=%d: rel rms %.2e, max abs %.2e of range %.2f" % (r.shape[1], rel_rms(g, r), err, r.abs().max().item()))
        # LayerNormed features (values up to ~6): fp16-operand GEMMs through up to 16 blocks (measured 4e-4 .. 1.8e-3 rms)
        assert rel_rms(g, r) <= 5e-3
        assert err <= 1e-2 * max(1.0, r.abs().max().item())
    assert net.engine.last_launch_count > 200


def test_mvit_batch_and_raw_view(net):
    """Batch of two, handed over as the loader's raw 4-D view: clip 0 equals the single-clip result bit for bit."""
    x = synth.make_video_input(2)
    one = [t.clone() for t in net(x[:1].cuda())]
    two = net(x.reshape(-1, 3, 224, 384).cuda())
    torch.cuda.synchronize()
    for a, b in zip(one, two):
        assert b.shape[0] == 2 and torch.equal(a[0], b[0])


def test_decoder_on_mvit_features(net):
    """features in -> map out: one SalUNet evaluation conditioned on the B200 encoder's features against the oracle chain
    (fp32 MViT -> fp32 SalUNet); north_star tolerance on the min-max-normalised map."""
    from diff_sal_b200.salunet import SalUNetB200
    from oracle import mvit, salunet, samplers
    xv = synth.make_video_input(1)
    x, _, aud = synth.make_inputs(1, audio=True)
    feats = net(xv.cuda())
    dec = SalUNetB200(max_batch=1, audio_visual=True)
    dec.load_state_dict(synth.make_state_dict("wide"))
    t = torch.tensor([500.0])
    y = dec(x.cuda(), t.cuda(), feats, aud.cuda()).cpu()
    ref_feats = mvit.forward(synth.make_mvit_state_dict(), xv)
    ref = salunet.forward(synth.make_state_dict("wide"), x, t, ref_feats, aud)
    dec.engine.close()
    assert (samplers.minmax_map(y) - samplers.minmax_map(ref)).abs().max().item() <= 1e-2


def test_mvit_rejects_other_configurations():
    from diff_sal_b200.engine import DsbError
    from diff_sal_b200.mvit import MViTB200
    with pytest.raises(DsbError):
        MViTB200(arch="base")
    with pytest.raises(DsbError):
        MViTB200().forward(torch.zeros(1, 3, 16, 224, 384, device="cuda"))          # no weights
    m = MViTB200()
    sd = synth.make_mvit_state_dict()
    del sd["blocks.7.attn.rel_pos_w"]
    with pytest.raises(DsbError):
        m.load_state_dict(sd)
