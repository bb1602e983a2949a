"""GPU: the model-container boundary (models/diff_model.py:53-58,70-114).

The UNMODIFIED reference classes (imported through oracle/ref_loader.py: /root/reference in the build container,
the bytecode build oracle/_ref/ on the GPU box) are used as they are: ``VideoSaliencyModel`` builds its sub-networks from
config dicts through the reference's own ``OBJECT_REGISTRY``, a DDP-prefixed checkpoint is loaded on the PARENT with
``strict=False`` (model.py:17-22), and ``DiffusionTrainer.sample_ddim`` (diffusion_trainer.py:439-480) -- including its
``copy.deepcopy`` of the feature list on every step -- drives the B200 decoder.
"""
import copy
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from diff_sal_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])


def minmax(x):
    flat = x.reshape(x.shape[0], -1)
    lo = flat.min(dim=1, keepdim=True).values
    hi = flat.max(dim=1, keepdim=True).values
    return ((flat - lo) / (hi - lo)).reshape(x.shape)


def decoder_cfg():
    """The literal decoder_net kwargs of cfgs/audio_visual.py:50-82."""
    from oracle import ref_loader
    return ref_loader.decoder_kwargs()


def audio_cfg():
    """cfgs/audio_visual.py:34-48."""
    return dict(depth=1, heads=2, dim=512, mlp_dim=256, patch_dim=512, num_patches=16, height=7, width=12, pool="cls",
                dim_head=64, dropout=0.0, emb_dropout=0.0)


def checkpoint(prefix="module."):
    """A DDP-style checkpoint of the whole container with the reference's key names."""
    ck = {}
    for k, v in synth.make_state_dict("wide").items():
        ck[prefix + "decoder_net." + k] = v
    for k, v in synth.make_audio_attn_state_dict().items():
        ck[prefix + "spatiotemp_net." + k] = v
    for k, v in synth.make_vggish_state_dict().items():
        ck[prefix + "audio_net." + k] = v
    return ck


class FakeDDP(nn.Module):
    """What DistributedDataParallel contributes to the key names and the attribute path: a ``module.`` level."""

    def __init__(self, m):
        super().__init__()
        self.module = m


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    ref_loader.load()
    from models.diff_model import VideoSaliencyModel      # the reference's class, unmodified
    from util.registry import OBJECT_REGISTRY
    import diffusion_trainer
    return types.SimpleNamespace(VideoSaliencyModel=VideoSaliencyModel, registry=OBJECT_REGISTRY, dt=diffusion_trainer)


@pytest.fixture(scope="module")
def container(ref):
    from diff_sal_b200.audio_attention import AudioAttnNetB200
    from diff_sal_b200.salunet import SalUNetB200
    from diff_sal_b200.vggish import VGGishB200
    model = ref.VideoSaliencyModel(
        channel_list=None, visual_net=None,
        spatiotemp_net=dict(type=AudioAttnNetB200, **audio_cfg()),
        audio_net=dict(type=VGGishB200, pretrained=False),
        decoder_net=dict(type=SalUNetB200, **decoder_cfg()))
    wrapped = FakeDDP(model).cuda().eval()
    msg = wrapped.load_state_dict(checkpoint(), strict=False)        # model.py:20: strict=0 on the wrapped parent
    # the reference container's own 128->512->768 `fc` head is not in our synthetic checkpoint; nothing else may be missing
    assert all(k.startswith("module.fc.") for k in msg.missing_keys), msg.missing_keys
    assert not msg.unexpected_keys
    yield wrapped
    for m in (model.decoder_net, model.spatiotemp_net, model.audio_net):
        m.engine.close()


def make_trainer(ref, wrapped, S):
    from oracle import samplers
    tb = samplers.DdimTables()
    tr = ref.dt.DiffusionTrainer.__new__(ref.dt.DiffusionTrainer)     # __init__ needs the argparse / yaml world
    tr.device = torch.device("cuda")
    tr.num_timesteps = 1000
    tr.training_target = "x0"
    for k in ("alphas_hat", "sqrt_alphas_hat", "sqrt_recip_alphas_hat", "sqrt_recipm1_alphas_hat"):
        setattr(tr, k, getattr(tb, k).cuda())
    tr.config = types.SimpleNamespace(sampling=types.SimpleNamespace(timesteps=S, eta=0.0))
    tr.model = wrapped
    return tr


def test_registry_build_and_parent_state_dict_roundtrip(ref, container):
    from diff_sal_b200.model import register_b200_modules
    from diff_sal_b200.salunet import SalUNetB200
    # by-name swap in the reference's registry (force: SalUNet / AudioAttnNet / VGGish are already registered there)
    register_b200_modules(ref.registry)
    built = ref.registry.build(dict(type="SalUNet", **decoder_cfg()))
    assert isinstance(built, SalUNetB200)
    sd = container.state_dict()
    want = checkpoint()
    for k, v in want.items():
        if k.endswith("num_batches_tracked"):
            continue
        assert k in sd, k
        assert torch.equal(sd[k].cpu().float(), v.float()), k


def test_reference_sample_ddim_drives_b200_decoder(ref, container):
    """DiffusionTrainer.sample_ddim, unmodified, 5 steps: same map as the reference decoder's fixture; the loop-invariant
    conditioning runs once although every step hands the decoder a fresh deep copy of the features."""
    x, feats, aud = synth.make_inputs(1, audio=True)
    eng = container.module.decoder_net.engine
    calls = []
    orig = eng.set_condition
    eng.set_condition = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        torch.manual_seed(0)
        y = make_trainer(ref, container, 5).sample_ddim(x.cuda(), [f.cuda() for f in feats], aud.cuda()).cpu()
    finally:
        eng.set_condition = orig
    assert (minmax(y) - minmax(gold("ddim5_wide_av"))).abs().max().item() <= 1e-2
    assert len(calls) == 1, "conditioning ran %d times for 5 steps" % len(calls)


def test_condition_cache_sees_content_not_addresses(container):
    """fp16 / CPU feature sets are converted to fresh tensors, so two DIFFERENT clips can present the same
    device addresses to the engine; the cache must not confuse them, and equal content at new addresses must hit."""
    dec = container.module.decoder_net
    xa, fa, aa = synth.make_inputs(1, audio=True, seed=100)
    xb, fb, ab = synth.make_inputs(1, audio=True, seed=200)
    t = torch.tensor([500.0], device="cuda")
    x = xa.cuda()
    ya = dec(x, t, [f.half() for f in fa], aa.half()).clone()         # CPU fp16 -> converted copies, then freed
    yb = dec(x, t, [f.half() for f in fb], ab.half()).clone()         # may land at the same addresses
    ya2 = dec(x, t, [f.half().cuda() for f in fa], aa.half().cuda()).clone()
    torch.cuda.synchronize()
    assert not torch.equal(ya, yb)
    assert torch.equal(ya, ya2)
    with torch.inference_mode():                                       # tensors without a version counter
        yi = dec(x.clone(), t.clone(), [f.half().cuda() for f in fa], aa.half().cuda()).clone()
    assert torch.equal(ya, yi)
    # in-place edit of a live tensor (same address, bumped version) must be seen
    fl = [f.cuda() for f in fa]
    y0 = dec(x, t, fl, aa.cuda()).clone()
    fl[0].mul_(1.5)
    y1 = dec(x, t, fl, aa.cuda()).clone()
    assert not torch.equal(y0, y1)


def test_reference_container_forward_end_to_end(ref, container):
    """VideoSaliencyModel.forward(data, t) (diff_model.py:83-114), reference code: VGGishB200.forward_feat ->
    AudioAttnNetB200 -> SalUNetB200, against the fp32 oracle chain on the same inputs.  visual_net=None makes the
    reference draw random placeholder features (diff_model.py:105-111), so the generator is seeded and the same draw is
    replayed for the oracle."""
    from oracle import audio_attention, salunet, samplers, vggish
    audio = synth.make_audio_input(1)
    x, _, _ = synth.make_inputs(1, audio=False)
    t = torch.tensor([500.0])
    data = {"img": torch.zeros(1, device="cuda"), "input": x.cuda(), "audio": audio.cuda()}
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    with torch.no_grad():
        y = container.module(data, t.cuda()).cpu()
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    vis = [torch.randn(s, device="cuda").cpu() for s in [(1, 768, 8, 7, 12), (1, 384, 8, 14, 24), (1, 192, 8, 28, 48), (1, 96, 8, 56, 96)]]
    feat = vggish.forward_feat(synth.make_vggish_state_dict(), audio.view(-1, 1, 112, 192))
    feat = feat.reshape(1, 9, 512, 7, 12).permute(0, 2, 1, 3, 4).contiguous()
    emb = audio_attention.forward(synth.make_audio_attn_state_dict(), feat)
    refy = salunet.forward(synth.make_state_dict("wide"), x, t, vis, emb)
    assert (samplers.minmax_map(y) - samplers.minmax_map(refy)).abs().max().item() <= 1e-2


def test_product_container_matches_reference_container(ref, container):
    """diff_sal_b200.model.VideoSaliencyModelB200 (no reference tree needed) == the reference container, bitwise."""
    from diff_sal_b200.model import VideoSaliencyModelB200
    m = VideoSaliencyModelB200(channel_list=None, visual_net=None,
                               spatiotemp_net=dict(type="AudioAttnNet", **audio_cfg()),
                               audio_net=dict(type="VGGish", pretrained=False),
                               decoder_net=dict(type="SalUNet", **decoder_cfg())).cuda().eval()
    FakeDDP(m).load_state_dict(checkpoint(), strict=False)
    audio = synth.make_audio_input(1).cuda()
    a1, a2 = m.forward_vggish(audio)
    b1, _ = container.module.forward_vggish(audio)
    assert a1 is a2 and torch.equal(a1, b1)
    x, _, _ = synth.make_inputs(1, audio=False)
    data = {"img": torch.zeros(1, device="cuda"), "input": x.cuda(), "audio": audio}
    t = torch.tensor([250.0], device="cuda")
    torch.manual_seed(3); torch.cuda.manual_seed(3)
    ya = m(data, t).clone()
    torch.manual_seed(3); torch.cuda.manual_seed(3)
    with torch.no_grad():
        yb = container.module(data, t).clone()
    assert torch.equal(ya, yb)
    for sub in (m.decoder_net, m.spatiotemp_net, m.audio_net):
        sub.engine.close()


def test_full_model_video_and_audio_in_map_out(ref):
    """The reference's VideoSaliencyModel with ALL four sub-networks built from the B200 classes (MViTB200, VGGishB200,
    AudioAttnNetB200, SalUNetB200): raw video + log-mel audio in, map out, against the fp32 oracle chain."""
    from diff_sal_b200.audio_attention import AudioAttnNetB200
    from diff_sal_b200.mvit import MViTB200
    from diff_sal_b200.salunet import SalUNetB200
    from diff_sal_b200.vggish import VGGishB200
    from oracle import audio_attention, mvit, salunet, samplers, vggish
    model = ref.VideoSaliencyModel(
        channel_list=None,
        visual_net=dict(type=MViTB200, arch="small", pretrained=None, out_scales=[0, 1, 2, 3]),
        spatiotemp_net=dict(type=AudioAttnNetB200, **audio_cfg()),
        audio_net=dict(type=VGGishB200, pretrained=False),
        decoder_net=dict(type=SalUNetB200, **decoder_cfg()))
    wrapped = FakeDDP(model).cuda().eval()
    ck = checkpoint()
    for k, v in synth.make_mvit_state_dict().items():
        ck["module.visual_net." + k] = v
    msg = wrapped.load_state_dict(ck, strict=False)
    assert all(k.startswith("module.fc.") for k in msg.missing_keys) and not msg.unexpected_keys
    video, audio = synth.make_video_input(1), synth.make_audio_input(1)
    x, _, _ = synth.make_inputs(1, audio=False)
    t = torch.tensor([500.0])
    # the loader hands the clip over as [B*16, 3, H, W] (diffusion_trainer.py:101-105; MViT.forward re-views it)
    data = {"img": video.reshape(-1, 3, 224, 384).cuda(), "input": x.cuda(), "audio": audio.cuda()}
    with torch.no_grad():
        y = wrapped.module(data, t.cuda()).cpu()
    vis = mvit.forward(synth.make_mvit_state_dict(), video)
    feat = vggish.forward_feat(synth.make_vggish_state_dict(), audio.view(-1, 1, 112, 192))
    feat = feat.reshape(1, 9, 512, 7, 12).permute(0, 2, 1, 3, 4).contiguous()
    emb = audio_attention.forward(synth.make_audio_attn_state_dict(), feat)
    refy = salunet.forward(synth.make_state_dict("wide"), x, t, vis, emb)
    for m in (model.decoder_net, model.spatiotemp_net, model.audio_net, model.visual_net):
        m.engine.close()
    assert (samplers.minmax_map(y) - samplers.minmax_map(refy)).abs().max().item() <= 1e-2
