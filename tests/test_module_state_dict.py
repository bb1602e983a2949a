"""CPU: the nn.Module checkpoint hooks of the drop-in modules (diff_sal_b200/_module.py) with a fake engine --
what a parent module's load_state_dict(strict=0) / state_dict() (model.py:17-22, diffusion_trainer.py:263-280) reach."""
import pytest
import torch
import torch.nn as nn

from diff_sal_b200 import synth
from diff_sal_b200.engine import DsbError


class FakeEngine:
    def __init__(self):
        self.loaded = None
        self.closed = False

    def load_state_dict(self, sd):
        self.loaded = dict(sd)

    def close(self):
        self.closed = True


@pytest.fixture()
def SalUNet(monkeypatch):
    from diff_sal_b200.salunet import SalUNetB200
    monkeypatch.setattr(SalUNetB200, "_make_engine", lambda self: FakeEngine())
    return SalUNetB200


class Parent(nn.Module):
    def __init__(self, dec):
        super().__init__()
        self.backbone = nn.Linear(4, 4)
        self.decoder_net = dec


class Wrapped(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.module = m


def test_parent_load_reaches_the_engine_and_roundtrips(SalUNet):
    dec = SalUNet(max_batch=1)
    top = Wrapped(Parent(dec))
    sd = synth.make_state_dict("wide")
    ck = {"module.decoder_net." + k: v for k, v in sd.items()}
    ck.update({"module.backbone." + k: v for k, v in nn.Linear(4, 4).state_dict().items()})
    msg = top.load_state_dict(ck, strict=False)
    assert not msg.missing_keys and not msg.unexpected_keys
    eng = dec.engine
    assert set(eng.loaded) == {k for k in sd if not k.endswith("num_batches_tracked")}
    assert torch.equal(eng.loaded["conv_in.weight"], sd["conv_in.weight"])
    out = top.state_dict()
    for k, v in sd.items():
        assert torch.equal(out["module.decoder_net." + k], v)
    assert "module.backbone.weight" in out
    # strict load of the same checkpoint also passes and replaces the engine
    top.load_state_dict(ck, strict=True)
    assert eng.closed and dec.engine is not eng


def test_checkpoint_without_the_decoder_is_skipped_when_not_strict(SalUNet):
    dec = SalUNet(max_batch=1)
    top = Parent(dec)
    msg = top.load_state_dict({"backbone.weight": torch.zeros(4, 4), "backbone.bias": torch.zeros(4)}, strict=False)
    assert any(k.startswith("decoder_net.") for k in msg.missing_keys)
    with pytest.raises(DsbError):
        dec.engine                                          # no weights: the product path fails loudly


def test_partial_decoder_weights_are_an_error_even_when_not_strict(SalUNet):
    dec = SalUNet(max_batch=1)
    sd = synth.make_state_dict("wide")
    del sd["logits.linear_pred.bias"]
    with pytest.raises(RuntimeError, match="partial weights"):
        Parent(dec).load_state_dict({"decoder_net." + k: v for k, v in sd.items()}, strict=False)
    with pytest.raises(DsbError):
        dec.load_state_dict(sd)


def test_shape_mismatch_and_unexpected_keys(SalUNet):
    dec = SalUNet(max_batch=1)
    sd = synth.make_state_dict("wide")
    bad = dict(sd)
    bad["conv_in.weight"] = torch.zeros(96, 1, 5, 5)
    with pytest.raises(DsbError, match="size mismatch"):
        dec.load_state_dict(bad)
    extra = dict(sd)
    extra["not_a_reference_key"] = torch.zeros(1)
    with pytest.raises(DsbError, match="not_a_reference_key"):
        dec.load_state_dict(extra, strict=True)
    dec.load_state_dict(extra, strict=False)                # tolerated, like nn.Module
    assert "not_a_reference_key" not in dec.state_dict()


def test_direct_load_with_prefix(SalUNet):
    dec = SalUNet(max_batch=1)
    sd = synth.make_state_dict("ref_init")
    dec.load_state_dict({"module.decoder_net." + k: v for k, v in sd.items()}, prefix="module.decoder_net.")
    assert torch.equal(dec.state_dict()["temb.dense.0.weight"], sd["temb.dense.0.weight"])


def test_vggish_and_audio_modules(monkeypatch):
    from diff_sal_b200.audio_attention import AudioAttnNetB200
    from diff_sal_b200.vggish import VGGishB200
    monkeypatch.setattr(VGGishB200, "_make_engine", lambda self: FakeEngine())
    monkeypatch.setattr(AudioAttnNetB200, "_make_engine", lambda self: FakeEngine())
    v = VGGishB200()
    v.load_state_dict(synth.make_vggish_state_dict())                       # features only: the embeddings head is optional
    assert all(k.startswith("features.") for k in v.engine.loaded)
    full = synth.make_vggish_state_dict(with_embeddings=False)
    full["embeddings.0.bias"] = torch.zeros(4096)
    v.load_state_dict(full)
    assert "embeddings.0.bias" in v.state_dict() and "embeddings.0.bias" not in v.engine.loaded
    a = AudioAttnNetB200(depth=1, heads=2, mlp_dim=256, dim=512, patch_dim=512, height=7, width=12)
    sd = synth.make_audio_attn_state_dict()
    a.load_state_dict(sd)
    assert all(k.startswith("transformer.") for k in a.engine.loaded)
    assert set(a.state_dict()) == set(sd)
    only_used = {k: t for k, t in sd.items() if k.startswith("transformer.")}
    with pytest.raises(DsbError):
        a.load_state_dict(only_used, strict=True)           # the reference module owns pos_embedding / to_patch_embedding too
    a.load_state_dict(only_used, strict=False)
