"""Synthetic weights and inputs of the reference's default shapes.

There is no network for checkpoints or datasets, so parity tests, smoke() and bench.py
run on random-init weights keyed exactly like the reference ``SalUNet.state_dict()``
(215 tensors, /root/reference/models/saliency_decoder/sal_unet.py:146-277) and on
synthetic clips shaped like cfgs/audio_visual.py (224x384 maps, MViT feature pyramid
of diff_model.py:106-111, VGGish/AudioAttnNet features [B,512,9,7,12]).

All tensors are drawn from seeded CPU generators so that the CUDA path, the oracle
and the committed golden fixtures see bit-identical inputs.
"""
import math

import torch

STAGE_C = (768, 384, 192, 96)
STAGE_HW = ((7, 12), (14, 24), (28, 48), (56, 96))
STAGE_S = (2, 4, 8, 16)
IMG_HW = (224, 384)
AUDIO_SHAPE = (512, 9, 7, 12)


def state_dict_spec():
    """[(key, shape)] in the reference's state_dict order."""
    spec = []
    dec = "invpt_decoder."
    for i, c in enumerate(STAGE_C):
        spec += [(dec + "norm_mts.%d.weight" % i, (c,)), (dec + "norm_mts.%d.bias" % i, (c,))]
    for i, c in enumerate(STAGE_C):
        spec.append((dec + "redu_chan_up.%d.proj.0.weight" % i, (768, c, 5, 1, 1)))
    for i, c in enumerate(STAGE_C):
        st = dec + "mid_stages.%d." % i
        if i > 0:
            cin = STAGE_C[i - 1]
            pe = st + "patch_embed.0.proj."
            for conv, bn, ci in (("1", "2", cin), ("4", "5", c)):
                spec.append((pe + conv + ".weight", (c, ci, 3, 3)))
                spec += [(pe + bn + ".weight", (c,)), (pe + bn + ".bias", (c,)),
                         (pe + bn + ".running_mean", (c,)), (pe + bn + ".running_var", (c,)),
                         (pe + bn + ".num_batches_tracked", ())]
        b = st + "blocks.0."
        s = STAGE_S[i]
        spec += [(b + "mlp.fc1.weight", (2 * c, c)), (b + "mlp.fc1.bias", (2 * c,)),
                 (b + "mlp.fc2.weight", (c, 2 * c)), (b + "mlp.fc2.bias", (c,)),
                 (b + "norm.weight", (c,)), (b + "norm.bias", (c,)),
                 (b + "attn.conv_proj_q.conv.weight", (c, 1, 3, 3, 3)),
                 (b + "attn.conv_proj_q.bn.weight", (c,)), (b + "attn.conv_proj_q.bn.bias", (c,)),
                 (b + "attn.conv_proj_k.conv.weight", (c, 1, 1, s, s)),
                 (b + "attn.conv_proj_k.bn.weight", (c,)), (b + "attn.conv_proj_k.bn.bias", (c,)),
                 (b + "attn.conv_proj_v.conv.weight", (c, 1, 1, s, s)),
                 (b + "attn.conv_proj_v.bn.weight", (c,)), (b + "attn.conv_proj_v.bn.bias", (c,))]
        for n in ("proj_q", "proj_k", "proj_v", "proj"):
            spec += [(b + "attn.%s.weight" % n, (c, c)), (b + "attn.%s.bias" % n, (c,))]
        spec += [(b + "norm2.weight", (c,)), (b + "norm2.bias", (c,)),
                 (b + "align_conv.weight", (c, 512, 1, 1)), (b + "align_conv.bias", (c,))]
    spec += [(dec + "mt_proj.0.weight", (96, 768, 3, 3)), (dec + "mt_proj.0.bias", (96,)),
             (dec + "mt_proj.1.weight", (96,)), (dec + "mt_proj.1.bias", (96,)),
             (dec + "mt_proj.1.running_mean", (96,)), (dec + "mt_proj.1.running_var", (96,)),
             (dec + "mt_proj.1.num_batches_tracked", ()),
             ("logits.linear_pred.weight", (1, 96, 1, 1)), ("logits.linear_pred.bias", (1,)),
             ("temb.dense.0.weight", (384, 96)), ("temb.dense.0.bias", (384,)),
             ("temb.dense.1.weight", (384, 384)), ("temb.dense.1.bias", (384,)),
             ("conv_in.weight", (96, 1, 3, 3)), ("conv_in.bias", (96,)),
             ("down1.conv.weight", (96, 96, 3, 3)), ("down1.conv.bias", (96,))]
    cin = 96
    for i, c in enumerate((192, 384, 768)):
        r = "res_encoder.%d.0." % i
        spec += [(r + "norm1.weight", (cin,)), (r + "norm1.bias", (cin,)),
                 (r + "conv1.weight", (c, cin, 3, 3)), (r + "conv1.bias", (c,)),
                 (r + "temb_proj.weight", (c, 384)), (r + "temb_proj.bias", (c,)),
                 (r + "norm2.weight", (c,)), (r + "norm2.bias", (c,)),
                 (r + "conv2.weight", (c, c, 3, 3)), (r + "conv2.bias", (c,)),
                 (r + "nin_shortcut.weight", (c, cin, 1, 1)), (r + "nin_shortcut.bias", (c,)),
                 ("res_encoder.%d.1.conv.weight" % i, (c, c, 3, 3)), ("res_encoder.%d.1.conv.bias" % i, (c,))]
        cin = c
    return spec


def _is_norm_affine(key):
    return (".norm" in key or "norm_mts" in key or ".bn." in key or key.endswith(".weight") and False)


def make_state_dict(kind="wide", seed=0):
    """Random weights keyed like the reference.

    kind="ref_init": what ``SalUNet.init_weights`` produces (sal_unet.py:263-277): every
        conv/linear weight ~ N(0, 0.01), biases 0, LayerNorm/BatchNorm weight 1 bias 0, BN
        running stats (0, 1), GroupNorm default affine (1, 0).  The output map is then nearly
        flat (span ~0.04), which makes min-max-normalised comparisons very strict.
    kind="wide": fan-in scaled weights, non-zero biases, randomised norm affines and BN
        running statistics, so that every term of every formula influences the output
        (a test on ref_init alone would not notice a dropped bias or a swapped BN stat).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in state_dict_spec():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
            continue
        leaf = key.rsplit(".", 1)[1]
        parent = key.rsplit(".", 1)[0]
        is_norm = (len(shape) == 1 and leaf in ("weight", "bias") and
                   (".norm" in key or "norm_mts" in key or parent.endswith(".bn")
                    or parent.endswith("proj.2") or parent.endswith("proj.5") or parent.endswith("mt_proj.1")))
        if kind == "ref_init":
            if leaf == "running_mean":
                v = torch.zeros(shape)
            elif leaf == "running_var":
                v = torch.ones(shape)
            elif is_norm:
                v = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif leaf == "bias":
                v = torch.zeros(shape)
            else:
                v = torch.randn(shape, generator=g) * 0.01
        elif kind == "wide":
            if leaf == "running_mean":
                v = torch.randn(shape, generator=g) * 0.1
            elif leaf == "running_var":
                v = 0.5 + torch.rand(shape, generator=g)
            elif is_norm:
                v = (1.0 + 0.2 * torch.randn(shape, generator=g)) if leaf == "weight" \
                    else 0.1 * torch.randn(shape, generator=g)
            elif leaf == "bias":
                v = 0.05 * torch.randn(shape, generator=g)
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                v = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        else:
            raise ValueError(kind)
        sd[key] = v.float().contiguous()
    return sd


def audio_attn_state_dict_spec(depth=1):
    """(key, shape) of ``AudioAttnNet(depth, heads=2, dim=512, mlp_dim=256, patch_dim=512, dim_head=64)`` in the
    reference's registration order (models/audio_attention.py:96-130, cfgs/audio_visual.py:34-48)."""
    spec = [("pos_embedding", (1, 1, 9, 1, 1)),
            ("to_patch_embedding.0.weight", (512,)), ("to_patch_embedding.0.bias", (512,)),
            ("to_patch_embedding.1.weight", (512, 512)), ("to_patch_embedding.1.bias", (512,)),
            ("to_patch_embedding.2.weight", (512,)), ("to_patch_embedding.2.bias", (512,)),
            ("transformer.norm.weight", (512,)), ("transformer.norm.bias", (512,))]
    for i in range(depth):
        a, f = "transformer.layers.%d.0." % i, "transformer.layers.%d.1." % i
        spec += [(a + "norm.weight", (512,)), (a + "norm.bias", (512,)), (a + "to_qkv.weight", (384, 512)),
                 (a + "to_out.0.weight", (512, 128)), (a + "to_out.0.bias", (512,)),
                 (f + "net.0.weight", (512,)), (f + "net.0.bias", (512,)), (f + "net.1.weight", (256, 512)),
                 (f + "net.1.bias", (256,)), (f + "net.4.weight", (512, 256)), (f + "net.4.bias", (512,))]
    return spec


def make_audio_attn_state_dict(seed=0, depth=1):
    """'wide' random weights for the audio transformer: fan-in scaled linears, non-zero biases, randomised LayerNorm
    affines (so that every term of every formula influences the output)."""
    g = torch.Generator().manual_seed(seed + 7919)
    sd = {}
    for key, shape in audio_attn_state_dict_spec(depth):
        leaf = key.rsplit(".", 1)[-1]
        is_norm = len(shape) == 1 and ("norm" in key or key.endswith("net.0.weight") or key.endswith("net.0.bias")
                                       or "to_patch_embedding.0" in key or "to_patch_embedding.2" in key)
        if key == "pos_embedding":
            v = torch.randn(shape, generator=g)
        elif is_norm:
            v = (1.0 + 0.2 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            v = 0.05 * torch.randn(shape, generator=g)
        else:
            v = torch.randn(shape, generator=g) * (1.0 / math.sqrt(shape[1]))
        sd[key] = v.float().contiguous()
    return sd


def vggish_state_dict_spec():
    """(key, shape) of the reference ``VGGish`` (models/vggish.py:66-124) in registration order."""
    spec, cin = [], 1
    for idx, cout in zip((0, 3, 6, 8, 11, 13), (64, 128, 256, 256, 512, 512)):
        spec += [("features.%d.weight" % idx, (cout, cin, 3, 3)), ("features.%d.bias" % idx, (cout,))]
        cin = cout
    for idx, (o, i) in zip((0, 2, 4), ((4096, 512 * 4 * 6), (4096, 4096), (128, 4096))):
        spec += [("embeddings.%d.weight" % idx, (o, i)), ("embeddings.%d.bias" % idx, (o,))]
    return spec


def make_vggish_state_dict(seed=0, with_embeddings=False):
    """He-scaled conv weights + small biases (ReLU stack keeps O(1) activations).  The ``embeddings`` MLP (113 M
    parameters, unused by forward_feat) is only materialised on request."""
    g = torch.Generator().manual_seed(seed + 4241)
    sd = {}
    for key, shape in vggish_state_dict_spec():
        if key.startswith("embeddings") and not with_embeddings:
            continue
        if key.endswith("bias"):
            v = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        sd[key] = v.float().contiguous()
    return sd


# (in_dims, out_dims, heads, stride_q, stride_kv, rel_pos length) of the 16 blocks of the reference's MViTv2-S
# (models/mvit.py:898-903,1024-1066; rel_pos tables are sized for a square 56 x 56 token grid, :586-592)
MVIT_BLOCKS = ([(96, 96, 1, 1, 8, 111), (96, 192, 2, 2, 4, 55), (192, 192, 2, 1, 4, 55), (192, 384, 4, 2, 2, 27)] +
               [(384, 384, 4, 1, 2, 27)] * 10 + [(384, 768, 8, 2, 1, 27), (768, 768, 8, 1, 1, 13)])


def mvit_state_dict_spec():
    """(key, shape) of ``MViT(arch="small", out_scales=[0, 1, 2, 3])`` (models/mvit.py:796-1105) in registration order."""
    spec = [("cls_token", (1, 1, 96)), ("patch_embed.projection.weight", (96, 3, 3, 7, 7)), ("patch_embed.projection.bias", (96,))]
    for i, (cin, cout, heads, sq, skv, rel) in enumerate(MVIT_BLOCKS):
        b = "blocks.%d." % i
        spec += [(b + "norm1.weight", (cin,)), (b + "norm1.bias", (cin,)),
                 (b + "attn.rel_pos_h", (rel, 96)), (b + "attn.rel_pos_w", (rel, 96)), (b + "attn.rel_pos_t", (15, 96)),
                 (b + "attn.qkv.weight", (3 * cout, cin)), (b + "attn.qkv.bias", (3 * cout,)),
                 (b + "attn.proj.weight", (cout, cout)), (b + "attn.proj.bias", (cout,))]
        for n in ("q", "k", "v"):
            spec += [(b + "attn.pool_%s.weight" % n, (96, 1, 3, 3, 3)), (b + "attn.norm_%s.weight" % n, (96,)),
                     (b + "attn.norm_%s.bias" % n, (96,))]
        spec += [(b + "norm2.weight", (cout,)), (b + "norm2.bias", (cout,)),
                 (b + "mlp.fc1.weight", (4 * cout, cout)), (b + "mlp.fc1.bias", (4 * cout,)),
                 (b + "mlp.fc2.weight", (cout, 4 * cout)), (b + "mlp.fc2.bias", (cout,))]
        if cin != cout:
            spec += [(b + "proj.weight", (cout, cin)), (b + "proj.bias", (cout,))]
    for s_, c in enumerate((96, 192, 384, 768)):
        spec += [("norm%d.weight" % s_, (c,)), ("norm%d.bias" % s_, (c,))]
    return spec


def make_mvit_state_dict(seed=0):
    """'wide' random weights for the video encoder: fan-in scaled linears / convolutions, non-zero biases, randomised
    LayerNorm affines, O(0.3) relative-position tables and depthwise pooling taps (so that every term matters)."""
    g = torch.Generator().manual_seed(seed + 1013)
    sd = {}
    for key, shape in mvit_state_dict_spec():
        leaf = key.rsplit(".", 1)[-1]
        if key == "cls_token":
            v = 0.5 * torch.randn(shape, generator=g)
        elif "rel_pos" in key:
            v = 0.3 * torch.randn(shape, generator=g)
        elif "norm" in key and len(shape) == 1:
            v = (1.0 + 0.2 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            v = 0.05 * torch.randn(shape, generator=g)
        elif ".pool_" in key:
            v = torch.randn(shape, generator=g) * 0.25
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[key] = v.float().contiguous()
    return sd


def make_video_input(batch, seed=7321, hw=IMG_HW):
    """ImageNet-normalised-range video clips as the loader hands them to the video encoder: [B, 3, 16, 224, 384]."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        out.append(torch.randn((1, 3, 16) + tuple(hw), generator=g))
    return torch.cat(out, 0)


def make_audio_input(batch, seed=4321, frames=9, hw=(112, 192)):
    """Log-mel-like audio patches as the loader hands them to forward_vggish (datasets/saliency_db.py:303-305,457):
    [B, 1, 9, 112, 192], values ~ N(-2, 1.5)."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        out.append(torch.randn((1, 1, frames) + tuple(hw), generator=g) * 1.5 - 2.0)
    return torch.cat(out, 0)


def make_inputs(batch, audio=True, seed=1234):
    """Seeded synthetic clip batch: x_T, the four MViT-shaped feature tensors and the
    audio feature tensor.  Clip ``i`` uses generator seed ``seed + i`` so that a clip's
    inputs do not depend on which rank / micro-batch it lands in."""
    xs, feats, auds = [], [[] for _ in range(4)], []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        xs.append(torch.randn((1, 1) + IMG_HW, generator=g))
        for j, (c, (h, w)) in enumerate(zip(STAGE_C, STAGE_HW)):
            feats[j].append(torch.randn((1, c, 8, h, w), generator=g))
        if audio:
            auds.append(torch.randn((1,) + AUDIO_SHAPE, generator=g))
    x = torch.cat(xs, 0)
    feat_list = [torch.cat(f, 0) for f in feats]
    aud = torch.cat(auds, 0) if audio else None
    return x, feat_list, aud
