"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's once-per-clip audio transformer.

Follows ``AudioAttnNet.forward`` (models/audio_attention.py:132-143) and the ``Transformer`` / ``Attention`` /
``FeedForward`` modules it drives (models/audio_attention.py:13-90) as configured by cfgs/audio_visual.py:34-48
(depth 1, heads 2 x 64, dim 512, mlp_dim 256).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product path never does.

Reference behaviour restated exactly, including its quirk: the result of ``to_patch_embedding`` + ``pos_embedding``
is discarded (audio_attention.py:134-141 rebinds ``x`` to the raw input tokens), so those parameters do not
influence the output and are ignored here.

Pinned against the real reference in tests/test_oracle_vs_reference.py and through tests/golden/audio_attn_*.npz.
"""
import torch
import torch.nn.functional as F

HEADS = 2
DIM_HEAD = 64


def _attention(sd, pre, x):
    """Attention.forward (audio_attention.py:55-69): pre-LN, bias-free qkv, softmax(q k^T * d^-0.5) v, to_out."""
    xn = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
    qkv = F.linear(xn, sd[pre + "to_qkv.weight"])
    b, n, _ = qkv.shape
    q, k, v = [t.reshape(b, n, HEADS, DIM_HEAD).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
    dots = torch.matmul(q, k.transpose(-1, -2)) * (DIM_HEAD ** -0.5)
    attn = dots.softmax(dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, HEADS * DIM_HEAD)
    return F.linear(out, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def _feed_forward(sd, pre, x):
    """FeedForward.forward (audio_attention.py:15-27): LN -> Linear -> exact-erf GELU -> Linear."""
    xn = F.layer_norm(x, (x.shape[-1],), sd[pre + "net.0.weight"], sd[pre + "net.0.bias"], 1e-5)
    h = F.gelu(F.linear(xn, sd[pre + "net.1.weight"], sd[pre + "net.1.bias"]))
    return F.linear(h, sd[pre + "net.4.weight"], sd[pre + "net.4.bias"])


def forward(sd, audio):
    """audio [B, 512, T, H, W] fp32 -> same shape (audio_attention.py:132-143)."""
    b, c, t, h, w = audio.shape
    x = audio.permute(0, 2, 3, 4, 1).reshape(b, t * h * w, c)           # 'b c t h w -> b (t h w) c'
    depth = 0
    while ("transformer.layers.%d.0.to_qkv.weight" % depth) in sd:
        depth += 1
    for i in range(depth):                                                # Transformer.forward (:85-90)
        x = _attention(sd, "transformer.layers.%d.0." % i, x) + x
        x = _feed_forward(sd, "transformer.layers.%d.1." % i, x) + x
    x = F.layer_norm(x, (c,), sd["transformer.norm.weight"], sd["transformer.norm.bias"], 1e-5)
    return x.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3).contiguous()  # 'b (t h w) c -> b c t h w'
