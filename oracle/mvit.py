"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of the reference's MViTv2-S video encoder.

Restates, as plain functional torch on a reference-keyed state_dict (arch 'small', cfgs/audio_visual.py:27-32):
  MViT.forward                  /root/reference/models/mvit.py:1107-1152  (incl. ``x = norm_s(x)`` replacing the token
                                stream at every output scale, :1133, and the reversed output list, :1152)
  PatchEmbed3D.forward          mvit.py:232-254   (Conv3d (3,7,7) / (2,4,4) / (1,3,3))
  MultiScaleBlock.forward       mvit.py:763-793
  MultiScaleAttention.forward   mvit.py:603-648
  attention_pool                mvit.py:459-510
  resize_decomposed_rel_pos / add_decomposed_rel_pos   mvit.py:330-404
Pinned against the imported reference in tests/test_oracle_vs_reference.py.
"""
import torch
import torch.nn.functional as F

HEAD_DIM = 96
# (in_dims, out_dims, heads, stride_q, stride_kv) of the 16 blocks of MViT-S (mvit.py:898-903,1024-1066)
BLOCKS = ([(96, 96, 1, 1, 8), (96, 192, 2, 2, 4), (192, 192, 2, 1, 4), (192, 384, 4, 2, 2)] +
          [(384, 384, 4, 1, 2)] * 10 + [(384, 768, 8, 2, 1), (768, 768, 8, 1, 1)])
STAGE_AFTER = {0: 0, 2: 1, 13: 2, 15: 3}          # block index -> output scale


def _ln(x, sd, key, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps=eps)


def resize_rel_pos(rel_pos, q_size, k_size):
    """mvit.py:330-363."""
    max_rel_dist = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel_dist:
        resized = F.interpolate(rel_pos.transpose(0, 1).unsqueeze(0), size=max_rel_dist, mode="linear")
        resized = resized.squeeze(0).transpose(0, 1)
    else:
        resized = rel_pos
    q_ratio = max(k_size / q_size, 1.0)
    k_ratio = max(q_size / k_size, 1.0)
    q_coords = torch.arange(q_size)[:, None] * q_ratio
    k_coords = torch.arange(k_size)[None, :] * k_ratio
    rel = (q_coords - k_coords) + (k_size - 1) * k_ratio
    return resized[rel.long()]                       # [q_size, k_size, C]


def _pool(x, w, size, stride, norm_w, norm_b):
    """attention_pool (mvit.py:459-510) with the cls token: x [B, heads, 1 + THW, 96]."""
    B, nh, L, C = x.shape
    T, H, W = size
    cls_tok, x = x[:, :, :1, :], x[:, :, 1:, :]
    x = x.reshape(B * nh, T, H, W, C).permute(0, 4, 1, 2, 3).contiguous()
    x = F.conv3d(x, w, None, stride=(1, stride, stride), padding=1, groups=C)
    out_size = tuple(x.shape[2:])
    x = x.reshape(B, nh, C, -1).transpose(2, 3)
    x = torch.cat((cls_tok, x), dim=2)
    x = F.layer_norm(x, (C,), norm_w, norm_b, eps=1e-5)
    return x, out_size


def attention(sd, p, x, size, out_dims, heads, stride_q, stride_kv):
    """MultiScaleAttention.forward (mvit.py:603-648)."""
    B, N, _ = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B, N, 3, heads, -1)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    q, q_shape = _pool(q, sd[p + "pool_q.weight"], size, stride_q, sd[p + "norm_q.weight"], sd[p + "norm_q.bias"])
    k, k_shape = _pool(k, sd[p + "pool_k.weight"], size, stride_kv, sd[p + "norm_k.weight"], sd[p + "norm_k.bias"])
    v, _ = _pool(v, sd[p + "pool_v.weight"], size, stride_kv, sd[p + "norm_v.weight"], sd[p + "norm_v.bias"])
    attn = (q * (HEAD_DIM ** -0.5)) @ k.transpose(-2, -1)
    # add_decomposed_rel_pos (mvit.py:366-404), with_cls_token=True
    q_t, q_h, q_w = q_shape
    k_t, k_h, k_w = k_shape
    Rt = resize_rel_pos(sd[p + "rel_pos_t"], q_t, k_t)
    Rh = resize_rel_pos(sd[p + "rel_pos_h"], q_h, k_h)
    Rw = resize_rel_pos(sd[p + "rel_pos_w"], q_w, k_w)
    r_q = q[:, :, 1:].reshape(B, heads, q_t, q_h, q_w, HEAD_DIM)
    rel_t = torch.einsum("bythwc,tkc->bythwk", r_q, Rt)
    rel_h = torch.einsum("bythwc,hkc->bythwk", r_q, Rh)
    rel_w = torch.einsum("bythwc,wkc->bythwk", r_q, Rw)
    rel = (rel_t[:, :, :, :, :, :, None, None] + rel_h[:, :, :, :, :, None, :, None] + rel_w[:, :, :, :, :, None, None, :])
    attn[:, :, 1:, 1:] = attn[:, :, 1:, 1:] + rel.reshape(B, heads, q_t * q_h * q_w, k_t * k_h * k_w)
    attn = attn.softmax(dim=-1)
    x = attn @ v
    x = torch.cat((x[:, :, :1], x[:, :, 1:] + q[:, :, 1:]), dim=2)          # residual pooling (:639-643)
    x = x.transpose(1, 2).reshape(B, -1, out_dims)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"]), q_shape


def block(sd, i, x, size):
    """MultiScaleBlock.forward (mvit.py:763-793), dim_mul_in_attention=True."""
    in_dims, out_dims, heads, stride_q, stride_kv = BLOCKS[i]
    p = "blocks.%d." % i
    x_norm = _ln(x, sd, p + "norm1")
    x_attn, out_size = attention(sd, p + "attn.", x_norm, size, out_dims, heads, stride_q, stride_kv)
    skip = F.linear(x_norm, sd[p + "proj.weight"], sd[p + "proj.bias"]) if in_dims != out_dims else x
    if stride_q > 1:                                  # pool_skip = MaxPool3d((1,3,3), (1,2,2), (0,1,1)) (:746-752)
        B, L, C = skip.shape
        T, H, W = size
        cls_tok, s = skip[:, :1], skip[:, 1:]
        s = s.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3).contiguous()
        s = F.max_pool3d(s, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        s = s.reshape(B, C, -1).transpose(1, 2)
        skip = torch.cat((cls_tok, s), dim=1)
    x = skip + x_attn
    x_norm = _ln(x, sd, p + "norm2")
    h = F.gelu(F.linear(x_norm, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x, out_size


def forward(sd, x):
    """x: [B, 3, 16, H, W] (or the raw 4-D view of it, mvit.py:1110-1111) -> [scale3, scale2, scale1, scale0] features
    [B,768,8,H/32,W/32], [B,384,8,H/16,W/16], [B,192,8,H/8,W/8], [B,96,8,H/4,W/4]."""
    if x.dim() == 4:
        x = x.view(-1, x.shape[-3], 16, x.shape[-2], x.shape[-1])
    B = x.shape[0]
    t = F.conv3d(x, sd["patch_embed.projection.weight"], sd["patch_embed.projection.bias"], stride=(2, 4, 4), padding=(1, 3, 3))
    size = tuple(t.shape[2:])
    t = t.flatten(2).transpose(1, 2)
    x = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    outs = []
    for i in range(len(BLOCKS)):
        x, size = block(sd, i, x, size)
        if i in STAGE_AFTER:
            x = _ln(x, sd, "norm%d" % STAGE_AFTER[i])
            outs.append(x.transpose(1, 2)[:, :, 1:].reshape(B, x.shape[-1], *size))
    return outs[::-1]
