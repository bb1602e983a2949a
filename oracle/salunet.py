"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of one
SalUNet denoiser evaluation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  The product path (diff_sal_b200/) never does.

Restates, as plain functional torch on a reference-keyed state_dict:
  SalUNet.forward            /root/reference/models/saliency_decoder/sal_unet.py:302-328
  noise_downsample           sal_unet.py:279-300  (Downsample4x4 :67-84, Downsample :47-64,
                             ResnetBlock :87-142, Normalize :41-44, temb :15-38)
  Decoder.forward            sal_unet.py:457-491
  TransformerStage.forward   transformer.py:259-289   (UpEmbed common_block.py:176-223)
  TransformerBlock.forward   transformer.py:124-159
  Attention.forward          attention.py:86-113
  Mlp / ReduceTemp / MLPHead common_block.py:111-173

Parity of this restatement is pinned against the imported reference in
tests/test_oracle_vs_reference.py (runs where /root/reference exists) and against
the committed fixtures tests/golden/*.npz (generated from the reference by
tests/golden/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F

STAGE_C = (768, 384, 192, 96)
STAGE_HW = ((7, 12), (14, 24), (28, 48), (56, 96))
STAGE_S = (2, 4, 8, 16)
N_FRAMES = 9          # 8 MViT temporal slices + 1 noise slice (sal_unet.py:317)
N_REDUCE = 5          # ReduceTemp kernel/stride (cfgs/audio_visual.py:62)
OUT_HW = (224, 384)
MID_HW = (112, 192)


def timestep_embedding(t, dim=96):
    """sal_unet.py:15-33 (half=48, f_k = exp(-k ln(1e4)/(half-1)), [sin, cos])."""
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32)
    freq = torch.exp(k * -(math.log(10000.0) / (half - 1)))
    arg = t.float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def swish(x):
    return x * torch.sigmoid(x)


def temb_mlp(sd, t):
    """sal_unet.py:304-307."""
    e = timestep_embedding(t, 96)
    e = F.linear(e, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    e = swish(e)
    return F.linear(e, sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def resnet_block(sd, p, x, temb):
    """sal_unet.py:123-142 (dropout is identity in eval)."""
    h = swish(_gn(x, sd, p + ".norm1"))
    h = F.conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    h = swish(_gn(h, sd, p + ".norm2"))
    h = F.conv2d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def pad_conv_down(x, w, b, stride):
    """sal_unet.py:57-61 / :77-81: zero-pad right/bottom by one, 3x3 conv, stride 2 or 4."""
    return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=stride)


def noise_encoder(sd, x, temb):
    """sal_unet.py:279-300 -> [768@7x12, 384@14x24, 192@28x48], each [B,C,1,h,w]."""
    h = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    h = pad_conv_down(h, sd["down1.conv.weight"], sd["down1.conv.bias"], 4)
    outs = []
    for i in range(3):
        h = resnet_block(sd, "res_encoder.%d.0" % i, h, temb)
        h = pad_conv_down(h, sd["res_encoder.%d.1.conv.weight" % i], sd["res_encoder.%d.1.conv.bias" % i], 2)
        outs.append(h.unsqueeze(2))
    return outs[::-1]


def _bn_eval(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=1e-5)


def up_embed(sd, p, x):
    """common_block.py:196-223 on frames [(B T),C,h,w]: x2 bilinear, two dilated convs + BN + ReLU."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    x = F.relu(_bn_eval(F.conv2d(x, sd[p + ".proj.1.weight"], None, padding=2, dilation=2), sd, p + ".proj.2"))
    x = F.relu(_bn_eval(F.conv2d(x, sd[p + ".proj.4.weight"], None, padding=2, dilation=2), sd, p + ".proj.5"))
    return x


def audio_key_source(sd, p, x5, audio):
    """transformer.py:128-147.  x5 [B,C,T,H,W]; audio [B,512,T,7,12].
    Returns the K source as frames [(B T), C, H, W] after the raw .view reinterpretation."""
    B, C, T, H, W = x5.shape
    a = audio.permute(0, 2, 1, 3, 4).reshape(B * T, 512, audio.shape[3], audio.shape[4])
    a = F.conv2d(a, sd[p + ".align_conv.weight"], sd[p + ".align_conv.bias"])
    h, w = a.shape[-2:]
    if h != H and w != W:
        a = F.interpolate(a, scale_factor=H // h, mode="nearest")
    a5 = a.reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4)
    gate = torch.softmax((a5 * x5).mean(dim=2, keepdim=True), dim=-1)
    a5 = (a5 * gate).contiguous()
    # raw reinterpretation of the [B,C,T,H,W] buffer as [(B T),(H W),C]  (transformer.py:146)
    tok = a5.view(B * T, H * W, C)
    return tok.permute(0, 2, 1).reshape(B * T, C, H, W)


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _tokens(img):
    n, c, h, w = img.shape
    return img.reshape(n, c, h * w).permute(0, 2, 1)


def attention(sd, p, xn_tok, H, W, s, k_src_img=None):
    """attention.py:86-113 with fea_no=1.  xn_tok [(B T), HW, C] (already LayerNormed)."""
    n, _, C = xn_tok.shape
    img = xn_tok.permute(0, 2, 1).reshape(n, C, H, W)
    wq = sd[p + ".conv_proj_q.conv.weight"][:, :, 1]        # depth-1 volume: only the middle temporal tap
    wk = sd[p + ".conv_proj_k.conv.weight"][:, :, 0]
    wv = sd[p + ".conv_proj_v.conv.weight"][:, :, 0]
    q = _ln(_tokens(F.conv2d(img, wq, None, padding=1, groups=C)), sd, p + ".conv_proj_q.bn")
    ksrc = img if k_src_img is None else k_src_img
    k = _ln(_tokens(F.conv2d(ksrc, wk, None, stride=s, groups=C)), sd, p + ".conv_proj_k.bn")
    v = _ln(_tokens(F.conv2d(img, wv, None, stride=s, groups=C)), sd, p + ".conv_proj_v.bn")
    q = F.linear(q, sd[p + ".proj_q.weight"], sd[p + ".proj_q.bias"])
    k = F.linear(k, sd[p + ".proj_k.weight"], sd[p + ".proj_k.bias"])
    v = F.linear(v, sd[p + ".proj_v.weight"], sd[p + ".proj_v.bias"])
    d = C // 2
    q = q.reshape(n, -1, 2, d).permute(0, 2, 1, 3)
    k = k.reshape(n, -1, 2, d).permute(0, 2, 1, 3)
    v = v.reshape(n, -1, 2, d).permute(0, 2, 1, 3)
    score = torch.matmul(q, k.transpose(-1, -2)) * (C ** -0.5)   # full-dim scale (attention.py:33)
    o = torch.matmul(torch.softmax(score, dim=-1), v)
    o = o.permute(0, 2, 1, 3).reshape(n, -1, C)
    return F.linear(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def transformer_block(sd, p, x5, s, audio):
    """transformer.py:124-159.  x5 [B,C,T,H,W] -> same."""
    B, C, T, H, W = x5.shape
    k_src = audio_key_source(sd, p, x5, audio) if audio is not None else None
    frames = x5.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    x = _tokens(frames)
    x = attention(sd, p + ".attn", _ln(x, sd, p + ".norm"), H, W, s, k_src) + x
    h = F.linear(_ln(x, sd, p + ".norm2"), sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
    h = F.gelu(h)
    x = x + F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x.permute(0, 2, 1).reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4)


def decoder(sd, back_fea, audio, taps=None):
    """sal_unet.py:457-491 + transformer.py:259-289.  back_fea[i]: [B,C_i,9,h_i,w_i] for i<3.
    ``taps`` (optional dict) receives intermediates in channels-last form for per-stage parity checks."""
    x5 = back_fea[0]
    B = x5.shape[0]
    acc = 0
    for i in range(4):
        p = "invpt_decoder.mid_stages.%d" % i
        if i > 0:
            C, T = x5.shape[1], x5.shape[2]
            fr = x5.permute(0, 2, 1, 3, 4).reshape(B * T, C, x5.shape[3], x5.shape[4])
            fr = up_embed(sd, p + ".patch_embed.0", fr)
            x5 = fr.reshape(B, T, fr.shape[1], fr.shape[2], fr.shape[3]).permute(0, 2, 1, 3, 4)
            if i in (1, 2):                      # stage 3 has no skip (transformer.py:265-270)
                x5 = x5 + back_fea[i]
        x5 = transformer_block(sd, p + ".blocks.0", x5, STAGE_S[i], audio)
        Bc, C, T, H, W = x5.shape
        tok = x5.permute(0, 2, 3, 4, 1)                         # [B,T,H,W,C]
        if taps is not None:
            taps["x%d" % i] = tok.contiguous()
        tok = _ln(tok, sd, "invpt_decoder.norm_mts.%d" % i)
        y = tok.permute(0, 4, 1, 2, 3)                          # [B,C,T,H,W]
        y = F.relu(F.conv3d(y, sd["invpt_decoder.redu_chan_up.%d.proj.0.weight" % i], None,
                            stride=(N_REDUCE, 1, 1)))
        y = y.squeeze(2)
        if taps is not None:
            taps["r%d" % i] = y.permute(0, 2, 3, 1).contiguous()
        acc = acc + F.interpolate(y, size=MID_HW, mode="bilinear", align_corners=False)
    y = F.conv2d(acc, sd["invpt_decoder.mt_proj.0.weight"], sd["invpt_decoder.mt_proj.0.bias"], padding=1)
    return F.relu(_bn_eval(y, sd, "invpt_decoder.mt_proj.1"))


def forward(sd, x, t, feat_list, audio=None, taps=None):
    """One denoiser evaluation, sal_unet.py:302-328.  Never mutates feat_list.
    x [B,1,224,384]; t [B] (int or float); feat_list 4 tensors [B,C_i,8,h_i,w_i];
    audio [B,512,9,7,12] or None.  Returns [B,1,224,384] in (0,1)."""
    with torch.no_grad():
        x = x.float()
        temb = temb_mlp(sd, t)
        noise = noise_encoder(sd, x, temb)
        back = [torch.cat([feat_list[i].float(), noise[i]], dim=2) for i in range(3)]
        y = decoder(sd, back, None if audio is None else audio.float(), taps)
        y = torch.sigmoid(F.conv2d(y, sd["logits.linear_pred.weight"], sd["logits.linear_pred.bias"]))
        if taps is not None:
            for i in range(3):
                taps["noise%d" % i] = noise[i].squeeze(2).permute(0, 2, 3, 1).contiguous()
            taps["p"] = y.squeeze(1).contiguous()
        return F.interpolate(y, size=OUT_HW, mode="bilinear", align_corners=False)
