"""GPU: every memory-bound kernel through its single-kernel C-ABI entry against plain fp32 torch ops (the same ops
the reference uses) on the same inputs.  bf16 outputs are compared with a bf16-rounding tolerance."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = 2.0 ** -7          # bf16 rounding of an O(1) value, with margin


@pytest.fixture(scope="module")
def L():
    from diff_sal_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.test_lib()
    lib.dsb_test_layernorm.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def rnd(*shape, seed=0, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).cuda()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def check(rc):
    assert rc == 0, "kernel entry returned %d" % rc
    torch.cuda.synchronize()


@pytest.mark.parametrize("Fr,H,W,C", [(2, 56, 96, 96), (3, 28, 48, 192), (2, 14, 24, 768)])
def test_groupnorm_swish(L, Fr, H, W, C):
    x = rnd(Fr, C, H, W, seed=1, scale=1.5, shift=0.3)
    g, b = rnd(C, seed=2, scale=0.2, shift=1.0), rnd(C, seed=3, scale=0.1)
    a = nhwc(x)
    act = torch.empty(Fr, H, W, C, device="cuda", dtype=torch.bfloat16)
    raw = torch.empty_like(act)
    scratch = torch.zeros(Fr * 64 * 64, device="cuda", dtype=torch.float64)
    check(L.test_lib().dsb_test_groupnorm_swish(L.ptr(a), Fr, H * W, C, L.ptr(g), L.ptr(b), L.ptr(scratch), L.ptr(act),
                                           L.ptr(raw), L.stream_ptr()))
    ref = F.group_norm(x, 32, g, b, eps=1e-6)
    ref = nhwc(ref * torch.sigmoid(ref))
    assert (act.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())
    assert torch.equal(raw, a.to(torch.bfloat16))


@pytest.mark.parametrize("Fr,H,W,C", [(2, 56, 96, 96), (3, 56, 96, 192), (3, 28, 48, 192), (2, 28, 48, 384), (8, 14, 24, 384),
                                       (2, 14, 24, 768), (1, 7, 9, 96)])
def test_groupnorm_swish_one_launch(L, Fr, H, W, C):
    """gn_fused_kernel: a thread-block cluster per frame, statistics exchanged through distributed shared memory (every
    GroupNorm of the noise encoder runs through it); reproducible bit for bit run to run."""
    x = rnd(Fr, C, H, W, seed=41, scale=1.5, shift=0.3)
    g, b = rnd(C, seed=42, scale=0.2, shift=1.0), rnd(C, seed=43, scale=0.1)
    a = nhwc(x)
    act = torch.empty(Fr, H, W, C, device="cuda", dtype=torch.bfloat16)
    raw = torch.empty_like(act)
    check(L.test_lib().dsb_test_groupnorm_swish(L.ptr(a), Fr, H * W, C, L.ptr(g), L.ptr(b), None, L.ptr(act), L.ptr(raw), L.stream_ptr()))
    ref = F.group_norm(x, 32, g, b, eps=1e-6)
    ref = nhwc(ref * torch.sigmoid(ref))
    assert (act.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())
    assert torch.equal(raw, a.to(torch.bfloat16))
    act2 = torch.empty_like(act)
    check(L.test_lib().dsb_test_groupnorm_swish(L.ptr(a), Fr, H * W, C, L.ptr(g), L.ptr(b), None, L.ptr(act2), None, L.stream_ptr()))
    assert torch.equal(act, act2)


@pytest.mark.parametrize("C", [96, 192, 384, 768])
def test_layernorm_with_frame_filter(L, C):
    T, hw, B = 9, 20, 2
    x = rnd(B * T * hw, C, seed=4, scale=2.0, shift=0.5)
    g, b = rnd(C, seed=5, scale=0.2, shift=1.0), rnd(C, seed=6, scale=0.1)
    out = torch.full((B * T * hw, C), 7.0, device="cuda", dtype=torch.bfloat16)
    check(L.test_lib().dsb_test_layernorm(L.ptr(x), B * T * hw, C, L.ptr(g), L.ptr(b), L.ptr(out), hw, T, 5, L.stream_ptr()))
    ref = F.layer_norm(x, (C,), g, b, eps=1e-5)
    o = out.float().reshape(B, T, hw, C)
    r = ref.reshape(B, T, hw, C)
    assert (o[:, :5] - r[:, :5]).abs().max().item() <= BF * max(1.0, r.abs().max().item())
    assert (o[:, 5:] == 7.0).all()            # frames 5..8 untouched


@pytest.mark.parametrize("Fr,H,W,C", [(2, 7, 12, 768), (2, 14, 24, 384), (1, 28, 48, 192), (1, 56, 96, 96)])
def test_q_depthwise_layernorm(L, Fr, H, W, C):
    x = rnd(Fr, C, H, W, seed=7, scale=1.3)
    ng, nb = rnd(C, seed=8, scale=0.2, shift=1.0), rnd(C, seed=9, scale=0.1)
    w3 = rnd(C, 1, 3, 3, 3, seed=10, scale=0.3)
    qg, qb = rnd(C, seed=11, scale=0.2, shift=1.0), rnd(C, seed=12, scale=0.1)
    w9 = w3[:, 0, 1].reshape(C, 9).t().contiguous()                  # middle temporal tap, [9][C]
    out = torch.empty(Fr * H * W, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(Fr * H * W, 2, device="cuda")
    xl = nhwc(x)
    check(L.test_lib().dsb_test_q_dwln(L.ptr(xl), Fr, H, W, C, L.ptr(ng), L.ptr(nb), L.ptr(w9), L.ptr(qg), L.ptr(qb),
                                  L.ptr(stats), L.ptr(out), 1, 1, L.stream_ptr()))
    xn = F.layer_norm(nhwc(x), (C,), ng, nb, eps=1e-5).permute(0, 3, 1, 2)
    q = F.conv3d(xn.unsqueeze(2), w3, None, padding=1, groups=C).squeeze(2)     # the reference's Conv3d on depth 1
    ref = F.layer_norm(nhwc(q), (C,), qg, qb, eps=1e-5).reshape(Fr * H * W, C)
    assert (out.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("Fr,T,tmax,H,W,C,s", [(2, 1, 1, 56, 96, 96, 16), (3, 1, 1, 28, 48, 192, 8), (9, 9, 5, 28, 48, 192, 8)])
def test_fused_q_and_v_producer(L, Fr, T, tmax, H, W, C, s):
    """qv_tile_kernel (narrow stages, audio-visual): q = LN(dw3x3(LN(x))) and the 18 pooled v = LN(dw sxs(LN(x))) tokens
    of every live frame from one pass over x, LayerNorm statistics computed in the kernel."""
    x = rnd(Fr, C, H, W, seed=51, scale=1.3)
    ng, nb = rnd(C, seed=52, scale=0.2, shift=1.0), rnd(C, seed=53, scale=0.1)
    w3 = rnd(C, 1, 3, 3, 3, seed=54, scale=0.3)
    qg, qb = rnd(C, seed=55, scale=0.2, shift=1.0), rnd(C, seed=56, scale=0.1)
    wv = rnd(C, 1, 1, s, s, seed=57, scale=1.0 / s)
    vg, vb = rnd(C, seed=58, scale=0.2, shift=1.0), rnd(C, seed=59, scale=0.1)
    w9 = w3[:, 0, 1].reshape(C, 9).t().contiguous()
    wt = wv.reshape(C, s * s).t().contiguous()
    q_out = torch.full((Fr * H * W, C), 7.0, device="cuda", dtype=torch.bfloat16)
    v_out = torch.full((Fr * 18, C), 7.0, device="cuda", dtype=torch.bfloat16)
    xl = nhwc(x)
    check(L.test_lib().dsb_test_qv_tile(L.ptr(xl), Fr, H, W, C, s, L.ptr(ng), L.ptr(nb), L.ptr(w9), L.ptr(qg), L.ptr(qb), L.ptr(wt),
                                   L.ptr(vg), L.ptr(vb), L.ptr(q_out), L.ptr(v_out), T, tmax, L.stream_ptr()))
    xn = F.layer_norm(nhwc(x), (C,), ng, nb, eps=1e-5).permute(0, 3, 1, 2)
    q = F.conv3d(xn.unsqueeze(2), w3, None, padding=1, groups=C).squeeze(2)
    qref = F.layer_norm(nhwc(q), (C,), qg, qb, eps=1e-5).reshape(Fr, H * W, C)
    v = F.conv2d(xn, wv[:, :, 0], None, stride=s, groups=C)
    vref = F.layer_norm(nhwc(v), (C,), vg, vb, eps=1e-5).reshape(Fr, 18, C)
    qo, vo = q_out.float().reshape(Fr, H * W, C), v_out.float().reshape(Fr, 18, C)
    live = [f for f in range(Fr) if f % T < tmax]
    dead = [f for f in range(Fr) if f % T >= tmax]
    assert (qo[live] - qref[live]).abs().max().item() <= BF * max(1.0, qref.abs().max().item())
    assert (vo[live] - vref[live]).abs().max().item() <= BF * max(1.0, vref.abs().max().item())
    if dead:
        assert (qo[dead] == 7.0).all() and (vo[dead] == 7.0).all()          # dead frames untouched


@pytest.mark.parametrize("Fr,H,W,C,s", [(2, 7, 12, 768, 2), (2, 14, 24, 384, 4), (2, 28, 48, 192, 8), (1, 56, 96, 96, 16)])
def test_pooled_tokens(L, Fr, H, W, C, s):
    x = rnd(Fr, C, H, W, seed=13, scale=1.3)
    ng, nb = rnd(C, seed=14, scale=0.2, shift=1.0), rnd(C, seed=15, scale=0.1)
    wv = rnd(C, 1, 1, s, s, seed=16, scale=1.0 / s)
    vg, vb = rnd(C, seed=17, scale=0.2, shift=1.0), rnd(C, seed=18, scale=0.1)
    wt = wv.reshape(C, s * s).t().contiguous()
    out = torch.empty(Fr * 18, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(Fr * H * W, 2, device="cuda")
    xl = nhwc(x)
    check(L.test_lib().dsb_test_pool_ln(L.ptr(xl), Fr, H, W, C, s, L.ptr(ng), L.ptr(nb), L.ptr(wt), L.ptr(vg), L.ptr(vb),
                                   L.ptr(stats), L.ptr(out), 1, 1, L.stream_ptr()))
    xn = F.layer_norm(nhwc(x), (C,), ng, nb, eps=1e-5).permute(0, 3, 1, 2)
    v = F.conv2d(xn, wv[:, :, 0], None, stride=s, groups=C)
    ref = F.layer_norm(nhwc(v), (C,), vg, vb, eps=1e-5).reshape(Fr * 18, C)
    assert (out.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,C,s", [(2, 7, 12, 768, 2), (1, 14, 24, 384, 4), (2, 28, 48, 192, 8), (1, 56, 96, 96, 16)])
def test_audio_gate_and_scrambled_key(L, B, H, W, C, s):
    """transformer.py:128-147 + attention.py:89-91: gate, the raw .view reinterpretation, pooling, LayerNorm."""
    T = 9
    x5 = rnd(B, C, T, H, W, seed=19)
    a = rnd(B * T, C, 7, 12, seed=20)
    wk = rnd(C, 1, 1, s, s, seed=21, scale=1.0 / s)
    kg, kb = rnd(C, seed=22, scale=0.2, shift=1.0), rnd(C, seed=23, scale=0.1)
    xf = x5.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, C).contiguous()
    a_low = nhwc(a)
    gate = torch.empty(B, C, H, W, device="cuda")
    out = torch.empty(B * T * 18, C, device="cuda", dtype=torch.bfloat16)
    wkt = wk.reshape(C, s * s).t().contiguous()
    check(L.test_lib().dsb_test_av_key(L.ptr(xf), L.ptr(a_low), B, T, H, W, C, s, L.ptr(wkt), L.ptr(kg), L.ptr(kb), L.ptr(gate),
                                  L.ptr(out), T, L.stream_ptr()))
    au = a if H == 7 else F.interpolate(a, scale_factor=H // 7, mode="nearest")
    a5 = au.reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4)
    g = torch.softmax((a5 * x5).mean(dim=2, keepdim=True), dim=-1)
    assert (gate - g[:, :, 0]).abs().max().item() < 1e-5
    src = (a5 * g).contiguous().view(B * T, H * W, C).permute(0, 2, 1).reshape(B * T, C, H, W)
    k = F.conv2d(src, wk[:, :, 0], None, stride=s, groups=C)
    ref = F.layer_norm(nhwc(k), (C,), kg, kb, eps=1e-5).reshape(B * T * 18, C)
    assert (out.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("Fr,H,W,C", [(2, 7, 12, 768), (3, 14, 24, 384), (2, 28, 48, 192)])
def test_upsample2x(L, Fr, H, W, C):
    x = rnd(Fr, C, H, W, seed=24)
    out = torch.empty(Fr, 2 * H, 2 * W, C, device="cuda", dtype=torch.bfloat16)
    xl = nhwc(x)
    check(L.test_lib().dsb_test_upsample2x(L.ptr(xl), Fr, H, W, C, L.ptr(out), L.stream_ptr()))
    ref = nhwc(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    assert (out.float() - ref).abs().max().item() <= BF * max(1.0, ref.abs().max().item())


def test_multi_scale_sum_and_final_upsample(L):
    B = 2
    rs = [rnd(B, 768, 7 << k, 12 << k, seed=25 + k) for k in range(4)]
    out = torch.empty(B, 112, 192, 768, device="cuda", dtype=torch.float16)     # fp16: operand of the mt_proj GEMM
    rl = [nhwc(r) for r in rs]                      # keep the channels-last copies alive across the launch
    check(L.test_lib().dsb_test_ms_sum(*[L.ptr(r) for r in rl], B, L.ptr(out), L.stream_ptr()))
    ref = sum(F.interpolate(r, size=(112, 192), mode="bilinear", align_corners=False) for r in rs)
    assert (out.float() - nhwc(ref)).abs().max().item() <= BF / 8 * max(1.0, ref.abs().max().item())
    p = rnd(B, 1, 112, 192, seed=30)
    o2 = torch.empty(B, 1, 224, 384, device="cuda")
    check(L.test_lib().dsb_test_final_up(L.ptr(p), B, L.ptr(o2), L.stream_ptr()))
    assert (o2 - F.interpolate(p, size=(224, 384), mode="bilinear", align_corners=False)).abs().max().item() < 1e-6


def test_stem_composition(L):
    """conv_in (3x3 pad 1) followed by Downsample4x4 (pad right/bottom, 3x3 stride 4) as ONE 5x5 stride-4 conv."""
    B = 2
    x = rnd(B, 1, 224, 384, seed=31)
    w_in, b_in = rnd(96, 1, 3, 3, seed=32, scale=0.3), rnd(96, seed=33, scale=0.1)
    w_d, b_d = rnd(96, 96, 3, 3, seed=34, scale=0.05), rnd(96, seed=35, scale=0.1)
    w5, b5 = torch.empty(2400, device="cuda"), torch.empty(96, device="cuda")
    h0 = torch.empty(B, 56, 96, 96, device="cuda")
    check(L.test_lib().dsb_test_stem(L.ptr(x), B, L.ptr(w_in), L.ptr(b_in), L.ptr(w_d), L.ptr(b_d), L.ptr(w5), L.ptr(b5),
                                L.ptr(h0), L.stream_ptr()))
    ref = F.conv2d(F.pad(F.conv2d(x, w_in, b_in, padding=1), (0, 1, 0, 1)), w_d, b_d, stride=4)
    assert (h0 - nhwc(ref)).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("tvals", [[0.0, 999.0, 37.0], [886.9, 0.5]])
def test_timestep_embedding_mlp(L, tvals):
    B = len(tvals)
    t = torch.tensor(tvals).cuda()
    w0, b0 = rnd(384, 96, seed=36, scale=0.1), rnd(384, seed=37, scale=0.1)
    w1, b1 = rnd(384, 384, seed=38, scale=0.05), rnd(384, seed=39, scale=0.1)
    wps = [rnd(c, 384, seed=40 + i, scale=0.05) for i, c in enumerate((192, 384, 768))]
    bps = [rnd(c, seed=43 + i, scale=0.1) for i, c in enumerate((192, 384, 768))]
    outs = [torch.empty(B, c, device="cuda") for c in (192, 384, 768)]
    w0t, w1t = w0.t().contiguous(), w1.t().contiguous()          # [in][out], kept alive across the launch
    wpt = [w.t().contiguous() for w in wps]
    check(L.test_lib().dsb_test_temb(L.ptr(t), B, L.ptr(w0t), L.ptr(b0), L.ptr(w1t), L.ptr(b1), L.ptr(wpt[0]), L.ptr(bps[0]),
                                L.ptr(wpt[1]), L.ptr(bps[1]), L.ptr(wpt[2]), L.ptr(bps[2]), L.ptr(outs[0]), L.ptr(outs[1]),
                                L.ptr(outs[2]), L.stream_ptr()))
    half = 48
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1))).cuda()
    e = torch.cat([torch.sin(t[:, None] * freq), torch.cos(t[:, None] * freq)], dim=1)
    h = F.linear(e, w0, b0)
    h = F.linear(h * torch.sigmoid(h), w1, b1)
    h = h * torch.sigmoid(h)
    for o, wp, bp in zip(outs, wps, bps):
        ref = F.linear(h, wp, bp)
        assert (o - ref).abs().max().item() <= 3e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("C,Fr,R", [(96, 3, 64), (192, 2, 64), (384, 2, 48), (768, 1, 48)])
def test_projections_folded_into_kv(L, C, Fr, R):
    """attention.py:97-113 with proj_q folded into proj_k and proj into proj_v at the WEIGHT level (plan.cu kFoldProj):
    the K / V projection GEMMs emit the score operand K'[f][h*18+j][c], the P.V operand V''[f][c][h*18+j] and the K pooling
    kernel the folded score bias; all three against the unfused fp32 formulas."""
    wq, wk_, wv, wp = (rnd(C, C, seed=70 + i, scale=C ** -0.5) for i in range(4))
    bq, bk, bv = (rnd(C, seed=80 + i, scale=0.2) for i in range(3))
    k_ln = rnd(Fr * 18, C, seed=90).to(torch.bfloat16)
    v_ln = rnd(Fr * 18, C, seed=91).to(torch.bfloat16)
    K1 = torch.zeros(Fr, R, C, device="cuda", dtype=torch.bfloat16)
    V2 = torch.zeros(Fr, C, 64, device="cuda", dtype=torch.bfloat16)
    mb = torch.zeros(2, C, device="cuda")
    cb = torch.zeros(2, device="cuda")
    check(L.test_lib().dsb_test_fold_kv(L.ptr(wq), L.ptr(bq), L.ptr(wk_), L.ptr(bk), L.ptr(wp), L.ptr(wv), L.ptr(bv), C, Fr, R,
                                        L.ptr(k_ln), L.ptr(v_ln), L.ptr(K1), L.ptr(V2), L.ptr(mb), L.ptr(cb), L.stream_ptr()))
    torch.cuda.synchronize()
    scale, d = C ** -0.5, C // 2
    K = (k_ln.float() @ wk_.t() + bk).reshape(Fr, 18, C)
    V = (v_ln.float() @ wv.t() + bv).reshape(Fr, 18, C)
    K1_ref = torch.zeros(Fr, R, C, device="cuda")
    V2_ref = torch.zeros(Fr, C, 64, device="cuda")
    sb_ref = torch.zeros(Fr, R, device="cuda")
    for h in range(2):
        sl = slice(h * d, (h + 1) * d)
        K1_ref[:, h * 18:(h + 1) * 18] = scale * K[:, :, sl] @ wq[sl, :]
        sb_ref[:, h * 18:(h + 1) * 18] = scale * K[:, :, sl] @ bq[sl]
        V2_ref[:, :, h * 18:(h + 1) * 18] = (V[:, :, sl] @ wp[:, sl].t()).transpose(1, 2)
    tol = 3 * BF
    assert (K1.float() - K1_ref).abs().max().item() <= tol * max(1.0, K1_ref.abs().max().item())
    assert (V2.float() - V2_ref).abs().max().item() <= tol * max(1.0, V2_ref.abs().max().item())
    assert K1[:, 36:].abs().max().item() == 0 and V2[:, :, 36:].abs().max().item() == 0      # padding untouched
    # folded score bias: sb[f][h*18+j] = k_ln[j] . mb[h] + cb[h] must equal scale * K_h[j] . bq_h
    sb_from_tables = torch.zeros(Fr, R, device="cuda")
    kl = k_ln.float().reshape(Fr, 18, C)
    for h in range(2):
        sb_from_tables[:, h * 18:(h + 1) * 18] = kl @ mb[h] + cb[h]
    assert (sb_from_tables - sb_ref).abs().max().item() <= 1e-4 * max(1.0, sb_ref.abs().max().item())


def test_score_bias_from_pooling_kernel(L):
    """the K pooling kernels write sb[f][h*18+j] = k_ln[f,j] . mb[h] + cb[h] beside the pooled tokens (kernels.cu ScoreBias)"""
    Fr, H, W, C, s, R = 2, 28, 48, 192, 8, 64
    x = rnd(Fr, C, H, W, seed=13, scale=1.3)
    ng, nb = rnd(C, seed=14, scale=0.2, shift=1.0), rnd(C, seed=15, scale=0.1)
    wv = rnd(C, 1, 1, s, s, seed=16, scale=1.0 / s)
    vg, vb = rnd(C, seed=17, scale=0.2, shift=1.0), rnd(C, seed=18, scale=0.1)
    mb, cb = rnd(2, C, seed=19, scale=0.1), rnd(2, seed=20)
    wt = wv.reshape(C, s * s).t().contiguous()
    out = torch.empty(Fr * 18, C, device="cuda", dtype=torch.bfloat16)
    sb = torch.zeros(Fr, R, device="cuda")
    stats = torch.empty(Fr * H * W, 2, device="cuda")
    check(L.test_lib().dsb_test_pool_ln_sb(L.ptr(nhwc(x)), Fr, H, W, C, s, L.ptr(ng), L.ptr(nb), L.ptr(wt), L.ptr(vg), L.ptr(vb),
                                           L.ptr(stats), L.ptr(out), 1, 1, L.ptr(mb), L.ptr(cb), L.ptr(sb), R, L.stream_ptr()))
    torch.cuda.synchronize()
    tok = out.float().reshape(Fr, 18, C)
    for h in range(2):
        ref = tok @ mb[h] + cb[h]
        assert (sb[:, h * 18:(h + 1) * 18] - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    assert sb[:, 36:].abs().max().item() == 0
