"""GPU: every hand-written kernel through its C-ABI test entry against plain fp32 torch ops on the same
inputs (TF32 disabled).  Inputs of tensor-core kernels are pre-rounded to bf16 so that only the accumulation
order differs."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CONV_3X3, CONV_3X3_S2, CONV_1X1, CONV_TEMPORAL = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


@pytest.fixture(scope="module")
def L():
    from diff_sal_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def run_conv(L, kind, A_nhwc, Wt, N, F_, H, W, Cin, dilation=1, T=1, kt=1, scale=None, shift=None, rowbias=None,
             residual=None, act=0, want="f32", out_frames=None, out_fmul=1, out_fadd=0, head_w=None, head_b=0.0):
    lib = L.test_lib()
    out_frames = out_frames or F_
    o32 = torch.zeros(out_frames, H, W, N, device="cuda") if want == "f32" else None
    o16 = torch.zeros(out_frames, H, W, N, device="cuda", dtype=torch.bfloat16) if want == "bf16" else None
    oh = torch.zeros(out_frames, H, W, device="cuda") if head_w is not None else None
    r = lib.dsb_test_conv(kind, F_, H, W, Cin, N, dilation, T, kt, L.ptr(A_nhwc), L.ptr(Wt), L.ptr(scale),
                          L.ptr(shift), L.ptr(rowbias), L.ptr(residual), act, L.ptr(o32), L.ptr(o16), out_fmul,
                          out_fadd, L.ptr(head_w), ctypes.c_float(head_b), L.ptr(oh), L.stream_ptr())
    assert r == 0, "dsb_test_conv returned %d" % r
    torch.cuda.synchronize()
    return o32 if want == "f32" else (o16 if want == "bf16" else oh)


def pack_w(w):
    """[N, C, kh, kw] -> [N, (kh kw C)] bf16 (k = tap*Cin + c)."""
    n = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(n, -1).to(torch.bfloat16).contiguous()


def close(a, b, tol):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1e-6), "max err %g vs ref max %g" % (err, ref)


@pytest.mark.parametrize("M,Cin,N,act,resid", [(1000, 192, 384, ACT_GELU, False), (777, 96, 96, ACT_NONE, True),
                                               (130, 768, 1536, ACT_GELU, False), (64 * 1024, 96, 192, ACT_NONE, False)])
def test_linear(L, M, Cin, N, act, resid):
    a = _rand(M, Cin, seed=1).to(torch.bfloat16)
    w = _rand(N, Cin, seed=2, scale=Cin ** -0.5).to(torch.bfloat16)
    b = _rand(N, seed=3)
    res = _rand(M, N, seed=4) if resid else None
    out = run_conv(L, CONV_1X1, a, w, N, 1, 1, M, Cin, shift=b, residual=res, act=act, want="f32")
    ref = F.linear(a.float(), w.float(), b)
    if act == ACT_GELU:
        ref = F.gelu(ref)
    if resid:
        ref = ref + res
    close(out.reshape(M, N), ref, 2e-3)


@pytest.mark.parametrize("halo,two", [(0, -1), (1, 1)])
def test_fp16_operands(L, halo, two):
    """GemmParams::ab_f16: the output-head GEMMs (ReduceTemp, mt_proj) take fp16 instead of bf16 operands -- same
    kind::f16 instruction, A / B format bits of the instruction descriptor cleared.  Values are chosen so that an fp16
    buffer misread as bf16 (or the reverse) is wrong by orders of magnitude."""
    lib = L.test_lib()
    lib.dsb_test_set_ab_f16(1)
    lib.dsb_test_set_halo(halo)
    lib.dsb_test_set_two_cta(two)
    try:
        Fr, H, W, Cin, N = 2, 32, 48, 128, 96
        x = _rand(Fr, Cin, H, W, seed=71).to(torch.float16)
        w = _rand(N, Cin, 3, 3, seed=72, scale=(9 * Cin) ** -0.5).to(torch.float16)
        a = x.permute(0, 2, 3, 1).contiguous()
        wt = w.permute(0, 2, 3, 1).reshape(N, -1).contiguous()
        out = run_conv(L, CONV_3X3, a, wt, N, Fr, H, W, Cin, act=ACT_RELU, want="f32")
        ref = F.relu(F.conv2d(x.float(), w.float(), padding=1)).permute(0, 2, 3, 1)
        close(out, ref, 1e-3)
        # temporal (5,1,1) reduction, the other fp16 GEMM of the plan
        T, kt, C2 = 9, 5, 192
        xt = _rand(2 * T, 10, 12, C2, seed=73).to(torch.float16)
        w3 = _rand(768, C2, kt, seed=74, scale=(kt * C2) ** -0.5).to(torch.float16)
        wt3 = w3.permute(0, 2, 1).reshape(768, -1).contiguous()
        o = run_conv(L, CONV_TEMPORAL, xt, wt3, 768, 2, 10, 12, C2, T=T, kt=kt, act=ACT_RELU, want="f32")
        xr = xt.float().reshape(2, T, 10, 12, C2)[:, :kt]
        ref3 = F.relu(torch.einsum("bthwc,nct->bhwn", xr, w3.float()))
        close(o, ref3, 1e-3)
    finally:
        lib.dsb_test_set_ab_f16(0)
        lib.dsb_test_set_halo(0)
        lib.dsb_test_set_two_cta(0)


@pytest.mark.parametrize("Fr,H,W,Cin,N,dil", [(2, 28, 48, 192, 384, 1), (3, 14, 24, 768, 384, 2), (1, 56, 96, 96, 192, 1),
                                              (9, 56, 96, 96, 96, 2), (4, 7, 12, 768, 768, 1)])
def test_conv3x3(L, Fr, H, W, Cin, N, dil):
    x = _rand(Fr, Cin, H, W, seed=5).to(torch.bfloat16)
    w = _rand(N, Cin, 3, 3, seed=6, scale=(9 * Cin) ** -0.5).to(torch.bfloat16)
    scale = 1.0 + 0.1 * _rand(N, seed=7)
    shift = _rand(N, seed=8)
    rowbias = _rand(Fr, N, seed=9)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, Cin, dilation=dil, scale=scale, shift=shift,
                   rowbias=rowbias, act=ACT_RELU, want="bf16")
    ref = F.conv2d(x.float(), w.float(), None, padding=dil, dilation=dil)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None] + rowbias[:, :, None, None])
    close(out.permute(0, 3, 1, 2), ref, 1e-2)


@pytest.mark.parametrize("Fr,H,W,C", [(2, 14, 24, 384), (3, 28, 48, 192), (2, 7, 12, 768)])
def test_conv3x3_stride2_into_frame_slot(L, Fr, H, W, C):
    x = _rand(Fr, C, 2 * H, 2 * W, seed=10).to(torch.bfloat16)
    w = _rand(C, C, 3, 3, seed=11, scale=(9 * C) ** -0.5).to(torch.bfloat16)
    b = _rand(C, seed=12)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3_S2, a, pack_w(w), C, Fr, H, W, C, shift=b, want="f32", out_frames=Fr * 9,
                   out_fmul=9, out_fadd=8)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b, stride=2)
    out = out.reshape(Fr, 9, H, W, C)
    close(out[:, 8].permute(0, 3, 1, 2), ref, 2e-3)
    assert out[:, :8].abs().max().item() == 0.0


@pytest.mark.parametrize("B,H,W,C", [(2, 7, 12, 768), (3, 28, 48, 192), (1, 56, 96, 96)])
def test_temporal_reduce(L, B, H, W, C):
    x = _rand(B, C, 9, H, W, seed=13).to(torch.bfloat16)
    w = _rand(768, C, 5, 1, 1, seed=14, scale=(5 * C) ** -0.5).to(torch.bfloat16)
    a = x.permute(0, 2, 3, 4, 1).contiguous()                       # [B,T,H,W,C]
    wt = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(768, 5 * C).contiguous()
    out = run_conv(L, CONV_TEMPORAL, a, wt, 768, B, H, W, C, T=9, kt=5, act=ACT_RELU, want="f32")
    ref = F.relu(F.conv3d(x.float(), w.float(), None, stride=(5, 1, 1))).squeeze(2)
    close(out.permute(0, 3, 1, 2), ref, 2e-3)


def test_head_epilogue(L):
    Fr, H, W, Cin, N = 2, 16, 64, 768, 96
    x = _rand(Fr, Cin, H, W, seed=15).to(torch.bfloat16)
    w = _rand(N, Cin, 3, 3, seed=16, scale=(9 * Cin) ** -0.5).to(torch.bfloat16)
    scale = 1.0 + 0.1 * _rand(N, seed=17)
    shift = _rand(N, seed=18)
    hw = _rand(N, seed=19, scale=0.3)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, Cin, scale=scale, shift=shift, act=ACT_RELU,
                   want="head", head_w=hw, head_b=0.25)
    ref = F.relu(F.conv2d(x.float(), w.float(), None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    ref = torch.sigmoid((ref * hw[None, :, None, None]).sum(1) + 0.25)
    close(out, ref, 2e-3)


@pytest.fixture
def two_cta(L):
    lib = L.test_lib()
    lib.dsb_test_set_two_cta(1)
    yield
    lib.dsb_test_set_two_cta(0)


@pytest.mark.parametrize("Fr,H,W,Cin,N,dil", [(2, 28, 48, 192, 384, 1), (3, 14, 24, 768, 384, 2), (1, 56, 96, 96, 192, 1),
                                              (9, 56, 96, 96, 96, 2), (5, 7, 12, 768, 768, 1), (1, 28, 48, 384, 768, 1)])
def test_conv3x3_cta_pairs(L, two_cta, Fr, H, W, Cin, N, dil):
    """Same contraction through the cta_group::2 path (cluster of 2 CTAs, M = 256 MMA, B split across the pair),
    including odd M-tile counts (the pair's second tile is a dummy)."""
    test_conv3x3(L, Fr, H, W, Cin, N, dil)


def test_linear_and_head_cta_pairs(L, two_cta):
    test_linear(L, 1000, 192, 384, ACT_GELU, False)
    test_linear(L, 777, 96, 96, ACT_NONE, True)
    test_linear(L, 64 * 1024, 96, 192, ACT_NONE, False)
    test_head_epilogue(L)
    test_temporal_reduce(L, 3, 28, 48, 192)
    test_conv3x3_stride2_into_frame_slot(L, 3, 28, 48, 192)


@pytest.mark.parametrize("C,HW,F_,fg,fu", [(192, 1344, 3, 0, 0), (96, 5376, 2, 0, 0), (96, 640, 10, 9, 5), (192, 200, 4, 0, 0)])
def test_fused_mlp(L, C, HW, F_, fg, fu):
    """fc1 -> exact-erf GELU -> fc2 -> + residual with the hidden activation kept on chip (TMEM -> smem operand)."""
    src_frames = F_ if not fg else (F_ // fu) * fg
    a = _rand(src_frames, HW, C, seed=1).to(torch.bfloat16)
    w1 = _rand(2 * C, C, seed=2, scale=C ** -0.5).to(torch.bfloat16)
    w2 = _rand(C, 2 * C, seed=3, scale=(2 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = _rand(2 * C, seed=4, scale=0.2), _rand(C, seed=5, scale=0.2)
    res = _rand(src_frames, HW, C, seed=6)
    out = torch.full((src_frames, HW, C), 7.0, device="cuda")
    r = L.test_lib().dsb_test_mlp_fused(C, HW, F_, fg, fu, L.ptr(a), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(res),
                                   L.ptr(out), L.stream_ptr())
    assert r == 0, r
    torch.cuda.synchronize()
    hid = F.gelu(F.linear(a.float(), w1.float(), b1)).to(torch.bfloat16).float()     # the operand of fc2 is bf16
    ref = F.linear(hid, w2.float(), b2) + res
    live = [f for f in range(src_frames) if not fg or f % fg < fu]
    dead = [f for f in range(src_frames) if f not in live]
    close(out[live], ref[live], 3e-3)
    if dead:
        assert (out[dead] == 7.0).all()


@pytest.fixture
def split_ws(L):
    lib = L.test_lib()
    ws = torch.zeros(148 * 128 * 256, device="cuda")
    lib.dsb_test_set_split_ws.argtypes = [ctypes.c_void_p, ctypes.c_long]
    lib.dsb_test_set_split_ws(L.ptr(ws), ws.numel())
    yield lib
    lib.dsb_test_set_split_ws(None, 0)


@pytest.mark.parametrize("Fr,H,W,C,N", [(8, 7, 12, 768, 768), (8, 14, 24, 384, 384), (2, 14, 24, 384, 768)])
def test_splitk_conv_matches_unsplit(L, split_ws, Fr, H, W, C, N):
    """Tile-poor, K-long convs (the deep encoder layers at small batch) take the split-K path; same epilogue
    semantics (bias + row bias + residual, second output slot) and bitwise repeatable."""
    x = _rand(Fr, C, H, W, seed=30).to(torch.bfloat16)
    w = _rand(N, C, 3, 3, seed=31, scale=(9 * C) ** -0.5).to(torch.bfloat16)
    b, rb, res = _rand(N, seed=32), _rand(Fr, N, seed=33), _rand(Fr, H, W, N, seed=34)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, C, shift=b, rowbias=rb, residual=res, want="f32")
    assert split_ws.dsb_test_last_ksplit() > 1
    out2 = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, C, shift=b, rowbias=rb, residual=res, want="f32")
    assert torch.equal(out, out2)
    ref = F.conv2d(x.float(), w.float(), b, padding=1) + rb[:, :, None, None] + res.permute(0, 3, 1, 2)
    close(out.permute(0, 3, 1, 2), ref, 2e-3)


def test_splitk_temporal_reduce_and_stride2(L, split_ws):
    B, H, W, C = 8, 7, 12, 768
    x = _rand(B, C, 9, H, W, seed=35).to(torch.bfloat16)
    w = _rand(768, C, 5, 1, 1, seed=36, scale=(5 * C) ** -0.5).to(torch.bfloat16)
    a = x.permute(0, 2, 3, 4, 1).contiguous()
    wt = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(768, 5 * C).contiguous()
    out = run_conv(L, CONV_TEMPORAL, a, wt, 768, B, H, W, C, T=9, kt=5, act=ACT_RELU, want="f32")
    assert split_ws.dsb_test_last_ksplit() > 1
    ref = F.relu(F.conv3d(x.float(), w.float(), None, stride=(5, 1, 1))).squeeze(2)
    close(out.permute(0, 3, 1, 2), ref, 2e-3)
    # stride-2 down conv writing frame slot 8 of 9
    Fr, H, W, C = 8, 7, 12, 768
    x = _rand(Fr, C, 2 * H, 2 * W, seed=37).to(torch.bfloat16)
    w = _rand(C, C, 3, 3, seed=38, scale=(9 * C) ** -0.5).to(torch.bfloat16)
    b = _rand(C, seed=39)
    out = run_conv(L, CONV_3X3_S2, x.permute(0, 2, 3, 1).contiguous(), pack_w(w), C, Fr, H, W, C, shift=b, want="f32",
                   out_frames=Fr * 9, out_fmul=9, out_fadd=8)
    assert split_ws.dsb_test_last_ksplit() > 1
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b, stride=2)
    out = out.reshape(Fr, 9, H, W, C)
    close(out[:, 8].permute(0, 3, 1, 2), ref, 2e-3)
    assert out[:, :8].abs().max().item() == 0.0


def test_umma_row_shifted_descriptor(L):
    """Hardware-semantics probe behind the halo-tile convolution: a SWIZZLE_128B K-major operand descriptor
    may start at any 128-byte row of a TMA-written tile and space its 8-row groups by any multiple of 128 bytes (the
    swizzle is a function of the shared-memory address; the descriptor's base-offset field must stay 0)."""
    lib = L.test_lib()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(512, 64, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(32, 64, generator=g).to(torch.bfloat16).cuda()
    m = torch.arange(128)
    for sbo in (1024, 1280, 1536):
        for shift in (0, 1, 3, 8, 13):
            out = torch.zeros(128, 32, device="cuda")
            assert lib.dsb_test_umma_shift(L.ptr(A), L.ptr(B), L.ptr(out), shift, sbo, 0, L.stream_ptr()) == 0
            torch.cuda.synchronize()
            rows = ((m // 8) * (sbo // 128) + m % 8 + shift).cuda()
            ref = A[rows].float() @ B.float().t()
            assert (out - ref).abs().max().item() < 1e-3, (sbo, shift)


@pytest.fixture
def halo(L):
    lib = L.test_lib()
    lib.dsb_test_set_halo(1)
    yield lib
    lib.dsb_test_set_halo(0)


@pytest.mark.parametrize("Fr,H,W,Cin,N,dil", [(8, 56, 96, 192, 96, 2), (4, 56, 96, 64, 96, 1), (3, 112, 192, 128, 96, 1),
                                              (5, 60, 92, 128, 64, 2), (8, 56, 96, 96, 96, 2), (8, 56, 96, 96, 192, 1), (9, 28, 48, 192, 192, 2)])
def test_conv3x3_halo_tiles(L, halo, Fr, H, W, Cin, N, dil):
    """Halo-tile path: one TMA box per (8x16 tile, 64-channel block) serves all nine taps as row-shifted descriptors.
    Same results as the per-tap path (bitwise: same products, same K order per output) and as torch."""
    x = _rand(Fr, Cin, H, W, seed=40).to(torch.bfloat16)
    w = _rand(N, Cin, 3, 3, seed=41, scale=(9 * Cin) ** -0.5).to(torch.bfloat16)
    scale, shift = 1.0 + 0.1 * _rand(N, seed=42), _rand(N, seed=43)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, Cin, dilation=dil, scale=scale, shift=shift, act=ACT_RELU, want="bf16")
    assert halo.dsb_test_last_halo() == 1
    ref = F.relu(F.conv2d(x.float(), w.float(), None, padding=dil, dilation=dil) * scale[None, :, None, None] + shift[None, :, None, None])
    close(out.permute(0, 3, 1, 2), ref, 1e-2)
    halo.dsb_test_set_halo(0)
    plain = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, Cin, dilation=dil, scale=scale, shift=shift, act=ACT_RELU, want="bf16")
    assert halo.dsb_test_last_halo() == 0
    halo.dsb_test_set_halo(1)
    assert (out.float() - plain.float()).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_conv3x3_halo_head_epilogue(L, halo):
    Fr, H, W, Cin, N = 2, 112, 192, 256, 96
    x = _rand(Fr, Cin, H, W, seed=44).to(torch.bfloat16)
    w = _rand(N, Cin, 3, 3, seed=45, scale=(9 * Cin) ** -0.5).to(torch.bfloat16)
    scale, shift, hw = 1.0 + 0.1 * _rand(N, seed=46), _rand(N, seed=47), _rand(N, seed=48, scale=0.3)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = run_conv(L, CONV_3X3, a, pack_w(w), N, Fr, H, W, Cin, scale=scale, shift=shift, act=ACT_RELU, want="head", head_w=hw,
                   head_b=0.25)
    assert halo.dsb_test_last_halo() == 1
    ref = F.relu(F.conv2d(x.float(), w.float(), None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    ref = torch.sigmoid((ref * hw[None, :, None, None]).sum(1) + 0.25)
    close(out, ref, 2e-3)
