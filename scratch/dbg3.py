import torch, sys, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch.nn.functional as F
from diff_sal_b200 import _lib as L
from test_kernels_gpu import run_conv, pack_w, _rand, CONV_3X3, CONV_3X3_S2, CONV_1X1
def check(name, fn, n=30):
    ref = fn().clone(); bad = 0; mx = 0
    for i in range(n):
        o = fn()
        if not torch.equal(o, ref):
            bad += 1; mx = max(mx, (o.float()-ref.float()).abs().max().item())
    print(name, "nondeterministic runs: %d/%d maxdiff %.3e" % (bad, n, mx))
# block-1 shapes
x = _rand(1, 192, 28, 48, seed=5).to(torch.bfloat16); a = x.permute(0,2,3,1).contiguous()
w = pack_w(_rand(384, 192, 3, 3, seed=6, scale=0.02).to(torch.bfloat16))
sh = _rand(384, seed=8); rb = _rand(1, 384, seed=9)
check("conv1 192->384 28x48", lambda: run_conv(L, CONV_3X3, a, w, 384, 1, 28, 48, 192, shift=sh, rowbias=rb, want="f32"))
w1 = _rand(384, 192, seed=2, scale=0.05).to(torch.bfloat16)
check("shortcut 1x1 192->384", lambda: run_conv(L, CONV_1X1, a, w1, 384, 1, 28, 48, 192, shift=sh, want="f32"))
x2 = _rand(1, 384, 28, 48, seed=7).to(torch.bfloat16); a2 = x2.permute(0,2,3,1).contiguous()
w2 = pack_w(_rand(384, 384, 3, 3, seed=6, scale=0.02).to(torch.bfloat16))
res = _rand(1, 28, 48, 384, seed=11)
check("conv2 384->384 28x48 +res bf16", lambda: run_conv(L, CONV_3X3, a2, w2, 384, 1, 28, 48, 384, shift=sh, residual=res, want="bf16"))
check("down s2 384 -> 14x24", lambda: run_conv(L, CONV_3X3_S2, a2, w2, 384, 1, 14, 24, 384, shift=sh, want="f32"))
# block 0 shapes
x0 = _rand(1, 96, 56, 96, seed=5).to(torch.bfloat16); a0 = x0.permute(0,2,3,1).contiguous()
w0 = pack_w(_rand(192, 96, 3, 3, seed=6, scale=0.02).to(torch.bfloat16)); sh0=_rand(192, seed=8)
check("conv1 96->192 56x96", lambda: run_conv(L, CONV_3X3, a0, w0, 192, 1, 56, 96, 96, shift=sh0, want="f32"))
# big
x3 = _rand(9, 192, 56, 96, seed=5).to(torch.bfloat16); a3 = x3.permute(0,2,3,1).contiguous()
w3 = pack_w(_rand(96, 192, 3, 3, seed=6, scale=0.02).to(torch.bfloat16)); sh3=_rand(96, seed=8)
check("upembed 192->96 56x96 F=9 dil2", lambda: run_conv(L, CONV_3X3, a3, w3, 96, 9, 56, 96, 192, dilation=2, shift=sh3, act=1, want="bf16"), n=10)
