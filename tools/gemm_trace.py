# per-role clock trace of one GEMM-class launch (GemmParams::trace): where a tile's time goes -- operand arrival, MMA issue,
# accumulator completion, epilogue -- for the producer / MMA / epilogue roles of CTAs 0..3.
# usage: python tools/gemm_trace.py M Cin N [act] [bf16_out]                             (token linear)
#        python tools/gemm_trace.py conv F H W Cin N dilation [act] [bf16_out] [residual] [scale] [halo]   (3x3 conv)
import sys, ctypes, torch
sys.path.insert(0, '.')
from diff_sal_b200 import _lib as L

lib = L.test_lib()
g = torch.Generator().manual_seed(0)
a = sys.argv[1:]
if a[0] == "conv":
    F_, H, W, Cin, N, dil = [int(v) for v in a[1:7]]
    act = int(a[7]) if len(a) > 7 else 0
    want_bf16 = int(a[8]) if len(a) > 8 else 0
    resid = int(a[9]) if len(a) > 9 else 0
    use_scale = int(a[10]) if len(a) > 10 else 0
    halo = int(a[11]) if len(a) > 11 else 0
    dbg, kind, taps = 0, 0, 9
else:
    F_, H, W, Cin, N, dil = 1, 1, int(a[0]), int(a[1]), int(a[2]), 1
    act = int(a[3]) if len(a) > 3 else 0
    want_bf16 = int(a[4]) if len(a) > 4 else 0
    dbg = int(a[5]) if len(a) > 5 else 0
    resid, use_scale, halo, kind, taps = 0, 0, 0, 2, 1
M = F_ * H * W
x = torch.randn(M, Cin, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(N, taps * Cin, generator=g) * (taps * Cin) ** -0.5).to(torch.bfloat16).cuda()
b = torch.randn(N, generator=g).cuda()
sc = (torch.rand(N, generator=g) + 0.5).cuda() if use_scale else None
res = torch.randn(M, N, generator=g).cuda() if resid else None
o32 = torch.zeros(M, N, device="cuda") if not want_bf16 else None
o16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if want_bf16 else None
trace = torch.zeros(784, dtype=torch.int64, device="cuda")
if halo:
    lib.dsb_test_set_halo(1)

def run():
    r = lib.dsb_test_conv(kind, F_, H, W, Cin, N, dil, 1, 1, L.ptr(x), L.ptr(w), L.ptr(sc), L.ptr(b), None, L.ptr(res), act,
                          L.ptr(o32), L.ptr(o16), 1, 0, None, ctypes.c_float(0.0), None, L.stream_ptr())
    assert r == 0, r

for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / 20
print("M=%d K=%d N=%d act=%d: %.2f us per launch, %.1f TF/s" % (M, taps * Cin, N, act, us, 2.0 * M * taps * Cin * N / (us * 1e-6) / 1e12))
trace[780] = dbg
lib.dsb_test_set_gemm_trace(ctypes.c_void_p(trace.data_ptr()))
run(); torch.cuda.synchronize()
lib.dsb_test_set_gemm_trace(None)
t = trace.cpu().tolist()
for cta in range(2):
    k0, k1 = t[768 + cta * 2], t[768 + cta * 2 + 1]
    print("CTA %d: kernel body %d cycles" % (cta, k1 - k0))
    for u in range(16):
        pr = [t[((cta * 3 + 0) * 16 + u) * 4 + k] for k in range(2)]
        mm = [t[((cta * 3 + 1) * 16 + u) * 4 + k] for k in range(3)]
        ep = [t[((cta * 3 + 2) * 16 + u) * 4 + k] for k in range(3)]
        if not ep[0]: break
        f = lambda v: (v - k0) if v else -1
        print("  unit %2d | producer start %6d issued-all %6d | mma acc-free %6d first-stage %6d issued-all %6d | epi wait-from %6d acc-done %6d done %6d"
              % (u, f(pr[0]), f(pr[1]), f(mm[0]), f(mm[1]), f(mm[2]), f(ep[0]), f(ep[1]), f(ep[2])))
