# per-role clock trace of one GEMM-class launch (GemmParams::trace): where a tile's time goes -- operand arrival, MMA issue,
# accumulator completion, epilogue -- for the producer / MMA / epilogue roles of CTAs 0..3.
# usage: python tools/gemm_trace.py M Cin N [act] [kind taps: 1x1 only]
import sys, ctypes, torch
sys.path.insert(0, '.')
from diff_sal_b200 import _lib as L

M, Cin, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
want_bf16 = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dbg = int(sys.argv[6]) if len(sys.argv) > 6 else 0
lib = L.test_lib()
g = torch.Generator().manual_seed(0)
a = (torch.randn(M, Cin, generator=g)).to(torch.bfloat16).cuda()
w = (torch.randn(N, Cin, generator=g) * Cin ** -0.5).to(torch.bfloat16).cuda()
b = torch.randn(N, generator=g).cuda()
o32 = torch.zeros(M, N, device="cuda") if not want_bf16 else None
o16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if want_bf16 else None
trace = torch.zeros(784, dtype=torch.int64, device="cuda")

def run():
    r = lib.dsb_test_conv(2, 1, 1, M, Cin, N, 1, 1, 1, L.ptr(a), L.ptr(w), None, L.ptr(b), None, None, act, L.ptr(o32), L.ptr(o16),
                          1, 0, None, ctypes.c_float(0.0), None, L.stream_ptr())
    assert r == 0, r

for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("M=%d K=%d N=%d act=%d: %.2f us per launch, %.1f TF/s" % (M, Cin, N, act, 1e3 * e0.elapsed_time(e1) / 20, 2.0 * M * Cin * N / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e12))
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
