import torch, ctypes, sys
sys.path.insert(0, '.')
from diff_sal_b200 import _lib as L
lib = L.lib()
lib.dsb_test_set_chain_trace.argtypes = [ctypes.c_void_p]
def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed); return (torch.randn(*s, generator=g) * scale).cuda()
for C, HW, F_ in [(192, 1344, 72), (96, 5376, 40)]:
    a = rnd(F_, HW, C, seed=1).to(torch.bfloat16)
    w1 = rnd(2 * C, C, seed=2, scale=C ** -0.5).to(torch.bfloat16)
    w2 = rnd(C, 2 * C, seed=3, scale=(2 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = rnd(2 * C, seed=4, scale=0.2), rnd(C, seed=5, scale=0.2)
    res = rnd(F_, HW, C, seed=6)
    out = torch.zeros(F_, HW, C, device="cuda")
    trace = torch.zeros(24 * 8, dtype=torch.int64, device="cuda")
    args = (C, HW, F_, 0, 0, L.ptr(a), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(res), L.ptr(out), L.stream_ptr())
    for _ in range(2): lib.dsb_test_mlp_fused(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.dsb_test_mlp_fused(*args); e1.record(); torch.cuda.synchronize()
    print("C=%d HW=%d F=%d: %.1f us" % (C, HW, F_, e0.elapsed_time(e1) * 1e3))
    lib.dsb_test_set_chain_trace(L.ptr(trace))
    lib.dsb_test_mlp_fused(*args); torch.cuda.synchronize()
    lib.dsb_test_set_chain_trace(None)
    t = trace.cpu().reshape(24, 8)
    t0 = int(t[0, 0])
    print(" it | A-issue  g1-first  g2-last  act-start act-end  out-start out-end   (cycles since start)")
    for i in range(12):
        if t[i, 0] == 0: break
        print(" %2d | %s" % (i, "  ".join("%8d" % (int(v) - t0) for v in t[i, :7])))
