# experiment: one B=8 sampler pass versus two concurrent B=4 passes (two engines, two streams); DDIM-like 10 evaluations
import sys, torch
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine

def mk(B):
    e = Engine(B, True); e.load_state_dict(synth.make_state_dict("wide"))
    x, feats, aud = synth.make_inputs(B, audio=True)
    e.set_condition([f.cuda() for f in feats], aud.cuda())
    return e, x.cuda()

ops = []
for i in range(10):
    ops.append(("eval", 900.0 - 90 * i))
    ops.append(("axpy", 0, [(0, 0.9), (1, 0.1)], 0.0, -1))

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for total in (8, 16):
    e8, x8 = mk(total)
    t8 = timed(lambda: e8.sample(ops, x8))
    del e8
    h = total // 2
    (ea, xa), (eb, xb) = mk(h), mk(h)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        cur = torch.cuda.current_stream()
        sa.wait_stream(cur); sb.wait_stream(cur)
        with torch.cuda.stream(sa): ea.sample(ops, xa)
        with torch.cuda.stream(sb): eb.sample(ops, xb)
        cur.wait_stream(sa); cur.wait_stream(sb)
    t44 = timed(both)
    tser = timed(lambda: (ea.sample(ops, xa), eb.sample(ops, xb)))
    print("total %d clips: one batch %.3f ms | two halves concurrent %.3f ms | two halves serial %.3f ms" % (total, t8, t44, tser))
    del ea, eb
