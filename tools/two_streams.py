# experiment: one B=8 engine vs two concurrent B=4 engines on two streams (same 8 clips, same program)
import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth, sampler as S
from diff_sal_b200.engine import Engine
from oracle import samplers as O
ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
ops, _ = S.build_dpm_program(ns, 9, 2, "dpmsolver", "x_start", "logSNR", False, True)
sd = synth.make_state_dict("wide")
x, feats, aud = synth.make_inputs(8, audio=True)
x, feats, aud = x.cuda(), [f.cuda() for f in feats], aud.cuda()

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

e8 = Engine(8, True); e8.load_state_dict(sd); e8.set_condition(feats, aud)
def one():
    e8.sample(ops, x.clone())
t8 = timeit(one)
print("1 x B=8: %.2f ms -> %.1f clips/s" % (t8, 8e3 / t8))
for parts in (2, 4):
    bs = 8 // parts
    engs, streams, xs = [], [], []
    for k in range(parts):
        e = Engine(bs, True); e.load_state_dict(sd)
        e.set_condition([f[k*bs:(k+1)*bs].contiguous() for f in feats], aud[k*bs:(k+1)*bs].contiguous())
        engs.append(e); streams.append(torch.cuda.Stream()); xs.append(x[k*bs:(k+1)*bs].contiguous())
    torch.cuda.synchronize()
    def many():
        cur = torch.cuda.current_stream()
        for e, st, xx in zip(engs, streams, xs):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                e.sample(ops, xx.clone())
        for st in streams: cur.wait_stream(st)
    t = timeit(many)
    print("%d x B=%d concurrent: %.2f ms -> %.1f clips/s" % (parts, bs, t, 8e3 / t))
    for e in engs: e.close()
