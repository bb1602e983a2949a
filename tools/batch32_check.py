# B = 32 sanity: per-clip batch invariance against B = 1 and throughput of one evaluation (BASELINE config 3)
import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
sd = synth.make_state_dict("wide")
x, feats, aud = synth.make_inputs(32, audio=True)
e32 = Engine(32, True); e32.load_state_dict(sd)
e32.set_condition([f.cuda() for f in feats], aud.cuda())
t = torch.full((32,), 500.0)
y32 = e32.denoise(x.cuda(), t)
e1 = Engine(1, True); e1.load_state_dict(sd)
for i in (0, 17, 31):
    e1.set_condition([f[i:i+1].cuda() for f in feats], aud[i:i+1].cuda())
    y1 = e1.denoise(x[i:i+1].cuda(), t[:1])
    print("clip %d bitwise equal to its B=1 run:" % i, torch.equal(y1, y32[i:i+1]))
for _ in range(3): e32.denoise(x.cuda(), t)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): e32.denoise(x.cuda(), t)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("B=32: %.2f ms per evaluation -> %.0f clip-evals/s, %.1f%% of 1390 TF/s" % (ms, 32e3 / ms, 32 * 152.73e9 / (ms * 1e-3) / 1390e12 * 100))
print("workspace GB:", torch.cuda.memory_allocated() / 1e9, "(torch) + handle")
