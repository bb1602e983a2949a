# one evaluation at batch B on its real streams: per-launch start/end (CUDA events), critical path and idle time
import sys, torch
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
e = Engine(B, True); e.load_state_dict(synth.make_state_dict("wide"))
x, feats, aud = synth.make_inputs(B, audio=True)
e.set_condition([f.cuda() for f in feats], aud.cuda())
t = torch.full((B,), 500.0)
xg = x.cuda()
for _ in range(3):
    tl = e.profile_timeline(xg, t)
end = max(b for _, _, b, _ in tl)
print("# B=%d one evaluation: %d launches, makespan %.3f ms" % (B, len(tl), end))
busy = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
for name, a, b, s in tl:
    busy[s] += b - a
print("# busy per stream (ms):", {k: round(v, 3) for k, v in busy.items()})
prev_end = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
print("%-22s %2s %9s %9s %8s %8s" % ("launch", "st", "start_us", "end_us", "dur_us", "gap_us"))
for name, a, b, s in tl:
    print("%-22s %2d %9.1f %9.1f %8.1f %8.1f" % (name[:22], s, 1e3 * a, 1e3 * b, 1e3 * (b - a), 1e3 * (a - prev_end[s])))
    prev_end[s] = b
