import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
e = Engine(max_batch=B, audio_visual=True)
e.load_state_dict(synth.make_state_dict("wide"))
x, feats, aud = synth.make_inputs(B, audio=True)
e.set_condition([f.cuda() for f in feats], aud.cuda())
xs = x.cuda(); t = torch.full((B,), 500.0)
e.profile_denoise(xs, t)
reps = 5
rows = None
for _ in range(reps):
    r = e.profile_denoise(xs, t)
    if rows is None: rows = [[n, 0.0, fl, by] for n, m, fl, by in r]
    for i, (n, m, fl, by) in enumerate(r): rows[i][1] += m / reps
tot = sum(r[1] for r in rows)
print("B=%d total %.3f ms per eval (%d launches)" % (B, tot, len(rows)))
agg = {}
for n, m, fl, by in rows:
    a = agg.setdefault(n, [0.0, 0, 0.0]); a[0] += m; a[1] += 1; a[2] += fl
for n, (m, c, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-22s n=%2d %8.1f us %5.1f%%  %s" % (n, c, m * 1e3, 100 * m / tot, ("%.0f TF/s" % (fl / (m * 1e-3) / 1e12)) if fl else ""))
