import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
for audio in (True, False):
    e = Engine(max_batch=1, audio_visual=audio)
    e.load_state_dict(synth.make_state_dict("wide"))
    x, feats, aud = synth.make_inputs(1, audio=audio)
    e.set_condition([f.cuda() for f in feats], aud.cuda() if audio else None)
    xs = x.cuda(); t = torch.tensor([10.0])
    names = ["tp0","tp1","tp2","h0","noise2", "noise1", "noise0", "x0", "r0", "x1", "r1", "x2", "r2", "x3", "r3", "p"]
    runs = []
    for it in range(3):
        out = e.denoise(xs, t).clone()
        torch.cuda.synchronize()
        taps = {n: e.debug_read(n, 9*516096 if n.startswith('x') else 5376*768).clone() for n in names}
        taps["out"] = out
        runs.append(taps)
    for n in names + ["out"]:
        d1 = (runs[0][n] - runs[1][n]).abs().max().item()
        d2 = (runs[1][n] - runs[2][n]).abs().max().item()
        print(audio, n, "run0-run1 %.3e run1-run2 %.3e" % (d1, d2), "absmax %.3e" % runs[0][n].abs().max().item())
