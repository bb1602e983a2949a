# B=8: conditioning + 1 warm evaluation + 1 measured evaluation (for ncu --launch-skip)
import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
B = 8
e = Engine(max_batch=B, audio_visual=True)
e.load_state_dict(synth.make_state_dict("wide"))
x, feats, aud = synth.make_inputs(B, audio=True)
e.set_condition([f.cuda() for f in feats], aud.cuda())
xs = x.cuda(); t = torch.full((B,), 500.0)
for _ in range(2):
    e.denoise(xs, t)
torch.cuda.synchronize()
