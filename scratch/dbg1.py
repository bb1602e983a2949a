import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.engine import Engine
e = Engine(max_batch=2, audio_visual=True)
e.load_state_dict(synth.make_state_dict("wide"))
torch.cuda.synchronize(); print("finalized")
x, feats, aud = synth.make_inputs(2, audio=True)
e.set_condition([f.cuda() for f in feats], aud.cuda())
torch.cuda.synchronize(); print("conditioned")
out = e.denoise(x.cuda(), torch.tensor([500.0, 37.0]))
torch.cuda.synchronize(); print("denoised", out.mean().item())
