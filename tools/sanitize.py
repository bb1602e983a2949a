# small end-to-end run for compute-sanitizer: conditioning, one evaluation, a 2-step DDPM loop with thresholding ops,
# the audio transformer and split-K / corrector kernels (B=1)
import torch, sys
sys.path.insert(0, '.')
from diff_sal_b200 import synth, sampler as S
from diff_sal_b200.engine import Engine
from diff_sal_b200.audio_attention import AudioAttnNetB200
from oracle import samplers as O
x, feats, aud = synth.make_inputs(1, audio=True)
net = AudioAttnNetB200(depth=1, heads=2, dim=512, mlp_dim=256, patch_dim=512, height=7, width=12, max_batch=1)
net.load_state_dict(synth.make_audio_attn_state_dict())
emb = net(aud.cuda())
e = Engine(1, True); e.load_state_dict(synth.make_state_dict("wide"))
e.set_condition([f.cuda() for f in feats], emb)
y = e.denoise(x.cuda(), torch.tensor([500.0]))
ns = S.NoiseScheduleVP("discrete", betas=O.betas_fp32())
ops, _ = S.build_dpm_program(ns, 2, 2, "dpmsolver++", "x_start", "logSNR", False, True, thresholding=(0.995, 0.2))
ops.append(("clamp", 0, 0.0, 1.0))
z = e.sample(ops, x.cuda().clone(), use_graph=False)
torch.cuda.synchronize()
print("ok", float(y.mean()), float(z.mean()))
