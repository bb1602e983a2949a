import sys, torch
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.mvit import MViTB200
B = 2
m = MViTB200(max_batch=B); m.load_state_dict(synth.make_mvit_state_dict())
x = synth.make_video_input(B).cuda()
m(x); torch.cuda.synchronize()
m(x); torch.cuda.synchronize()
