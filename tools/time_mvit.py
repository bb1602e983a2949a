# MViTv2-S encoder timing on the B200 path (once per clip, outside the loop), CUDA events
import sys, torch
sys.path.insert(0, '.')
from diff_sal_b200 import synth
from diff_sal_b200.mvit import MViTB200
for B in (1, 2, 4):
    m = MViTB200(max_batch=B); m.load_state_dict(synth.make_mvit_state_dict())
    x = synth.make_video_input(B).cuda()
    for _ in range(3): m(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): m(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("MViT-S B=%d: %.2f ms per batch, %.1f clips/s, %.1f TFLOP/s (255.65 GF/clip), %d launches" % (B, ms, B / ms * 1e3, B * 255.65e9 / (ms * 1e-3) / 1e12, m.engine.last_launch_count))
    m.engine.close()
