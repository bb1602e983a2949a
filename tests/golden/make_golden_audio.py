"""Generates the audio-transformer fixtures (tests/golden/audio_attn_*.npz, step_wide_av_attn_t500.npz) by running the
UNMODIFIED reference ``AudioAttnNet`` (models/audio_attention.py, built with the kwargs of cfgs/audio_visual.py:34-48)
and ``SalUNet`` on the seeded synthetic inputs/weights of diff_sal_b200/synth.py.
Run in the build container:  python tests/golden/make_golden_audio.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from diff_sal_b200 import synth  # noqa: E402


def main():
    ref_loader.load()
    from models.audio_attention import AudioAttnNet
    net = AudioAttnNet(depth=1, heads=2, dim=512, mlp_dim=256, patch_dim=512, num_patches=16, height=7, width=12,
                       pool="cls", dim_head=64, dropout=0.0, emb_dropout=0.0).eval()
    net.load_state_dict(synth.make_audio_attn_state_dict(), strict=True)
    x, feats, aud = synth.make_inputs(1, audio=True)
    out = {}
    with torch.no_grad():
        emb = net(aud.clone())
        out["audio_attn_wide_b1"] = emb
        # diff_model.py:70-113: the decoder is conditioned on the transformer's output
        dec = ref_loader.build_salunet()
        dec.load_state_dict(synth.make_state_dict("wide"), strict=True)
        out["step_wide_av_attn_t500"] = dec(x, torch.tensor([500]), list(feats), emb)
    # VGGish feature stack (models/vggish.py:87-103) on one clip's 9 log-mel-like patches; every 4th channel is kept to
    # bound the fixture size (the CUDA path is additionally compared with the full oracle output on the GPU box)
    from models.vggish import VGGish
    vg = VGGish(pretrained=False).eval()
    vg.load_state_dict(synth.make_vggish_state_dict(), strict=False)
    with torch.no_grad():
        feat = vg.forward_feat(synth.make_audio_input(1).view(-1, 1, 112, 192))
    out["vggish_feat_b1_c4"] = feat[:, ::4]
    for k, v in out.items():
        v = v.detach().float().numpy()
        np.savez_compressed(os.path.join(HERE, k + ".npz"), y=v)
        print(k, v.shape, float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
