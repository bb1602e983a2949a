"""Golden fixtures for the configuration bench.py measures (BASELINE.json configs[1]) and for longer loops.

Runs the UNMODIFIED reference (NoiseScheduleVP + model_wrapper + DPM_Solver.sample(method="multistep"),
/root/reference/models/dpm_solver/sampler.py:1048-1247, around the reference SalUNet) on the seeded synthetic
inputs/weights of diff_sal_b200/synth.py.  Run in the build container:  python tests/golden/make_golden_cfg2.py

  cfg2_dpm_wide_av_b8_s9      audio-visual, batch 8, 'wide' weights, dpmsolver, x_start, order 2, steps 9 (10 NFE)
  cfg2_dpm_wide_av_b8_s9_noise  same loop with model_type="noise" (what the trainer literally passes,
                              diffusion_trainer.py:601) as ONE batch of 8 through the reference
  cfg2_dpm_refinit_av_b8_s9   as the first with the reference's own init (N(0, 0.01): nearly flat maps, min-max amplifies 26x)
  dpm_wide_av_b2_s24          25 NFE, batch 2, 'wide'
  ddim_wide_av_b2_s25 / _s10  the trainer's default sampler (sample_ddim) at 25 and 10 steps, batch 2 (BASELINE config 5)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader, samplers  # noqa: E402
from diff_sal_b200 import synth  # noqa: E402


def main():
    ns = ref_loader.load()
    model = ref_loader.build_salunet()
    import diffusion_trainer as dt
    tb = samplers.DdimTables()
    out = {}

    def ref_dpm_batch(x, feats, aud, steps, mtype):
        nsv = ns.NoiseScheduleVP(schedule="discrete", betas=tb.betas)
        closure = lambda x_, t_, vis, **kw: model(x_, t_, list(vis), aud)
        mf = ns.model_wrapper(closure, nsv, model_type=mtype, model_kwargs={}, guidance_type="uncond")
        solver = ns.DPM_Solver(mf, nsv, algorithm_type="dpmsolver")
        with torch.no_grad():
            return solver.sample(x, feats, steps=steps, order=2, skip_type="logSNR", method="multistep",
                                 lower_order_final=False, denoise_to_zero=True)

    def ref_dpm(sd_kind, B, steps, mtype="x_start"):
        """model_type="x_start": the reference's conversion `(x - alpha_t * output) / sigma_t` (sampler.py:290-292)
        multiplies a [B] vector with a [B,1,H,W] map without expanding it, so the unmodified reference only runs at
        batch 1 there; clips are independent (eval-mode norms), so the batch is the stack of B single-clip runs.
        model_type="noise" runs as one true batch."""
        model.load_state_dict(synth.make_state_dict(sd_kind), strict=True)
        x, feats, aud = synth.make_inputs(B, audio=True)
        if mtype == "noise":
            return ref_dpm_batch(x, feats, aud, steps, mtype)
        outs = []
        for b in range(B):
            outs.append(ref_dpm_batch(x[b:b + 1], [f[b:b + 1] for f in feats], aud[b:b + 1], steps, mtype))
            print("  clip", b, flush=True)
        return torch.cat(outs, 0)

    def ref_ddim(sd_kind, B, S):
        tr = dt.DiffusionTrainer.__new__(dt.DiffusionTrainer)
        tr.device = torch.device("cpu")
        tr.num_timesteps = 1000
        tr.training_target = "x0"
        tr.alphas_hat = tb.alphas_hat
        tr.sqrt_alphas_hat = tb.sqrt_alphas_hat
        tr.sqrt_recip_alphas_hat = tb.sqrt_recip_alphas_hat
        tr.sqrt_recipm1_alphas_hat = tb.sqrt_recipm1_alphas_hat
        tr.config = types.SimpleNamespace(sampling=types.SimpleNamespace(timesteps=S, eta=0.0))
        model.load_state_dict(synth.make_state_dict(sd_kind), strict=True)
        tr.model = types.SimpleNamespace(module=types.SimpleNamespace(decoder_net=model))
        x, feats, aud = synth.make_inputs(B, audio=True)
        torch.manual_seed(0)
        with torch.no_grad():
            return tr.sample_ddim(x, feats, aud)

    which = sys.argv[1:] or ["cfg2w", "cfg2r", "cfg2n", "dpm25", "ddim25", "ddim10"]
    if "cfg2w" in which:
        out["cfg2_dpm_wide_av_b8_s9"] = ref_dpm("wide", 8, 9)
    if "cfg2r" in which:
        out["cfg2_dpm_refinit_av_b8_s9"] = ref_dpm("ref_init", 8, 9)
    if "cfg2n" in which:
        out["cfg2_dpm_wide_av_b8_s9_noise"] = ref_dpm("wide", 8, 9, "noise")
    if "dpm25" in which:
        out["dpm_wide_av_b2_s24"] = ref_dpm("wide", 2, 24)
    if "ddim25" in which:
        out["ddim_wide_av_b2_s25"] = ref_ddim("wide", 2, 25)
    if "ddim10" in which:
        out["ddim_wide_av_b2_s10"] = ref_ddim("wide", 2, 10)
    for k, v in out.items():
        v = v.detach().float().numpy()
        np.savez_compressed(os.path.join(HERE, k + ".npz"), y=v)
        print(k, v.shape, float(v.min()), float(v.max()), flush=True)


if __name__ == "__main__":
    main()
