"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py) on the seeded synthetic inputs/weights of
diff_sal_b200/synth.py.  Run in the build container:  python tests/golden/make_golden.py

The reference holds no golden vectors of its own, so these fixtures are what
pins the oracle (and through it the CUDA path) to the reference's behaviour.
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
    betas = tb.betas
    out = {}

    def ref_step(sd_kind, audio, t):
        sd = synth.make_state_dict(sd_kind)
        model.load_state_dict(sd, strict=True)
        x, feats, aud = synth.make_inputs(1, audio=audio)
        with torch.no_grad():
            return model(x, torch.tensor(t), [f.clone() for f in feats], aud)

    out["step_wide_av_t500"] = ref_step("wide", True, [500])
    out["step_wide_vis_t500"] = ref_step("wide", False, [500])
    out["step_refinit_av_t37"] = ref_step("ref_init", True, [37])
    out["step_wide_av_t886p9"] = ref_step("wide", True, [886.9])     # float timestep (DPM-solver)

    def trainer(sd_kind, S):
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
        return tr

    # O-ddim: reference DiffusionTrainer.sample_ddim, unmodified
    x, feats, aud = synth.make_inputs(1, audio=True)
    torch.manual_seed(0)
    out["ddim5_wide_av"] = trainer("wide", 5).sample_ddim(x, feats, aud)
    torch.manual_seed(0)
    out["ddim1_wide_av"] = trainer("wide", 1).sample_ddim(x, feats, aud)

    # O-dpm: reference NoiseScheduleVP + model_wrapper + DPM_Solver.sample(multistep)
    def ref_dpm(sd_kind, audio, algo, mtype, steps, order):
        model.load_state_dict(synth.make_state_dict(sd_kind), strict=True)
        x, feats, aud = synth.make_inputs(1, audio=audio)
        nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
        closure = lambda x_, t_, vis, **kw: model(x_, t_, list(vis), aud)
        mf = ns.model_wrapper(closure, nsv, model_type=mtype, model_kwargs={}, guidance_type="uncond")
        solver = ns.DPM_Solver(mf, nsv, algorithm_type=algo)
        return solver.sample(x, feats, steps=steps, order=order, skip_type="logSNR", method="multistep",
                             lower_order_final=False, denoise_to_zero=True)

    out["dpm_wide_av_xstart_o2_s4"] = ref_dpm("wide", True, "dpmsolver", "x_start", 4, 2)
    out["dpmpp_wide_av_xstart_o2_s4"] = ref_dpm("wide", True, "dpmsolver++", "x_start", 4, 2)
    out["dpm_wide_av_noise_o2_s4"] = ref_dpm("wide", True, "dpmsolver", "noise", 4, 2)
    # BASELINE config 1: visual-only, batch 1, reference init, DPM-solver multistep-2 (10 NFE)
    out["cfg1_dpm_refinit_vis_xstart_o2_s9"] = ref_dpm("ref_init", False, "dpmsolver", "x_start", 9, 2)

    for k, v in out.items():
        v = v.detach().float().numpy()
        np.savez_compressed(os.path.join(HERE, k + ".npz"), y=v)
        print(k, v.shape, float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
