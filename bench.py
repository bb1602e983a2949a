#!/usr/bin/env python
"""bench.py -- saliency clips/sec of the DiffSal sampling hot path on N B200s of one node.

Contract (see DESIGN.md "Measurement"):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  N > 1 is launched by the driver under torch.distributed.run (one rank per GPU, NCCL).

A "step" = one full sampling loop (DPM-solver multistep-2, logSNR, 10 NFE incl. denoise-to-zero, model_type x_start)
over one batch of 8 synthetic audio-visual clips per GPU (BASELINE.json configs[1]; weak scaling: 8 clips per GPU),
including the per-batch conditioning (layout conversion + loop-invariant align_conv) and, for N > 1, the NCCL gather
of the predicted maps.  ``value`` is device-timed (CUDA events, max over ranks) with inputs resident in HBM;
``e2e`` is the same work through the public API from pinned HOST buffers (H2D of x_T / features / audio and D2H of
the maps inside the timed region, wall clock; uploads of the next batch run on a copy stream while the current batch
computes).  ``--impl reference`` times the reference's own fp32 CPU path on the box's host cores: the
UNMODIFIED reference imported from oracle/_ref (its bytecode build, see oracle/build_ref.py; kind "reference"), or the
oracle port when that build is absent (kind "port").

``--config`` selects the BASELINE.json configuration (default 2 = the headline; the driver runs the default):
  2  audio-visual batch 8 per GPU, full 10-NFE DPM-solver loop (weak scaling)
  3  denoiser-only microbench: one SalUNet evaluation at batch 1 / 8 / 32 on one GPU
  4  256 clips sharded over the N GPUs in micro-batches of 32, maps gathered over NVLink (strong scaling)
  5  sampling-step sweep (1 / 5 / 10 / 25 NFE) at 8 clips per GPU with CC / NSS / SIM / AUC-J drift vs the fp32 oracle
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CLIPS_PER_GPU = 8
NFE = 10
GFLOP_PER_CLIP_EVAL = 152.73          # torch flop counter on the reference graph (audio-visual)


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM-class launch, from the newest committed ncu capture of
    tools/make_profiles.sh (profiles/r*_gemm_traffic_summary.txt; hardware counters cannot be read from inside the timed
    process); None if no capture is committed."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic_summary.txt")), reverse=True):
        try:
            for line in open(path):
                if line.startswith("avg_dram_bytes_per_launch"):
                    return float(line.split()[1]), os.path.basename(path)
        except Exception:
            pass
    return None, None


WORKLOAD = ("audio_visual (cfgs/audio_visual.py shapes) batch 8 clips/GPU, DPM-solver multistep order 2, logSNR steps, "
            "10 NFE (steps=9 + denoise_to_zero), x0-parameterised SalUNet, random-init 'wide' weights")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"tflops": float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), "gbs": float(d["hbm_gbs"]),
                    "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
        except Exception:
            pass
    return {"tflops": 1400.0, "gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def check_parity(y, golden="cfg2_dpm_wide_av_b8_s9", tol=1e-2):
    """Per-clip max-abs error of min-max-normalised maps against the committed reference fixture (north_star tolerance)."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", golden + ".npz")
    if not os.path.exists(path):
        return {"checked": False, "why": "fixture %s missing" % golden}
    ref = torch.from_numpy(np.load(path)["y"])[: y.shape[0]]

    def mm(v):
        f = v.reshape(v.shape[0], -1)
        lo, hi = f.min(dim=1, keepdim=True).values, f.max(dim=1, keepdim=True).values
        return (f - lo) / (hi - lo)
    err = (mm(y) - mm(ref)).abs().max(dim=1).values
    return {"checked": True, "golden": "tests/golden/%s.npz (unmodified reference, fp32 CPU)" % golden, "clips": int(y.shape[0]),
            "max_abs_minmax_per_clip": [round(float(e), 5) for e in err], "max_abs_minmax": float(err.max()), "tol": tol,
            "ok": bool(err.max().item() <= tol)}


def dpm_ops():
    from diff_sal_b200 import sampler as S
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    ops, times = S.build_dpm_program(ns, NFE - 1, 2, "dpmsolver", "x_start", "logSNR", False, True)
    assert len(times) == NFE
    return ops


# ================================================================================================ reference arm
def _quiet(fn):
    """The reference prints to stdout (e.g. get_beta_schedule); bench.py's stdout is ONE JSON line."""
    import contextlib

    def run(*a, **k):
        with contextlib.redirect_stdout(sys.stderr):
            return fn(*a, **k)
    return run


def reference_sampler():
    """(step_fn, kind, note): one clip through the 10-NFE DPM-solver loop on the CPU.  kind "reference" = the unmodified
    reference (SalUNet + NoiseScheduleVP / model_wrapper / DPM_Solver.sample, models/dpm_solver/sampler.py:1048-1247)
    imported from oracle/_ref or /root/reference; kind "port" = the oracle restatement when neither is present."""
    from diff_sal_b200 import synth
    sd = synth.make_state_dict("wide")
    x, feats, aud = synth.make_inputs(1, audio=True)
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            raise RuntimeError("no reference build")
        ns = ref_loader.load()
        model = _quiet(ref_loader.build_salunet)()
        model.load_state_dict(sd, strict=True)
        betas = _quiet(lambda: ns.to_torch(ns.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)))()
        nsv = ns.NoiseScheduleVP(schedule="discrete", betas=betas)
        closure = lambda x_, t_, vis, **kw: model(x_, t_, list(vis), aud)
        mf = ns.model_wrapper(closure, nsv, model_type="x_start", model_kwargs={}, guidance_type="uncond")
        solver = ns.DPM_Solver(mf, nsv, algorithm_type="dpmsolver")

        def step():
            with torch.no_grad():
                return solver.sample(x, feats, steps=NFE - 1, order=2, skip_type="logSNR", method="multistep",
                                     lower_order_final=False, denoise_to_zero=True)
        src = "sources" if ref_loader.is_source_tree() else "bytecode build oracle/_ref"
        return _quiet(step), "reference", "unmodified reference (%s): SalUNet + DPM_Solver.sample" % src
    except Exception as e:                                            # noqa: BLE001 -- fall back to the port, say why
        from oracle import salunet, samplers as O
        net = lambda x_, t_: salunet.forward(sd, x_, t_, feats, aud)
        why = "%s: %s" % (type(e).__name__, e)

        def step():
            return O.sample_dpm(net, x, steps=NFE - 1, order=2, algorithm_type="dpmsolver", model_type="x_start")
        return step, "port", "oracle port (reference build unavailable: %s)" % why[:120]


def run_reference(args):
    """The reference's own fp32 PyTorch sampler on the host cores.  One step = one clip through the same 10-NFE
    DPM-solver loop (batch 1 is the reference's best case per clip, SURVEY 6, and the only batch its x_start
    conversion supports, sampler.py:290-292)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    step, kind, note = reference_sampler()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = "1 clip per step through the same 10-NFE DPM-solver loop, torch fp32 CPU, %d threads; %s" % (cores, note)
    print(json.dumps({
        "impl": "reference", "metric": "saliency clips/sec", "value": val, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ================================================================================================ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from diff_sal_b200 import synth
    from diff_sal_b200.engine import Engine
    from diff_sal_b200.parallel import MapGatherer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = CLIPS_PER_GPU
    eng = Engine(max_batch=B, audio_visual=True)
    eng.load_state_dict(synth.make_state_dict("wide"))
    ops = dpm_ops()
    # two input sets (alternated) so that a step never finds its inputs in L2 from the previous step
    sets = []
    for k in range(2):
        x, feats, aud = synth.make_inputs(B, audio=True, seed=1234 + 1000 * k + 100 * rank)
        sets.append((x.to(dev), [f.to(dev) for f in feats[:3]], aud.to(dev)))
    n_clips_total = B * world
    gatherer = MapGatherer(n_clips_total, (1, 224, 384), torch.float32, dev)     # preallocated gather buffers
    xbuf = torch.empty_like(sets[0][0])

    # The gather stays in line: MapGatherer.start / finish can put it on a side stream beside the next step's loop, which
    # measured +1.7 % at N = 2 but -1.1 % at N = 8 (the NCCL kernel's CTAs take SMs from the persistent one-CTA-per-SM
    # tensor-core kernels and break up their CTA pairs) -- DESIGN.md section 7.
    def step(k):
        x, feats, aud = sets[k & 1]
        eng.set_condition(feats, aud)
        xbuf.copy_(x)
        y = eng.sample(ops, xbuf, use_graph=True)
        if world > 1:
            y = gatherer(y)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()
    launches_per_step = eng.last_launch_count() + eng.condition_launch_count()
    # parity of what is timed: input set 0 of rank 0 is the seeded batch of tests/golden/cfg2_dpm_wide_av_b8_s9.npz (maps of
    # the unmodified reference, tests/golden/make_golden_cfg2.py); every clip is compared after min-max normalisation
    parity = None
    if rank == 0:
        parity = check_parity(step(0)[:B].float().cpu())
        barrier()
    elif world > 1:
        step(0)
        barrier()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()

    # ---------------------------------------------------------------- e2e: public API from pinned host buffers
    host = []
    for k in range(2):
        x, feats, aud = sets[k]
        host.append((x.cpu().pin_memory(), [f.cpu().pin_memory() for f in feats], aud.cpu().pin_memory()))
    # the result a user of the reference keeps is normalize_data's uint8 map (diffusion_trainer.py:898-935 -> PNG): the
    # loop's last op produces it on the device (DSB_OP_POSTPROCESS), so 86 KB per clip cross PCIe instead of 344 KB
    out_host = torch.empty((B, 1, 224, 384), dtype=torch.uint8).pin_memory()
    h2d = sum(t.numel() * 4 for t in [host[0][0]] + host[0][1] + [host[0][2]])
    d2h = out_host.numel()
    ops_post = ops + [("post", 0)]
    u8_dev = [torch.empty((B, 1, 224, 384), dtype=torch.uint8, device=dev) for _ in range(2)]

    # Double-buffered: while step k computes, a copy stream uploads the inputs of step k + 1 from pinned host memory and
    # the result of step k goes back to pinned host memory asynchronously; every step's H2D and D2H happen inside the
    # timed region (they overlap compute, they are not skipped).
    copy_s = torch.cuda.Stream(device=dev)
    dev_in = [(torch.empty_like(sets[k][0]), [torch.empty_like(f) for f in sets[k][1]], torch.empty_like(sets[k][2]))
              for k in range(2)]
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(k):
        s_ = k & 1
        xh, fh, ah = host[s_]
        xd, fd, ad = dev_in[s_]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done[s_])                  # the step that last used this slot has finished with it
            xd.copy_(xh, non_blocking=True)
            for d_, h_ in zip(fd, fh):
                d_.copy_(h_, non_blocking=True)
            ad.copy_(ah, non_blocking=True)
            ready[s_].record(copy_s)

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        prefetch(0)
        for k in range(n):
            s_ = k & 1
            if k + 1 < n:
                prefetch(k + 1)
            cur.wait_event(ready[s_])
            xd, fd, ad = dev_in[s_]
            eng.set_condition(fd, ad)
            eng.sample(ops_post, xd, use_graph=True, out_u8=u8_dev[s_])
            out_hosts[s_].copy_(u8_dev[s_], non_blocking=True)
            done[s_].record(cur)
        torch.cuda.synchronize()

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = e2e_s.item()
    clk = clocks.stop() if rank == 0 else None

    # ---------------------------------------------------------------- roofline of the dominant kernel (rank 0)
    peaks = measured_peaks()
    roof = None
    cpu_base = None
    if rank == 0:
        x, feats, aud = sets[0]
        eng.set_condition(feats, aud)
        t = torch.full((B,), 500.0)
        eng.profile_denoise(x, t)                                   # warm
        acc = {}
        reps = 3
        for _ in range(reps):
            for name, m, fl, by in eng.profile_denoise(x, t):
                a = acc.setdefault(name, [0.0, 0.0, 0.0, 0])
                a[0] += m; a[1] += fl; a[2] += by; a[3] += 1
        gemm_ms = sum(v[0] for k, v in acc.items() if k.startswith("gemm:")) / reps
        gemm_fl = sum(v[1] for k, v in acc.items() if k.startswith("gemm:")) / reps
        gemm_n = sum(v[3] for k, v in acc.items() if k.startswith("gemm:")) / reps
        tot_ms = sum(v[0] for v in acc.values()) / reps
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12
        step_tf = n_clips_total * NFE * GFLOP_PER_CLIP_EVAL * 1e9 * args.steps / (ms * 1e-3) / 1e12 / world
        roof = {"bound": "tensor",
                "kernel": "gemm_tc_kernel<0|1> + mlp_fused_kernel (tcgen05 implicit GEMM / fused GEMM chains: every conv, linear and attention launch)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": ncu_traffic_per_launch()[0],
                "traffic_unit": "bytes per launch (ncu dram read+write, B=8; %s)" % ncu_traffic_per_launch()[1],
                "peak_source": peaks["source"],
                "launches_per_eval": gemm_n, "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1),
                "algorithmic_gflop_per_launch": gemm_fl / max(gemm_n, 1) / 1e9,
                "kernel_share_of_eval": gemm_ms / tot_ms,
                "whole_step": {"achieved": step_tf, "frac": step_tf / peaks["tflops"],
                               "note": "clips x NFE x 152.73 GFLOP / device time, per GPU; 152.73 GFLOP is the reference "
                                       "graph's count (torch flop counter, SURVEY 8d), i.e. it credits work the kernels "
                                       "skip exactly: attention/MLP of the dead frames 5..8 of the last stage (~3 %)"},
                "top_launches_ms": sorted(((k, round(v[0] / reps, 4)) for k, v in acc.items()), key=lambda kv: -kv[1])[:12]}

        # ------------------------------------------------------------ CPU baseline (the reference itself), bounded sample
        torch.set_num_threads(os.cpu_count() or 1)
        cores = torch.get_num_threads()
        cpu_step, cpu_kind, cpu_note = reference_sampler()
        cpu_step()                                                  # warm
        t0 = time.perf_counter()
        n_cpu = 2
        for _ in range(n_cpu):
            cpu_step()
        cpu_s = time.perf_counter() - t0
        cpu_base = {"value": n_cpu / cpu_s, "unit": "clips/s", "cores": cores, "kind": cpu_kind,
                    "sample": "%d clips (batch 1) through the same 10-NFE DPM-solver loop, torch fp32 on %d threads; %s"
                              % (n_cpu, cores, cpu_note)}

    if rank == 0:
        value = n_clips_total * args.steps / (ms * 1e-3)
        line = {
            "metric": "saliency clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "dtype_note": "16-bit tensor-core operands, fp32 accumulate / epilogues / residual streams / norm statistics; bf16 "
                          "operands everywhere except the two GEMMs of the output-head chain (ReduceTemp, mt_proj), which take "
                          "fp16 operands at the same kind::f16 rate (DESIGN.md section 2, precision policy)",
            "config": {"workload": WORKLOAD, "clips_per_gpu": B, "nfe": NFE, "parallelism": "clip-sharded x%d" % world,
                       "l2": "two input sets alternate between steps and the per-evaluation working set (~2 GB) exceeds "
                             "the 126 MB L2; no explicit flush", "cuda_graph": True},
            "e2e": {"value": n_clips_total * args.steps / e2e_s, "unit": "clips/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "note": "wall clock over the same K steps through Engine.set_condition / Engine.sample; inputs come "
                            "from pinned host memory every step (copy stream, double-buffered: the upload of step k+1 "
                            "and the download of step k overlap the compute of step k); the result read back is the "
                            "min-max-normalised uint8 map (normalize_data), produced by the loop's last op"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clk, "parity": parity, "roofline": roof, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ================================================================================================ other BASELINE configs
def _dist_setup(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    return world, rank, local, dev, barrier, max_over_ranks


def _timed(fn, steps, warmup, barrier, max_over_ranks):
    for k in range(max(warmup, 3)):
        fn(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn(k)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps


def run_config3(args):
    """BASELINE config 3: one SalUNet evaluation (conv blocks + cross-modal attention) at batch 1 / 8 / 32 on one GPU.
    Per batch: in-graph time of one evaluation (10 back-to-back evaluations in one CUDA graph), the serialised
    CUDA-event sum over its launches, and the roofline fraction (152.73 GFLOP per clip-evaluation)."""
    from diff_sal_b200 import synth
    from diff_sal_b200.engine import Engine
    world, rank, local, dev, barrier, mx = _dist_setup(args)
    peaks = measured_peaks()
    sd = synth.make_state_dict("wide")
    rows = []
    clocks = ClockSampler(local)
    clocks.start()
    for B in (1, 8, 32):
        eng = Engine(max_batch=B, audio_visual=True)
        eng.load_state_dict(sd)
        sets = []
        for k in range(2):
            x, feats, aud = synth.make_inputs(B, audio=True, seed=1234 + 1000 * k)
            sets.append((x.to(dev), [f.to(dev) for f in feats[:3]], aud.to(dev)))
        n_eval = 10
        ops = [("eval", 500.0 - 40.0 * i) for i in range(n_eval)]
        xbuf = torch.empty_like(sets[0][0])

        def step(k):
            x, feats, aud = sets[k & 1]
            eng.set_condition(feats, aud)
            xbuf.copy_(x)
            eng.sample(ops, xbuf, use_graph=True)
        ms = _timed(step, args.steps, args.warmup, barrier, mx) / n_eval
        x, feats, aud = sets[0]
        eng.set_condition(feats, aud)
        t = torch.full((B,), 500.0)
        eng.profile_denoise(x, t)
        prof = eng.profile_denoise(x, t)
        ser = sum(p[1] for p in prof)
        gemm = sum(p[1] for p in prof if p[0].startswith("gemm:"))
        tf = B * GFLOP_PER_CLIP_EVAL * 1e9 / (ms * 1e-3) / 1e12
        rows.append({"batch": B, "ms_per_evaluation_in_graph": ms, "ms_per_evaluation_serialised_events": ser,
                     "tensor_core_launch_share": gemm / ser, "launches": len(prof), "clip_evaluations_per_s": B / (ms * 1e-3),
                     "achieved_tflops": tf, "roofline_frac": tf / peaks["tflops"]})
        eng.close()
    clk = clocks.stop()
    best = max(rows, key=lambda r: r["clip_evaluations_per_s"])
    print(json.dumps({
        "metric": "denoiser clip-evaluations/sec", "value": best["clip_evaluations_per_s"], "unit": "clip-evals/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": best["ms_per_evaluation_in_graph"] * 10,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE config 3: one SalUNet.forward (audio-visual) at batch 1 / 8 / 32, 10 evaluations per "
                               "CUDA graph, conditioning per step; value = best batch (%d)" % best["batch"],
                   "l2": "two input sets alternate; working set per evaluation exceeds L2 from batch 8 on"},
        "per_batch": rows, "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": best["achieved_tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": best["roofline_frac"], "traffic": None, "peak_source": peaks["source"]}}))


def run_config4(args):
    """BASELINE config 4: 256 audio-visual clips sharded over the N GPUs (contiguous blocks of 256 / N), each rank working
    through its block in micro-batches of 32, predicted maps gathered over NVLink (uint8 maps are what the reference
    keeps; the fp32 maps are gathered, as in config 2).  Strong scaling: total work is fixed."""
    from diff_sal_b200 import synth
    from diff_sal_b200.engine import Engine
    from diff_sal_b200.parallel import MapGatherer, micro_batches, shard_range
    world, rank, local, dev, barrier, mx = _dist_setup(args)
    N_CLIPS, MB = 256, 32
    lo, hi = shard_range(N_CLIPS, rank, world)
    eng = Engine(max_batch=MB, audio_visual=True)
    eng.load_state_dict(synth.make_state_dict("wide"))
    ops = dpm_ops()
    # synthetic clips: two distinct micro-batches per rank, alternated (generating 256 distinct clips on the host would
    # only time the generator; every micro-batch still goes through conditioning + the full loop)
    sets = []
    for k in range(2):
        x, feats, aud = synth.make_inputs(MB, audio=True, seed=1234 + 1000 * k + 100 * rank)
        sets.append((x.to(dev), [f.to(dev) for f in feats[:3]], aud.to(dev)))
    local_maps = torch.empty((hi - lo, 1, 224, 384), dtype=torch.float32, device=dev)
    gatherer = MapGatherer(N_CLIPS, (1, 224, 384), torch.float32, dev)
    xbuf = torch.empty_like(sets[0][0])
    mbs = micro_batches(lo, hi, MB)

    def step(k):
        for j, (a, b) in enumerate(mbs):
            x, feats, aud = sets[(k + j) & 1]
            n = b - a
            eng.set_condition([f[:n] for f in feats], aud[:n])
            xb = xbuf[:n]
            xb.copy_(x[:n])
            eng.sample(ops, xb, use_graph=True)
            local_maps[a - lo:b - lo].copy_(xb)
        return gatherer(local_maps)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = _timed(step, args.steps, args.warmup, barrier, mx)
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        peaks = measured_peaks()
        val = N_CLIPS / (ms * 1e-3)
        tf = N_CLIPS * NFE * GFLOP_PER_CLIP_EVAL * 1e9 / (ms * 1e-3) / 1e12 / world
        cpu_step, cpu_kind, cpu_note = reference_sampler()
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_step()
        t0 = time.perf_counter()
        cpu_step()
        cpu_s = time.perf_counter() - t0
        print(json.dumps({
            "metric": "saliency clips/sec", "value": val, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: 256 audio-visual clips, %d per GPU in micro-batches of 32, 10-NFE "
                                   "DPM-solver loop, fp32 maps all-gathered over NVLink" % (hi - lo),
                       "parallelism": "clip-sharded x%d" % world, "cuda_graph": True,
                       "l2": "two micro-batch input sets alternate; per-evaluation working set far exceeds L2"},
            "clocks": clk, "gpu_launches": int((eng.last_launch_count() + eng.condition_launch_count()) * len(mbs) * args.steps),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"],
                         "traffic": None, "peak_source": peaks["source"], "note": "whole step, per GPU"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": cpu_kind,
                             "sample": "1 clip (batch 1) through the same loop; " + cpu_note}}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE config 5: sampling-step sweep -- 1 / 5 / 10 / 25 network evaluations (DPM-solver multistep; order 1 for a
    single step as the reference requires steps >= order, sampler.py:1174) at 8 clips per GPU, with the drift of
    CC / NSS / SIM / AUC-J against the fp32 CPU oracle on rank 0's first 2 clips (bounded CPU sample)."""
    from diff_sal_b200 import sampler as S, synth
    from diff_sal_b200.engine import Engine
    from diff_sal_b200.parallel import MapGatherer
    world, rank, local, dev, barrier, mx = _dist_setup(args)
    B = CLIPS_PER_GPU
    eng = Engine(max_batch=B, audio_visual=True)
    sd = synth.make_state_dict("wide")
    eng.load_state_dict(sd)
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    x, feats, aud = synth.make_inputs(B, audio=True, seed=1234 + 100 * rank)
    xd, fd, ad = x.to(dev), [f.to(dev) for f in feats[:3]], aud.to(dev)
    gatherer = MapGatherer(B * world, (1, 224, 384), torch.float32, dev)
    xbuf = torch.empty_like(xd)
    rows = []
    for nfe in (1, 5, 10, 25):
        if nfe == 1:
            kw = dict(steps=1, order=1, denoise_to_zero=False)
        else:
            kw = dict(steps=nfe - 1, order=2, denoise_to_zero=True)
        ops, times = S.build_dpm_program(ns, kw["steps"], kw["order"], "dpmsolver", "x_start", "logSNR", False, kw["denoise_to_zero"])
        assert len(times) == nfe

        def step(k):
            eng.set_condition(fd, ad)
            xbuf.copy_(xd)
            eng.sample(ops, xbuf, use_graph=True)
            return gatherer(xbuf)
        ms = _timed(step, max(3, args.steps // 2), args.warmup, barrier, mx)
        row = {"nfe": nfe, "ms_per_step": ms, "clips_per_s": B * world / (ms * 1e-3)}
        if rank == 0:
            from oracle import metrics, salunet, samplers as O
            y = step(0)[:2].float().cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            xc, fc, ac = x[:2], [f[:2] for f in feats], aud[:2]
            ref = O.sample_dpm(lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac), xc, steps=kw["steps"], order=kw["order"],
                               algorithm_type="dpmsolver", model_type="x_start", denoise_to_zero=kw["denoise_to_zero"])
            drift, err = {}, 0.0
            for b in range(2):
                rm = O.inverse_data_transform(ref[b:b + 1])[0, 0].double().numpy()
                gm = O.inverse_data_transform(y[b:b + 1])[0, 0].double().numpy()
                gt = metrics.ground_truth_from_map(rm, b)
                a_, r_ = metrics.all_metrics(gm, b, gt=gt), metrics.all_metrics(rm, b, gt=gt)
                for k_ in a_:
                    drift[k_] = max(drift.get(k_, 0.0), abs(a_[k_] - r_[k_]) / abs(r_[k_]))
            err = float((O.minmax_map(y) - O.minmax_map(ref)).abs().max())
            row.update({"max_abs_minmax_vs_fp32_oracle": err, "relative_metric_drift": drift, "clips_checked": 2})
        else:
            step(0)
        barrier()
        rows.append(row)
    if rank == 0:
        ten = [r for r in rows if r["nfe"] == 10][0]
        print(json.dumps({
            "metric": "saliency clips/sec", "value": ten["clips_per_s"], "unit": "clips/s", "n_gpus": world, "steps": max(3, args.steps // 2),
            "warmup": max(args.warmup, 3), "ms_per_step": ten["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: sampling-step sweep 1/5/10/25 NFE, DPM-solver multistep, %d clips (8 per "
                                   "GPU); value = the 10-NFE row" % (B * world), "parallelism": "clip-sharded x%d" % world},
            "sweep": rows}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {2: 20, 3: 10, 4: 3, 5: 6}[args.config]
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 3:
        run_config3(args)
    elif args.config == 4:
        run_config4(args)
    elif args.config == 5:
        run_config5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
