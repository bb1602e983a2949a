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
computes).  ``--impl reference`` times the reference's own fp32 CPU path (the
oracle port of it) on the box's host cores.
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
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM-class launch, from the committed ncu capture
    (profiles/r1_gemm_traffic_summary.txt); None if that file is missing."""
    try:
        for line in open(os.path.join(ROOT, "profiles", "r1_gemm_traffic_summary.txt")):
            if line.startswith("avg_dram_bytes_per_launch"):
                return float(line.split()[1])
    except Exception:
        pass
    return None


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


def dpm_ops():
    from diff_sal_b200 import sampler as S
    betas = S.to_torch(S.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
    ns = S.NoiseScheduleVP("discrete", betas=betas)
    ops, times = S.build_dpm_program(ns, NFE - 1, 2, "dpmsolver", "x_start", "logSNR", False, True)
    assert len(times) == NFE
    return ops


# ================================================================================================ reference arm
def run_reference(args):
    """The reference's own fp32 PyTorch sampler on the host cores (oracle port of it: the Python reference tree
    cannot travel to the GPU box).  One step = one clip through the same 10-NFE DPM-solver loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from diff_sal_b200 import synth
    from oracle import salunet, samplers as O
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sd = synth.make_state_dict("wide")
    x, feats, aud = synth.make_inputs(1, audio=True)
    net = lambda x_, t_: salunet.forward(sd, x_, t_, feats, aud)

    def step():
        return O.sample_dpm(net, x, steps=NFE - 1, order=2, algorithm_type="dpmsolver", model_type="x_start")

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = "1 clip per step through the same 10-NFE DPM-solver loop, torch fp32 CPU, %d threads" % cores
    print(json.dumps({
        "impl": "reference", "metric": "saliency clips/sec", "value": val, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ================================================================================================ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from diff_sal_b200 import synth
    from diff_sal_b200.engine import Engine
    from diff_sal_b200.parallel import gather_maps

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

    def step(k):
        x, feats, aud = sets[k & 1]
        eng.set_condition(feats, aud)
        y = eng.sample(ops, x.clone(), use_graph=True)
        if world > 1:
            y = gather_maps(y, n_clips_total)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()
    launches_per_step = eng.last_launch_count() + eng.condition_launch_count()

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
    out_host = torch.empty((B, 1, 224, 384), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * 4 for t in [host[0][0]] + host[0][1] + [host[0][2]])
    d2h = out_host.numel() * 4

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
            y = eng.sample(ops, xd, use_graph=True)
            out_hosts[s_].copy_(y, non_blocking=True)
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
                "traffic": ncu_traffic_per_launch(), "traffic_unit": "bytes per launch (ncu dram read+write, B=8)",
                "peak_source": peaks["source"],
                "launches_per_eval": gemm_n, "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1),
                "algorithmic_gflop_per_launch": gemm_fl / max(gemm_n, 1) / 1e9,
                "kernel_share_of_eval": gemm_ms / tot_ms,
                "whole_step": {"achieved": step_tf, "frac": step_tf / peaks["tflops"],
                               "note": "clips x NFE x 152.73 GFLOP / device time, per GPU"},
                "top_launches_ms": sorted(((k, round(v[0] / reps, 4)) for k, v in acc.items()), key=lambda kv: -kv[1])[:12]}

        # ------------------------------------------------------------ CPU baseline (oracle port), bounded sample
        from oracle import salunet, samplers as O
        torch.set_num_threads(os.cpu_count() or 1)
        cores = torch.get_num_threads()
        sd = synth.make_state_dict("wide")
        xc, fc, ac = synth.make_inputs(1, audio=True)
        net = lambda x_, t_: salunet.forward(sd, x_, t_, fc, ac)
        net(xc, torch.tensor([500.0]))                              # warm
        t0 = time.perf_counter()
        n_cpu = 2
        for _ in range(n_cpu):
            O.sample_dpm(net, xc, steps=NFE - 1, order=2, algorithm_type="dpmsolver", model_type="x_start")
        cpu_s = time.perf_counter() - t0
        cpu_base = {"value": n_cpu / cpu_s, "unit": "clips/s", "cores": cores, "kind": "port",
                    "sample": "%d clips (batch 1) through the same 10-NFE DPM-solver loop, torch fp32 on %d threads" % (n_cpu, cores)}

    if rank == 0:
        value = n_clips_total * args.steps / (ms * 1e-3)
        line = {
            "metric": "saliency clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": B, "nfe": NFE, "parallelism": "clip-sharded x%d" % world,
                       "l2": "two input sets alternate between steps and the per-evaluation working set (~2 GB) exceeds "
                             "the 126 MB L2; no explicit flush", "cuda_graph": True},
            "e2e": {"value": n_clips_total * args.steps / e2e_s, "unit": "clips/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "note": "wall clock over the same K steps through Engine.set_condition / Engine.sample; inputs come "
                            "from pinned host memory every step (copy stream, double-buffered: the upload of step k+1 "
                            "and the download of step k overlap the compute of step k)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
