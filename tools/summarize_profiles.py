"""Builds the tracked summaries under profiles/ from the raw ncu outputs in gpurun_out/ (tools/make_profiles.sh)."""
import csv, io, os, re, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = os.environ.get("ROUND", "r2")

def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"^dsb::", "", n)
    return re.sub(r"\(.*$", "", n)

def read_log_csv(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))

UNITS = []
def ncu_raw(rep):
    global UNITS
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    UNITS = r[1]
    return r[0], r[2:]

# ---------------------------------------------------------------- (1) launch list
rows = read_log_csv(os.path.join(G, R + "_launches.csv"))
shutil.copy(os.path.join(G, R + "_launches.csv"), os.path.join(P, R + "_launches.csv"))
agg = {}
tot = 0.0
for r in rows:
    us = float(r["Metric Value"]) / 1e3
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += us; tot += us
with open(os.path.join(P, R + "_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv python tools/ncu_one_eval.py\n")
    f.write("# B=8 audio-visual: weight prep + conditioning + 2 denoiser evaluations; per-launch times are cold-cache and\n")
    f.write("# serialised -> compare SHARES with bench.py's CUDA-event numbers, not absolutes\n")
    f.write("launches %d total us %.1f\n" % (len(rows), tot))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-60s n=%4d %11.1f us %6.1f%%\n" % (k[:60], n, us, 100 * us / tot))
gemm_us = sum(us for k, (n, us) in agg.items() if "gemm_tc" in k or "mlp_fused" in k or "splitk" in k)
prep_us = sum(us for k, (n, us) in agg.items() if any(t in k for t in ("pack_", "bn_fold", "stem_compose", "transpose", "nct_to_frames", "audio_tokens", "fold_weights", "fold_bias", "q_dw_prep", "dw_affine_prep", "audio_cmajor")))
print("launch list: %d launches, GEMM-class share of the two evaluations %.1f%%" % (len(rows), 100 * gemm_us / (tot - prep_us)))

# ---------------------------------------------------------------- (2) GEMM traffic (second evaluation)
rows = read_log_csv(os.path.join(G, R + "_gemm_traffic.csv"))
shutil.copy(os.path.join(G, R + "_gemm_traffic.csv"), os.path.join(P, R + "_gemm_traffic.csv"))
by_id = {}
for r in rows:
    d = by_id.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"])})
    v = float(r["Metric Value"])
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
    d[r["Metric Name"]] = v
ids = sorted(by_id)
n_tc = sum(1 for i in ids if "gemm_tc" in by_id[i]["name"] or "mlp_fused" in by_id[i]["name"])
per_eval_tc = (n_tc - 4) // 2
# second evaluation = trailing launches containing per_eval_tc tensor-core launches
sel, cnt = [], 0
for i in reversed(ids):
    sel.append(i)
    if "gemm_tc" in by_id[i]["name"] or "mlp_fused" in by_id[i]["name"]:
        cnt += 1
        if cnt == per_eval_tc: break
dram = sum(by_id[i].get("dram__bytes_read.sum", 0) + by_id[i].get("dram__bytes_write.sum", 0) for i in sel)
dur = sum(by_id[i]["gpu__time_duration.sum"] for i in sel)
with open(os.path.join(P, R + "_gemm_traffic_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
            "-k regex:'gemm_tc|mlp_fused|splitk_reduce' python tools/ncu_one_eval.py (B=8)\n# second (warm) evaluation only\n")
    f.write("launches %d\ndram_bytes_total %d\nduration_us_total %.1f\navg_dram_bytes_per_launch %d\n" % (per_eval_tc, dram, dur, dram / per_eval_tc))
print("traffic: %d tensor-core launches / evaluation, %.1f MB DRAM per launch" % (per_eval_tc, dram / per_eval_tc / 1e6))

# ---------------------------------------------------------------- (3) per-GEMM speed-of-light table
h, rr = ncu_raw(os.path.join(G, R + "_gemm_sol.ncu-rep"))
def col(name):
    return h.index(name) if name in h else None
want = [("gpu__time_duration.sum", "us"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("dram__bytes.sum.per_second", "GB/s"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ"), ("launch__grid_size", "grid"), ("launch__cluster_size", "clu"),
        ("sm__inst_executed_pipe_tensor.sum", "tc_inst")]
with open(os.path.join(P, R + "_gemm_sol.txt"), "w") as f:
    f.write("# ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy "
            "--clock-control none -k regex:'gemm_tc|mlp_fused' --launch-skip 47 --launch-count 43 python tools/ncu_one_eval.py\n")
    f.write("# every tensor-core launch of the second (warm) B=8 evaluation, in program order\n")
    f.write("%3s %-22s %9s %6s %6s %9s %9s %5s %3s\n" % ("#", "kernel", "us", "sm%", "dram%", "dram GB/s", "warps%", "grid", "clu"))
    ki = h.index("Kernel Name")
    for n, row in enumerate(rr):
        vals = []
        for name, _ in want[:7]:
            c = col(name)
            vals.append(row[c] if c is not None else "")
        f.write("%3d %-22s %9s %6s %6s %9s %9s %5s %3s\n" % (n, short(row[ki])[:22], vals[0][:9], vals[1][:6], vals[2][:6], vals[3][:9], vals[4][:9], vals[5], vals[6]))
print("sol table:", len(rr), "launches")

# ---------------------------------------------------------------- (4) full captures
keys = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for rep, outn, what in [(R + "_gemm_mtproj", R + "_gemm_mtproj_ncu.txt", "mt_proj 3x3 768->96 + BN + ReLU + fused 96->1 head, 112x192, B=8 (--launch-skip 81)"),
                        (R + "_gemm_upconv1", R + "_gemm_upconv1_ncu.txt", "upembed.conv1 of stage 1: dilated 3x3 768->384 + BN + ReLU on 72 frames of 14x24 (--launch-skip 62)"),
                        (R + "_mlp_fused", R + "_mlp_fused_ncu.txt", "mlp_fused_kernel, MLP chain of the last stage: fc1 -> GELU -> fc2 -> +residual, C = 96, 215 040 live tokens (-k regex:mlp_fused --launch-skip 7)")]:
    if not os.path.exists(os.path.join(G, rep + ".ncu-rep")):
        continue
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hh, units, row = r[0], r[1], r[2]
    with open(os.path.join(P, outn), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:gemm_tc ... python tools/ncu_one_eval.py (B=8)\n# %s\n" % what)
        for k in keys:
            if k in hh:
                i = hh.index(k)
                f.write("%-75s %s %s\n" % (k, row[i], units[i]))
    print(outn, "written")

# ---------------------------------------------------------------- (5) memory-bound kernels (second evaluation = last half)
h, rr = ncu_raw(os.path.join(G, R + "_membound.ncu-rep"))
ki = h.index("Kernel Name")
half = rr[len(rr) // 2:]
def g(row, name, default=""):
    return row[h.index(name)] if name in h else default
with open(os.path.join(P, R + "_membound_ncu.txt"), "w") as f:
    f.write("# ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats\n")
    f.write("# --clock-control none, memory-bound kernels of the second (warm) B=8 evaluation, program order; GB/s = DRAM bytes / duration\n")
    units = None
    f.write("%-34s %8s %10s %7s %7s %7s %7s %5s\n" % ("kernel", "us", "dram", "dram%", "sm%", "issue%", "warps%", "regs"))
    ui = h.index("dram__bytes.sum.per_second")
    for row in half:
        us = float(g(row, "gpu__time_duration.sum") or 0)
        f.write("%-34s %8.1f %10s %7s %7s %7s %7s %5s\n" % (short(row[ki])[:34], us, row[ui][:7] + " " + UNITS[ui][:6],
                g(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[:6], g(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed")[:6],
                g(row, "smsp__issue_active.avg.pct_of_peak_sustained_active")[:6], g(row, "sm__warps_active.avg.pct_of_peak_sustained_active")[:6],
                g(row, "launch__registers_per_thread")))
print("membound table:", len(half), "launches")
