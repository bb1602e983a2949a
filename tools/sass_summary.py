"""profiles/<round>_sass_summary.txt: Blackwell-native instruction counts per kernel of the built product library."""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.environ.get("ROUND", "r2")
lib = os.path.join(ROOT, "diff_sal_b200", "libdiffsal_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cnt = collections.defaultdict(lambda: collections.Counter())
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn is None:
        continue
    if "UTCHMMA" in line:
        cnt[fn]["UTCHMMA"] += 1
        if ".2CTA" in line:
            cnt[fn]["2CTA"] += 1
    elif re.search(r"[^C]HMMA\.", line):
        cnt[fn]["HMMA"] += 1
    if "LDTM" in line:
        cnt[fn]["LDTM"] += 1
    if "UTMALDG" in line:
        cnt[fn]["UTMALDG"] += 1
    if "UBLKCP" in line:
        cnt[fn]["UBLKCP"] += 1
    if "FFMA2" in line or "FMUL2" in line or "FADD2" in line:
        cnt[fn]["F2"] += 1
demangle = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(cnt, demangle))
path = os.path.join(ROOT, "profiles", R + "_sass_summary.txt")
with open(path, "w") as f:
    f.write("# cuobjdump -sass diff_sal_b200/libdiffsal_b200.so : Blackwell-native instruction counts per kernel (tools/sass_summary.py)\n")
    f.write("# UTCHMMA = tcgen05.mma kind::f16 (bf16 / fp16 operands), .2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),\n")
    f.write("# UBLKCP = cp.async.bulk (1-D TMA), F2 = packed fp32 (FFMA2 / FMUL2 / FADD2).  gemm_tc_kernel<pair, epilogue flavour>: see gemm_tc.cu\n")
    f.write("%-96s %7s %6s %5s %8s %6s %4s\n" % ("kernel", "UTCHMMA", ".2CTA", "LDTM", "UTMALDG", "UBLKCP", "F2"))
    tot = collections.Counter()
    for k, c in sorted(cnt.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
        if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["UBLKCP"] or c["HMMA"]):
            continue
        n = re.sub(r"\(.*$", "", names[k]).replace("void ", "").replace("dsb::", "")
        f.write("%-96s %7d %6d %5d %8d %6d %4d\n" % (n[:96], c["UTCHMMA"], c["2CTA"], c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["F2"]))
        tot.update(c)
    f.write("\n# totals\nUTCHMMA total %d\nUTCHMMA.2CTA total %d\nLDTM total %d\nUTMALDG total %d\nUBLKCP total %d\n" % (tot["UTCHMMA"], tot["2CTA"], tot["LDTM"], tot["UTMALDG"], tot["UBLKCP"]))
    f.write("legacy HMMA (mma.sync, not preceded by UTC) total %d\n" % tot["HMMA"])
    nm = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    f.write("exported dsb_test_* symbols in the product library: %d (the per-kernel test entry points live in libdiffsal_b200_test.so)\n" % sum(1 for l in nm.splitlines() if "dsb_test_" in l))
print(open(path).read())
