"""Builds diff_sal_b200/libdiffsal_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdiffsal_b200.so")
TEST_LIB = os.path.join(HERE, "libdiffsal_b200_test.so")       # per-kernel test entries: NOT part of the product library
TEST_SRCS = ("test_api.cu", "probe.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _flags():
    return [f for f in FLAGS if not f.startswith("--use_fast_math")]


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(TEST_LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + _flags() + ["-I", os.path.join(os.path.dirname(HERE), "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    test_objs = [o for o in objs if os.path.basename(o)[:-2] + ".cu" in TEST_SRCS]
    prod_objs = [o for o in objs if o not in test_objs]
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + prod_objs)
    # the test library resolves the internal launchers (dsb::conv_lower ...) from the product library next to it
    subprocess.check_call([NVCC, "-shared", "-o", TEST_LIB] + test_objs +
                          ["-L" + HERE, "-l:libdiffsal_b200.so", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
