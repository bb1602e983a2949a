"""TEST / BASELINE INFRASTRUCTURE ONLY -- builds oracle/_ref/: the UNMODIFIED reference compiled to sourceless bytecode.

The reference is pure Python, so its "build" is byte-compilation: every /root/reference/**/*.py is compiled with
py_compile from where it lies into oracle/_ref/<same relative path with the suffix .refbc> (a .pyc under another suffix: snapshot
tools tend to drop *.pyc; oracle/ref_loader.py installs a small importer that executes these code objects).  No reference source is copied into the repository;
oracle/_ref/ is git-ignored (it stays out of history) but NOT gpurun-ignored, so it travels to the GPU box like our own
built .so files.  There it lets
  * ``bench.py --impl reference`` time the reference's own fp32 PyTorch sampler on the box's host cores
    (``cpu_baseline.kind = "reference"``), and
  * the container-boundary GPU tests run the reference's own ``VideoSaliencyModel`` / ``DiffusionTrainer.sample_ddim``
    around the B200 drop-in modules.
oracle/ref_loader.py prefers /root/reference when it exists and falls back to oracle/_ref/.

Run in the build container:  python oracle/build_ref.py        (also done by __graft_entry__.build())
The bytecode is tied to the interpreter's minor version (the GPU box runs the same image).
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DIFFSAL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "models", "saliency_decoder")):
        return None                                   # GPU box: nothing to build from, the prebuilt tree is used
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            src = os.path.join(root, f)
            dst = os.path.join(DST, rel, f[:-3] + ".refbc")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
                continue
            py_compile.compile(src, cfile=dst, dfile=os.path.join("<reference>", rel, f), doraise=True, quiet=1)
            n += 1
            if verbose:
                print("compiled", os.path.join(rel, f))
    with open(os.path.join(DST, "PYTHON_VERSION"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return n


if __name__ == "__main__":
    r = build(verbose="-v" in sys.argv)
    print("reference tree not present: nothing built" if r is None else "compiled %d files into %s" % (r, DST))
