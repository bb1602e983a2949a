"""CPU: bench.py's reference arm (the reference's own fp32 path on the host cores, here the oracle port of it) prints one
JSON line with the keys the measurement contract names; the B200 arm refuses to run without a GPU (no fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "saliency clips/sec" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_runs_the_unmodified_reference_from_its_bytecode_build():
    """oracle/build_ref.py compiles /root/reference to sourceless bytecode under oracle/_ref/ (what travels to the GPU
    box); pointed at that tree alone, the reference arm must run the real reference (kind "reference")."""
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    if build_ref.build() is None and not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "models")):
        import pytest
        pytest.skip("no reference tree and no prebuilt oracle/_ref")
    env = dict(os.environ, DIFFSAL_REFERENCE=os.path.join(ROOT, "oracle", "_ref"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert d["cpu_baseline"]["kind"] == "reference", d["cpu_baseline"]
    assert "bytecode build" in d["cpu_baseline"]["sample"]


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0            # fails loudly: no CPU fallback
