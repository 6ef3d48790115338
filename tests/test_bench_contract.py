"""bench.py's reference arm runs without a GPU (it times the reference's CPU kernels): check that it
prints ONE JSON line with the keys of the bench contract.  A tiny sample keeps it to seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ)
    env.pop("RANK", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-reference", "16", "--nmesh", "64"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "ms" and d["higher_is_better"] is False and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # ranks other than 0 stay silent under torchrun
    env["RANK"] = "1"
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-reference", "16", "--nmesh", "64"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600, env=env)
    assert q.returncode == 0 and q.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm must raise, not print a number"""
    import ctypes
    sys.path.insert(0, ROOT)
    from pmesh_b200 import _lib
    n = ctypes.c_int(0)
    if _lib.load().pmb_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--nmesh", "32", "--steps", "1", "--warmup", "0",
                        "--no-e2e", "--no-cpu"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
