"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints one JSON line that carries the
contract keys; the B200 arm refuses to run without a GPU (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_contract():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"] == "images/s masked-ViT encode+rate" and d["value"] > 0 and d["n_gpus"] == 1
    for key in ("steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly(monkeypatch):
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


def test_b200_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--no-cpu-baseline")
    assert p.returncode != 0 and "CUDA" in (p.stderr + p.stdout)
