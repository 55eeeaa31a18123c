"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) prints exactly ONE
JSON line on stdout with the keys the driver's contract names; the GPU arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    # oracle/_ref (the staged, unmodified reference) is present wherever build() ran with /root/reference mounted
    staged = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "reference", "models", "vince_model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged or os.path.isdir("/root/reference") else "port")
    assert d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["phases"]
    assert "configs[2]" in d["config"]["workload"] and "forward-only" in d["config"]["step"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr
    assert out.stdout.strip() == ""
