"""bench.py contract (CPU part): the reference arm prints exactly one JSON line with the keys the driver reads, for both
workloads.  (The CUDA arm is exercised on the GPU box by the driver itself.)"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


@pytest.mark.parametrize("workload", ["muzero", "efficientzero"])
def test_reference_arm_prints_one_json_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                          "--warmup", "0", "--cpu-sample-roots", "8", "--sims", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "simulations/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and ("EfficientZero" in d["config"]["workload"]) == (workload == "efficientzero")
