"""The CPU arm of bench.py (`--impl reference`: the oracle port on the host cores, the one place besides tests/ where oracle/ is
executed) prints the contract's JSON line, with the same `config` object the GPU arm prints for the same arguments."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--svs-per-gpu", "6", "--steps", "1",
                          "--warmup", "0", "--ref-step-seconds", "0.2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GCUPS" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert cfg["workload"].startswith("colo829_shaped (BASELINE.json configs[1])") and cfg["parallelism"] == "single GPU"
    assert set(cfg) == {"workload", "sv_candidates_per_gpu_per_step", "reads_per_gpu_per_step", "read_bases_per_gpu_mb", "l2", "parallelism"}
