"""
The part of bench.py's contract that runs without a GPU: the reference arm (`--impl reference` times the CPU exact path -
the oracle's tuned arm - on a bounded sample and prints ONE JSON line with the same metric / unit / config keys as the GPU
arm), and the refusal of the product arm to run without the CUDA device (no CPU fallback).
"""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--rows", "50000", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg3") and d["config"]["k"] == 100
    base = d["cpu_baseline"]
    assert base["kind"] in ("port", "reference") and base["cores"] >= 1 and base["value"] == d["value"] and base["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_product_arm_needs_the_gpu():
    import torch

    if torch.cuda.is_available():
        return  # covered by the GPU suite and the driver's own bench run
    r = _run(["--rows", "50000", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0, "bench.py must not produce a number without a CUDA device"
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{") and '"value"' in ln]
