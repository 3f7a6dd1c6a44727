"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly one JSON line with
the tier's keys; the default arm fails loudly without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"] == "range-image frames/s (fwd+bwd, 64x2650)" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg-5 rangedet_veh_wo_aug_4_18e train step") and d["gpu_launches"] == 0
    assert "1 frame" in d["cpu_baseline"]["sample"]


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--quick", timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
