"""not gpu: the reference arm of bench.py (`--impl reference`: the CPU restatement of the reference path on the host cores)
prints ONE JSON line with the keys the driver reads, and the engine arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "clip_frames_per_sec" and j["unit"] == "clip-frames/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0
    assert j["value"] > 0 and abs(j["value"] - 16 / (j["ms_per_step"] * 1e-3)) < 1e-6 * j["value"] + 1e-9
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["vs_baseline"] is None and j["data"] == "synthetic"


def test_engine_arm_needs_a_gpu():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]      # no number is ever printed from a fallback


def test_reference_arm_other_configs():
    """--config selects the BASELINE.json workload for both arms: the torch-stack flicker (c4) and sparse (c5) configs"""
    for cfgname, arch in (("c4", "r2plus1d_18"), ("c5", "r3d_18")):
        r = _run("--impl", "reference", "--config", cfgname, "--steps", "1", "--warmup", "0", "--frames", "4")
        assert r.returncode == 0, r.stderr[-2000:]
        j = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
        assert j["impl"] == "reference" and j["config"]["name"] == cfgname and j["config"]["arch"] == arch
        assert j["value"] > 0 and j["config"]["frames"] == 4
