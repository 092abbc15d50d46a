"""bench.py on a machine without a GPU: the reference arm runs (it is the CPU oracle), our arm refuses."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "flat1m", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "entities_per_second_trs_world_mvp_cull" and d["unit"] == "M entities/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.0 < d["config"]["visible_fraction"] < 1.0


def test_reference_arm_under_torchrun_only_rank_0_works():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--workload", "flat1m", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
