"""bench.py contract (CPU side): the reference arm prints ONE JSON line on the real stdout with the keys the driver reads, and the
product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(monkeypatch, capfd):
    sys.path.insert(0, ROOT)
    import bench
    orig = bench.run_cpu_arm
    monkeypatch.setattr(bench, "run_cpu_arm", lambda steps, warmup, workload="ppo", threads=None, T_s=None: orig(1, 0, workload=workload, threads=threads, T_s=1))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"])
    monkeypatch.delenv("RANK", raising=False)
    bench.main()
    out = capfd.readouterr().out.strip().splitlines()
    lines = [l for l in out if l.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("cleanba_ppo a0-l0-d1")
    assert set(("value", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_product_arm_refuses_to_run_without_cuda():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
