"""bench.py contract checks that run without a GPU: the reference arm (CPU port of the reference's path) prints ONE JSON line
with the keys the driver reads, non-zero ranks of a multi-rank launch stay silent, and the B200 arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                          env=e, cwd=ROOT)


@pytest.mark.parametrize("workload", ["sparse", "dense"])
def test_reference_arm_prints_one_contract_line(workload):
    args = ["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0", "--n-docs", "3000", "--n-queries", "16"]
    proc = run_bench(args)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "top1000_qps" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_nonzero_rank_is_silent():
    proc = run_bench(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert proc.returncode == 0 and proc.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_b200_arm_fails_loudly_without_cuda():
    proc = run_bench(["--steps", "1", "--warmup", "0", "--n-docs", "1000", "--n-queries", "4"])
    assert proc.returncode != 0 and "CUDA" in (proc.stderr + proc.stdout)
