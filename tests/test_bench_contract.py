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


def test_reference_arm_default_line_covers_sparse_and_dense_with_all_cores_under_torchrun_env():
    """Default workload = both halves of the metric in ONE line (dense as a sub-record); the thread count is the box's cores even
    with the OMP_NUM_THREADS=1 that torch.distributed.run exports to every rank; `config` is built from
    the command line only, so the B200 arm of the same command prints the identical dict (`same_config`)."""
    args = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--n-docs", "3000", "--n-queries", "16"]
    proc = run_bench(args, env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "1", "LOCAL_RANK": "0"})
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == cores
    assert d["dense"]["impl"] == "reference" and d["dense"]["cpu_baseline"]["cores"] == cores and d["dense"]["value"] > 0
    assert "dense_4096" not in d
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(n_docs=3000, n_queries=16, gpus=1, workload="both", dim=2048)
    assert d["config"] == bench.sparse_config(ns) and d["dense"]["config"] == bench.dense_config(ns, 2048)
    assert bench.dense_dims(argparse.Namespace(workload="both", gpus=8, dim=2048)) == [2048, 4096]
