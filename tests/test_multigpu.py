"""Runs tests/multigpu_check.py under torchrun when at least two GPUs are visible (skipped on 1-GPU boxes; the CPU-side
world_size-2 plumbing is covered by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_search_equals_single_gpu():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "multigpu_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "multigpu_check ok" in proc.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_devices_in_one_process():
    """Kernel attributes (opt-in shared memory) are per device: the same process must be able to search on cuda:0 and cuda:1."""
    from scaling_retriever_b200 import ops, synth
    results = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            rows, cols, vals = synth.gen_sparse_docs(20000, n_terms=500, mean_nnz=20, seed=3, device=dev)
            index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, 500, 20000)
            q_off, q_t, q_w = synth.gen_sparse_queries(40, n_terms=500, mean_nnz=10, seed=4, device=dev)
            s, i, c = ops.sparse_search(index, q_off, q_t, q_w, 50, 0.0)
            corpus = synth.gen_dense(3000, 128, seed=5, device=dev, dtype=torch.bfloat16)
            q16 = ops.f32_to_bf16(synth.gen_dense(20, 128, seed=6, device=dev))
            ds, di, _ = ops.dense_search(corpus, q16, 10)
            results.append((s.cpu(), i.cpu(), c.cpu(), ds.cpu(), di.cpu()))
    for a, b in zip(*results):
        assert torch.equal(a, b)
