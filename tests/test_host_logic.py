"""Host-side logic that needs no GPU: query packing, shard plans, dict-of-arrays views, file-format helpers, the
synthetic generators — and that the product path refuses to run without CUDA (no CPU fallback)."""
import json
import os

import numpy as np
import pytest
import torch

from scaling_retriever_b200 import ops, shard, synth
from scaling_retriever_b200.indexer import DenseFlatIndexer, SparseRetrieval, pack_queries
from scaling_retriever_b200.inverted_index import IndexDictOfArray, _TermArrays, _resize_offsets
from scaling_retriever_b200.utils import obtain_doc_vec_dir_files


def test_pack_queries_roundtrip():
    vecs = [(np.array([1, 5, 9], np.int32), np.array([.1, .2, .3], np.float32)), (np.zeros(0, np.int32), np.zeros(0, np.float32)),
            (np.array([7], np.int64), np.array([2.0], np.float64))]
    off, t, w = pack_queries(vecs)
    assert off.tolist() == [0, 3, 3, 4] and t.dtype == np.int32 and w.dtype == np.float32
    assert t.tolist() == [1, 5, 9, 7] and np.allclose(w, [.1, .2, .3, 2.0])
    off, t, w = pack_queries([])
    assert off.tolist() == [0] and len(t) == 0


@pytest.mark.parametrize("n,g", [(8841823, 8), (10, 4), (3, 8), (0, 2), (100, 1)])
def test_shard_plan_partitions_rows(n, g):
    plan = shard.ShardPlan(n, g)
    bounds = [plan.bounds(r) for r in range(g)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a, b), (c, d) in zip(bounds[:-1], bounds[1:]):
        assert b == c and a <= b
    if n:
        for doc in (0, n // 2, n - 1):
            lo, hi = plan.bounds(plan.owner(doc))
            assert lo <= doc < hi


def test_shard_sparse_csr_on_cpu_tensors():
    rows, cols, vals = synth.gen_sparse_docs(1000, n_terms=50, mean_nnz=6, seed=1)
    order = torch.argsort(cols.long() * 1000 + rows.long())
    counts = torch.bincount(cols.long(), minlength=50)
    off = torch.zeros(51, dtype=torch.int64)
    off[1:] = torch.cumsum(counts, 0)
    ids, w = rows[order], vals[order]
    parts = [shard.shard_sparse_csr(off, ids, w, *shard.ShardPlan(1000, 3).bounds(r)) for r in range(3)]
    assert sum(p[1].numel() for p in parts) == ids.numel()
    for r, (o, i, v) in enumerate(parts):
        lo, hi = shard.ShardPlan(1000, 3).bounds(r)
        assert int(o[-1]) == i.numel() and (i.numel() == 0 or (int(i.min()) >= 0 and int(i.max()) < hi - lo))
        t = 7
        sel = (ids[off[t]:off[t + 1]] >= lo) & (ids[off[t]:off[t + 1]] < hi)
        assert torch.equal(i[o[t]:o[t + 1]].long(), ids[off[t]:off[t + 1]][sel].long() - lo)


def test_term_arrays_view_behaves_like_the_reference_dict():
    off = np.array([0, 2, 2, 5])
    view = _TermArrays(off, np.arange(5, dtype=np.int32), np.array([0, 2]))
    assert len(view) == 2 and list(view) == [0, 2] and 1 not in view and 2 in view
    assert view[2].tolist() == [2, 3, 4]
    with pytest.raises(KeyError):
        view[1]
    assert _resize_offsets(off, 5).tolist() == [0, 2, 2, 5, 5, 5]
    with pytest.raises(ValueError):
        _resize_offsets(off, 2)


def test_obtain_doc_vec_dir_files(tmp_path):
    for r in range(2):
        for c in range(2):
            np.save(tmp_path / f"embs_{r}_{c}.npy", np.zeros((1, 2), np.float32))
            np.save(tmp_path / f"ids_{r}_{c}.npy", np.zeros(1, np.int64))
    json.dump({"nranks": 2, "num_chunks": 2, "index_path": "x"}, open(tmp_path / "plan.json", "w"))
    vec, ids = obtain_doc_vec_dir_files(str(tmp_path))
    assert [os.path.basename(v) for v in vec] == ["embs_0_0.npy", "embs_0_1.npy", "embs_1_0.npy", "embs_1_1.npy"]
    assert len(ids) == 4


def test_synthetic_generators_are_seeded_and_shard_invariant():
    a = synth.gen_sparse_docs(3000, n_terms=500, mean_nnz=10, seed=5)
    b = synth.gen_sparse_docs(3000, n_terms=500, mean_nnz=10, seed=5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    rows, cols, _ = a
    key = rows.long() * 500 + cols.long()
    assert bool((key[1:] > key[:-1]).all())          # row-major, unique (doc, term)
    part = synth.gen_sparse_docs(3000, n_terms=500, mean_nnz=10, seed=5, doc_lo=1000, doc_hi=2000)
    m = (rows >= 1000) & (rows < 2000)
    assert torch.equal(part[0], rows[m]) and torch.equal(part[1], cols[m])
    q_off, q_t, q_w = synth.gen_sparse_queries(20, n_terms=500, mean_nnz=5)
    assert q_off[0] == 0 and int(q_off[-1]) == q_t.numel() == q_w.numel() and bool((q_w > 0).all())
    d = synth.gen_dense(100, 16, seed=3)
    assert torch.allclose(d.norm(dim=1), torch.ones(100), atol=1e-5)
    assert torch.equal(synth.gen_dense(100, 16, seed=3, row_lo=10, row_hi=20), d[10:20])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sparse_search(None, torch.zeros(2, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), torch.zeros(0), 10)
    idx = IndexDictOfArray(index_path=None, dim_voc=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        idx.add_batch_document(np.array([0]), np.array([1]), np.array([1.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SparseRetrieval(torch.nn.Linear(1, 1), {"out_dir": str(tmp_path)}, 8, 0, index_d={"index": idx, "ids_mapping": {}})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DenseFlatIndexer().init_index(8)


def test_select_topk_static_keeps_reference_semantics():
    idx = np.arange(6)
    neg = -np.array([1, 2, 2, 2, 2, 3], dtype=np.float32)
    sel, sc = SparseRetrieval.select_topk(idx, neg, 3)
    assert len(sel) == 3 and 5 in sel and np.all(sc >= 2)
    sel, sc = SparseRetrieval.select_topk(idx, neg, 10)
    assert np.array_equal(sel, idx) and np.array_equal(sc, -neg)


def test_metrics_oracle_hand_example():
    """oracle/metrics_oracle.py (restatement of utils/metrics.py:13-42 + trec_eval definitions) on a hand-computed case."""
    from oracle import metrics_oracle as m
    run = {"a": {"4": 4.0, "2": 3.0, "9": 2.0, "1": 1.0}, "b": {"7": 9.0, "8": 8.0, "3": 7.0, "0": 6.0}}
    qrel = {"a": {"9": 1, "1": 1, "5": 1}, "b": {"6": 1}, "c": {"1": 1}}
    assert m.mrr_k(run, qrel, 10) == (1 / 3 + 0) / 2 and m.mrr_k(run, qrel, 2) == 0.0
    assert m.recall_k(run, qrel, 3) == (1 / 3) / 2 and m.recall_k(run, qrel, 1000) == (2 / 3) / 2
    assert m.truncate_run(run, 2)["a"] == {"4": 4.0, "2": 3.0}
