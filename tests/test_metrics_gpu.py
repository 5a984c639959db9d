"""On-GPU rank metrics against the restatement of the reference's mrr_k / recall_k (oracle/metrics_oracle.py,
utils/metrics.py:13-42 + trec_eval's recip_rank / recall_<k> definitions)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle
from scaling_retriever_b200 import metrics
from scaling_retriever_b200.results import ExternalIds, LazyRun

pytestmark = pytest.mark.gpu


def make_case(seed, nq, k, n_docs, string_ids):
    rng = np.random.default_rng(seed)
    ids = np.stack([rng.choice(n_docs, size=k, replace=False) for _ in range(nq)]).astype(np.int64)
    scores = -np.sort(-rng.random((nq, k), dtype=np.float32) * 20, axis=1)          # distinct scores, descending rows
    counts = rng.choice([0, 1, k // 2, k, k], size=nq).astype(np.int32)
    ext = [f"D{3 * i}" for i in range(n_docs)] if string_ids else range(n_docs)
    qids = [f"q{i}" for i in range(nq)]
    run = LazyRun(qids, ids, scores, counts, ExternalIds(ext))
    qrel = {}
    for i, q in enumerate(qids):
        if i % 7 == 3:
            continue                                   # query without judgements: not evaluated
        judged = {}
        for d in rng.choice(n_docs, size=int(rng.integers(1, 6)), replace=False):
            judged[str(ext[int(d)])] = int(rng.integers(0, 3))          # 0 = judged non-relevant
        for j in rng.choice(k, size=int(rng.integers(0, 4)), replace=False):      # some relevant docs really are retrieved
            judged[str(ext[int(ids[i, j])])] = 1
        if i % 5 == 0:
            judged["not-in-collection"] = 2            # counts in the recall denominator only
        qrel[q] = judged
    return run, qrel


@pytest.mark.parametrize("string_ids", [False, True])
def test_mrr_and_recall_match_the_reference_definitions(cuda, string_ids):
    run, qrel = make_case(5, nq=83, k=200, n_docs=5000, string_ids=string_ids)
    eager = run.to_dict()
    for cut in (10, 100):
        assert abs(metrics.mrr_k(run, qrel, cut) - metrics_oracle.mrr_k(eager, qrel, cut)) < 1e-6
    for cut in (10, 100, 1000):
        assert abs(metrics.recall_k(run, qrel, cut) - metrics_oracle.recall_k(eager, qrel, cut)) < 1e-6
    per_q = metrics.mrr_k(run, qrel, 10, agg=False)
    ref = metrics_oracle.recip_rank(metrics_oracle.truncate_run(eager, 10), qrel)
    assert set(per_q) == set(ref)                      # queries with no retrieved doc or no judgement are not evaluated
    for q in ref:
        assert abs(per_q[q]["recip_rank"] - ref[q]) < 1e-6


def test_hand_computed_example(cuda):
    ids = np.array([[4, 2, 9, 1], [7, 8, 3, 0]], dtype=np.int64)
    scores = np.array([[4, 3, 2, 1], [9, 8, 7, 6]], dtype=np.float32)
    run = LazyRun(["a", "b"], ids, scores, None, ExternalIds(range(10)))
    qrel = {"a": {"9": 1, "1": 1, "5": 1}, "b": {"6": 1}, "c": {"1": 1}}
    assert metrics.mrr_k(run, qrel, 10) == pytest.approx((1 / 3 + 0) / 2)
    assert metrics.mrr_k(run, qrel, 2) == pytest.approx(0.0)
    assert metrics.recall_k(run, qrel, 3) == pytest.approx((1 / 3 + 0) / 2)
    assert metrics.recall_k(run, qrel, 1000) == pytest.approx((2 / 3 + 0) / 2)
