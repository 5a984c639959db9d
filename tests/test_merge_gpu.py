"""GPU parity test of the shard-merge kernel (b200ret_merge_topk through ops.merge_topk) against the numpy restatement of its
contract used by the gloo test (tests/test_dist_gloo.py::merge_rows_reference): [G, Q, k] per-shard rows -> [Q, k] under the
total order (score desc, doc id asc), padding rows (id -1, score -inf) dropped.  Bit-exact (ids, score bits, counts)."""
import numpy as np
import pytest
import torch

from scaling_retriever_b200 import ops
from test_dist_gloo import merge_rows_reference

pytestmark = pytest.mark.gpu


def make_rows(g, q, k, n_docs, seed, tie_levels=0):
    """Per-shard rows as the search kernels emit them: sorted by (score desc, id asc), unique global ids, ragged live counts."""
    rng = np.random.default_rng(seed)
    scores = np.full((g, q, k), -np.inf, dtype=np.float32)
    ids = np.full((g, q, k), -1, dtype=np.int64)
    per = n_docs // g
    for gi in range(g):
        for qi in range(q):
            n_live = int(rng.choice([0, 1, k // 2, k, k, k]))
            n_live = min(n_live, per)
            docs = rng.choice(per, size=n_live, replace=False).astype(np.int64) + gi * per
            s = rng.random(n_live, dtype=np.float32) * 10
            if tie_levels:
                s = np.floor(s * tie_levels / 10).astype(np.float32)          # many equal scores: the id breaks the tie
            order = np.lexsort((docs, -s.astype(np.float64)))
            scores[gi, qi, :n_live], ids[gi, qi, :n_live] = s[order], docs[order]
    return scores, ids


@pytest.mark.parametrize("g,q,k,ties", [(1, 5, 10, 0), (2, 33, 100, 0), (8, 17, 1000, 0), (8, 9, 1000, 7), (4, 6, 4096, 3),
                                        (16, 3, 1000, 0)])
def test_merge_topk_matches_reference(cuda, g, q, k, ties):
    scores, ids = make_rows(g, q, k, n_docs=200_000, seed=g * 1000 + k, tie_levels=ties)
    out_s, out_i, out_c = ops.merge_topk(torch.as_tensor(scores).to(cuda), torch.as_tensor(ids).to(cuda), k)
    ref_s, ref_i, ref_c = merge_rows_reference(scores, ids, k)
    assert np.array_equal(out_c.cpu().numpy(), ref_c)
    assert np.array_equal(out_i.cpu().numpy(), ref_i)
    assert np.array_equal(out_s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))


def test_merge_topk_rejects_too_many_candidates(cuda):
    s = torch.zeros((17, 1, 1000), dtype=torch.float32, device=cuda)
    i = torch.zeros((17, 1, 1000), dtype=torch.int64, device=cuda)
    with pytest.raises(Exception):
        ops.merge_topk(s, i, 1000)
