"""GPU parity test of the shard-merge kernel (b200ret_merge_topk through ops.merge_topk) against the numpy restatement of its
contract used by the gloo test (tests/test_dist_gloo.py::merge_rows_reference): [G, Q, k] per-shard rows -> [Q, k] under the
total order (score desc, doc id asc), padding rows (id -1, score -inf) dropped.  Bit-exact (ids, score bits, counts)."""
import numpy as np
import pytest
import torch

from scaling_retriever_b200 import ops
from test_dist_gloo import merge_rows_reference, pack_keys_reference, unpack_keys_reference

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


def test_merge_entry_rejects_too_many_candidates_and_ops_merges_in_passes(cuda):
    """The C entry points are bounded by shared memory (b200ret_merge_max_shards); ops.merge_topk / ops.merge_keys group the
    shards into passes, so 8 shards x k = 4096 and 40 x 1000 work."""
    from scaling_retriever_b200 import _lib
    lib = _lib.load()
    assert ops.merge_max_shards(1000) >= 16 and 2 <= ops.merge_max_shards(4096) < 8
    s = torch.zeros((40, 1, 1000), dtype=torch.float32, device=cuda)
    i = torch.zeros((40, 1, 1000), dtype=torch.int64, device=cuda)
    o = torch.zeros((1, 1000), dtype=torch.int64, device=cuda)
    assert lib.b200ret_merge_keys(i.data_ptr(), 40, 1, 1000, o.data_ptr(), None) != 0
    for g, q, k in [(8, 5, 4096), (40, 3, 1000)]:
        scores, ids = make_rows(g, q, k, n_docs=400_000, seed=g + k)
        ref = merge_rows_reference(scores, ids, k)
        out = ops.merge_topk(torch.as_tensor(scores).to(cuda), torch.as_tensor(ids).to(cuda), k)
        assert np.array_equal(out[1].cpu().numpy(), ref[1]) and np.array_equal(out[0].cpu().numpy().view(np.uint32), ref[0].view(np.uint32))
        keys = ops.pack_keys(torch.as_tensor(scores).to(cuda), torch.as_tensor(ids).to(cuda))
        u = ops.unpack_keys(ops.merge_keys(keys, k), k)
        assert np.array_equal(u[1].cpu().numpy(), ref[1]) and np.array_equal(u[2].cpu().numpy(), ref[2])


def test_merge_topk_keeps_64_bit_ids(cuda):
    """ids >= 2^31 (they used to be truncated): merge_topk orders by position and copies the id through."""
    g, q, k = 4, 7, 50
    scores, ids = make_rows(g, q, k, n_docs=200_000, seed=5)
    big = np.where(ids >= 0, ids + (1 << 33), -1)
    out_s, out_i, out_c = ops.merge_topk(torch.as_tensor(scores).to(cuda), torch.as_tensor(big).to(cuda), k)
    ref_s, ref_i, ref_c = merge_rows_reference(scores, big, k)
    assert np.array_equal(out_i.cpu().numpy(), ref_i) and np.array_equal(out_c.cpu().numpy(), ref_c)
    assert np.array_equal(out_s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))


@pytest.mark.parametrize("g,q,k,ties", [(2, 33, 100, 0), (8, 17, 1000, 0), (8, 9, 1000, 7), (3, 5, 37, 2)])
def test_packed_key_merge_matches_reference(cuda, g, q, k, ties):
    """pack_keys -> merge_keys -> unpack_keys (the exchange path of shard.merge_shards) == the reference merge, bit for bit;
    pack/unpack also against their numpy restatements (ids up to 2^32 - 2)."""
    scores, ids = make_rows(g, q, k, n_docs=200_000, seed=g * 77 + k, tie_levels=ties)
    ids = np.where(ids >= 0, ids + (2 ** 32 - 2 - 200_000), -1)          # the top of the id range a key can carry
    d_s, d_i = torch.as_tensor(scores).to(cuda), torch.as_tensor(ids).to(cuda)
    keys = ops.pack_keys(d_s, d_i)
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), pack_keys_reference(scores, ids))
    merged = ops.merge_keys(keys, k)
    out_s, out_i, out_c = ops.unpack_keys(merged, k)
    ref_s, ref_i, ref_c = merge_rows_reference(scores, ids, k)
    assert np.array_equal(out_c.cpu().numpy(), ref_c)
    assert np.array_equal(out_i.cpu().numpy(), ref_i)
    assert np.array_equal(out_s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))
    u_s, u_i, u_c = unpack_keys_reference(merged.cpu().numpy().view(np.uint64))
    assert np.array_equal(u_i, ref_i) and np.array_equal(u_c, ref_c)
