"""GPU parity tests of the dense flat-IP path (tcgen05 GEMM + fused top-k) through the C ABI, against the fp32 CPU
restatement of faiss.IndexFlatIP.search (oracle/dense_oracle.py; PARITY UNPINNED by reference vectors — faiss is absent
from the reference tree and this image, see its header).

Tolerances (BASELINE.json north_star): scores <= 1e-2 relative vs the fp32 reference on the ORIGINAL inputs; against the
fp32 reference on the bf16-ROUNDED inputs the only difference left is the fp32 summation order (<= 1e-5 absolute for
unit-norm rows), and ids must be identical except where neighbouring scores are closer than that."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle
from scaling_retriever_b200 import ops, synth

pytestmark = pytest.mark.gpu


def check_against_oracle(scores, ids, counts, docs16, q16, k, n_docs, id_base=0):
    o_scores, o_ids = dense_oracle.flat_ip_search(docs16, q16, k)
    kk = min(k, n_docs)
    assert np.all(counts == kk)
    assert np.all(ids[:, kk:] == -1) and np.all(np.isneginf(scores[:, kk:]))
    assert np.all(np.diff(scores[:, :kk], axis=1) <= 0)
    np.testing.assert_allclose(scores[:, :kk], o_scores[:, :kk], rtol=0, atol=2e-5)
    full = q16 @ docs16.T
    for qi in range(len(q16)):
        got = ids[qi, :kk] - id_base
        assert len(set(got.tolist())) == kk and got.min() >= 0 and got.max() < n_docs
        np.testing.assert_allclose(scores[qi, :kk], full[qi][got], rtol=0, atol=2e-5)      # every id carries its own score
        differ = got != o_ids[qi, :kk]
        if differ.any():   # only allowed between near-equal neighbours (summation-order noise)
            assert np.all(np.abs(scores[qi, :kk][differ] - o_scores[qi, :kk][differ]) <= 2e-5)
        kth = o_scores[qi, kk - 1]
        clear = full[qi] > kth + 4e-5
        assert set(np.nonzero(clear)[0].tolist()) <= set(got.tolist())


@pytest.mark.parametrize("n_docs,dim,n_queries,k", [
    (5000, 128, 37, 100),          # ragged: queries and docs not multiples of the 256-wide tiles
    (70000, 256, 300, 1000),       # several rounds of the doubling schedule, two query blocks
    (300, 64, 5, 1000),            # fewer docs than k -> padded rows
    (16384 + 17, 2048, 130, 10),   # Lion-DS-1B row width
])
def test_dense_search_matches_fp32_oracle(cuda, n_docs, dim, n_queries, k):
    docs = synth.gen_dense(n_docs, dim, seed=3, device=cuda)
    queries = synth.gen_dense(n_queries, dim, seed=4, device=cuda)
    d16, q16 = ops.f32_to_bf16(docs), ops.f32_to_bf16(queries)
    assert torch.equal(d16, docs.to(torch.bfloat16))                     # cast kernel == round-to-nearest-even
    scores, ids, counts = ops.dense_search(d16, q16, k)
    torch.cuda.synchronize()
    check_against_oracle(scores.cpu().numpy(), ids.cpu().numpy(), counts.cpu().numpy(), d16.float().cpu().numpy(),
                         q16.float().cpu().numpy(), k, n_docs)
    # vs the fp32 reference on the unrounded inputs: <= 1e-2 relative
    exact = dense_oracle.flat_ip_search(docs.cpu().numpy(), queries.cpu().numpy(), k)[0]
    kk = min(k, n_docs)
    np.testing.assert_allclose(scores.cpu().numpy()[:, :kk], exact[:, :kk], rtol=1e-2, atol=2e-3)


def test_dense_search_planted_neighbours_and_id_base(cuda):
    """Well separated top ranks: each query has 20 planted near-duplicates that must come out first, in order."""
    n_docs, dim, n_queries, k = 40000, 512, 64, 50
    g = torch.Generator(device=cuda).manual_seed(11)
    docs = synth.gen_dense(n_docs, dim, seed=5, device=cuda)
    queries = synth.gen_dense(n_queries, dim, seed=6, device=cuda)
    planted = torch.randperm(n_docs, device=cuda, generator=g)[:n_queries * 20].view(n_queries, 20)
    for j in range(20):
        mix = (1.0 - 0.03 * j)
        v = mix * queries + (1 - mix) * synth.gen_dense(n_queries, dim, seed=100 + j, device=cuda)
        docs[planted[:, j]] = torch.nn.functional.normalize(v, dim=1)
    d16, q16 = docs.to(torch.bfloat16), queries.to(torch.bfloat16)
    scores, ids, counts = ops.dense_search(d16, q16, k, doc_id_base=1_000_000)
    # the 20 planted docs (cosine >= 0.6) are the 20 best of every query; their exact order is checked against the oracle
    assert torch.equal(ids[:, :20].sort(dim=1).values, (planted + 1_000_000).sort(dim=1).values)
    assert torch.equal(ids[:, 0], planted[:, 0] + 1_000_000)
    check_against_oracle(scores.cpu().numpy(), ids.cpu().numpy(), counts.cpu().numpy(), d16.float().cpu().numpy(),
                         q16.float().cpu().numpy(), k, n_docs, id_base=1_000_000)


def test_dense_search_adversarial_order_uses_safe_schedule(cuda):
    """Scores increase with the row id: every later doc beats the running k-th score, the candidate lists overflow in the
    doubling rounds and the safe re-run must still return the exact top-k."""
    n_docs, dim, k = 60000, 64, 10
    base = torch.zeros(dim, device=cuda)
    base[0] = 1.0
    docs = base.repeat(n_docs, 1) * (torch.arange(n_docs, device=cuda, dtype=torch.float32).unsqueeze(1) / n_docs + 1.0)
    queries = base.repeat(3, 1)
    d16, q16 = docs.to(torch.bfloat16), queries.to(torch.bfloat16)
    scores, ids, counts = ops.dense_search(d16, q16, k)
    o_scores, o_ids = dense_oracle.flat_ip_search(d16.float().cpu().numpy(), q16.float().cpu().numpy(), k)
    assert np.array_equal(scores.cpu().numpy(), o_scores)           # products of bf16 values: exact in fp32
    assert np.array_equal(ids.cpu().numpy(), o_ids)                 # ties (bf16 plateaus) resolve to the lowest row id on both sides


def test_dense_rejects_bad_shapes(cuda):
    from scaling_retriever_b200._lib import B200RetError
    d = torch.zeros((10, 96), dtype=torch.bfloat16, device=cuda)
    with pytest.raises(B200RetError):
        ops.dense_search(d, d[:2].contiguous(), 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dense_search(d.cpu(), d.cpu(), 5)


def test_dense_search_matches_exact_construction_golden(cuda):
    """The exact, tie-free construction of tests/golden/make_golden_dense.py: ids and scores of the tcgen05 kernel must equal
    the stored int64-computed result bit for bit (single search, and sharded over 3 doc ranges + merge)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dense_golden.npz"))
    k = int(g["k"])
    docs = ops.f32_to_bf16(torch.as_tensor(g["docs"].astype(np.float32)).to(cuda))
    queries = ops.f32_to_bf16(torch.as_tensor(g["queries"]).to(cuda))
    assert torch.equal(docs.float().cpu(), torch.as_tensor(g["docs"].astype(np.float32)))     # bf16 holds the inputs exactly
    assert torch.equal(queries.float().cpu(), torch.as_tensor(g["queries"]))
    s, i, c = ops.dense_search(docs, queries, k)
    assert np.array_equal(i.cpu().numpy(), g["top_ids"])
    assert np.array_equal(s.cpu().numpy().view(np.uint32), g["top_scores"].view(np.uint32))
    assert (c.cpu().numpy() == k).all()
    bounds = [0, 4100, 8000, docs.shape[0]]
    parts = [ops.dense_search(docs[a:b].contiguous(), queries, k, doc_id_base=a) for a, b in zip(bounds[:-1], bounds[1:])]
    ms, mi, _ = ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    assert np.array_equal(mi.cpu().numpy(), g["top_ids"]) and np.array_equal(ms.cpu().numpy().view(np.uint32), g["top_scores"].view(np.uint32))


def test_dense_search_large_k(cuda):
    """k = 4096 (B200RET_MAX_K): candidate capacity k + 5 k; exact-construction style inputs so ids and scores are bit-exact."""
    g = torch.Generator().manual_seed(9)
    n, d, nq, k = 30000, 64, 16, 4096
    docs = torch.randint(-1, 2, (n, d), generator=g).float()
    ids = torch.arange(n)
    docs[:, d - 2] = (ids % 256).float()
    docs[:, d - 1] = (ids // 256).float()
    qs = torch.randint(-1, 2, (nq, d), generator=g).float()
    qs[:, d - 2] = 2.0 ** -16
    qs[:, d - 1] = 2.0 ** -8
    s, i, c = ops.dense_search(ops.f32_to_bf16(docs.to(cuda)), ops.f32_to_bf16(qs.to(cuda)), k)
    o_s, o_i = dense_oracle.flat_ip_search(docs.numpy(), qs.numpy(), k)
    assert np.array_equal(i.cpu().numpy(), o_i) and np.array_equal(s.cpu().numpy().view(np.uint32), o_s.view(np.uint32))
    assert (c.cpu().numpy() == k).all()
