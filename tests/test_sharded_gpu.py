"""The sharded search entry points (b200ret_sparse_search_sharded / b200ret_dense_search_sharded, include/b200ret.h (3b)) on ONE
GPU: the G doc-range shards are searched by G host threads on G streams and the per-round tau exchange is a thread barrier +
element-wise minimum (sharded_helpers.LocalExchange) instead of an NCCL all-reduce.  The merged rows must be bit-identical to
the unsharded search.  (The same checks over NCCL, one rank per GPU: tests/multigpu_check.py.)"""
import pytest
import torch

import sharded_helpers as sh
from scaling_retriever_b200 import _lib, ops, shard, synth

pytestmark = pytest.mark.gpu


def _merge(rows, k):
    s = torch.stack([r[0] for r in rows])
    i = torch.stack([r[1] for r in rows])
    return ops.merge_topk(s, i, k)


def _same(got, want):
    return (torch.equal(got[1], want[1]) and torch.equal(got[2], want[2])
            and torch.equal(got[0].view(torch.int32), want[0].view(torch.int32)))


@pytest.mark.parametrize("world,k,threshold", [(2, 100, 0.0), (3, 1000, 0.0), (4, 37, 1.5), (8, 1000, 0.0)])
def test_sparse_tau_exchange_threads(cuda, world, k, threshold):
    lib = _lib.load()
    growth = lib.b200ret_exchange_growth(world)
    assert growth == max(4, world + 1)
    bd = ops.block_docs()
    n_terms = 3000
    n_docs = world * 2 * growth * bd + 1        # shards 0..G-2: 2g blocks + 1 doc (3 rounds); the last one: 2g blocks (2 rounds)
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=20, seed=15, device=cuda)
    q_off, q_t, q_w = synth.gen_sparse_queries(150, n_terms=n_terms, mean_nnz=12, seed=6, device=cuda)
    want = ops.sparse_search(ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs), q_off, q_t, q_w, k, threshold)
    plan = shard.ShardPlan(n_docs, world)
    parts, los = [], []
    for g in range(world):
        lo, hi = plan.bounds(g)
        keep = (rows >= lo) & (rows < hi)
        parts.append(ops.SparseDeviceIndex.from_coo((rows[keep] - lo).contiguous(), cols[keep].contiguous(), vals[keep].contiguous(),
                                                    n_terms, hi - lo))
        los.append(lo)
    n_ex = lib.b200ret_sparse_exchange_rounds(plan.per_shard, world)
    assert n_ex == 2
    shared = sh.LocalExchange.Shared(world)
    exs = [sh.LocalExchange(shared, n_ex, growth, world, cuda) for _ in range(world)]
    for _ in range(2):                          # exchange objects are reused across searches
        got = sh.search_shards_threaded(
            lambda g: ops.sparse_search(parts[g], q_off, q_t, q_w, k, threshold, doc_id_base=los[g], exchange=exs[g]), world, cuda)
        plain = [ops.sparse_search(parts[g], q_off, q_t, q_w, k, threshold, doc_id_base=los[g]) for g in range(world)]
        # with the exchanged bound a shard keeps only what can still reach the global top-k: a subset of its plain local rows
        assert all(bool((got[g][2] <= plain[g][2]).all()) for g in range(world))
        assert _same(_merge(got, k), want)
    assert all(e.rounds_seen == 2 * n_ex for e in exs)      # the shard with fewer rounds still took part in every exchange


def test_sparse_sharded_overflow_tiers_threads(cuda):
    """A sharded search whose exchange-driven schedule overflows a candidate list falls back to the shard's own bounds with the
    plain schedule (middle tier) and, if that overflows too, to the safe schedule; launch counts pin which tier ran."""
    world, k = 2, 1000
    (l_rows, l_cols, l_vals), n_shard, n_terms, qa, qb = sh.rising_score_shard(cuda, k)
    g_rows = torch.cat([l_rows + g * n_shard for g in range(world)])
    full = ops.SparseDeviceIndex.from_coo(g_rows, l_cols.repeat(world), l_vals.repeat(world), n_terms, world * n_shard)
    part = ops.SparseDeviceIndex.from_coo(l_rows, l_cols, l_vals, n_terms, n_shard)
    shared = sh.LocalExchange.Shared(world)
    exs = [sh.LocalExchange(shared, 1, 100000, world, cuda) for _ in range(world)]     # round 2 = the rest of the shard
    for (off, t, w), launches_per_shard in ((qa, sh.LAUNCHES_MIDDLE_TIER), (qb, sh.LAUNCHES_SAFE_TIER)):
        want = ops.sparse_search(full, off, t, w, k, 0.0)
        ops.profile_enable(True)
        ops.profile_read(ops.PROF_SPARSE_SCORE)
        try:
            got = sh.search_shards_threaded(
                lambda g: ops.sparse_search(part, off, t, w, k, 0.0, doc_id_base=g * n_shard, exchange=exs[g]), world, cuda)
            _, launches, _ = ops.profile_read(ops.PROF_SPARSE_SCORE)
        finally:
            ops.profile_enable(False)
        assert launches == world * launches_per_shard
        assert _same(_merge(got, k), want)


@pytest.mark.parametrize("world,k", [(2, 200), (3, 1000)])
def test_dense_tau_exchange_threads(cuda, world, k):
    lib = _lib.load()
    growth = lib.b200ret_exchange_growth(world)
    dim, nq = 256, 70
    n_docs = world * 32 * growth * 256 + 1      # 32g + 1 tiles of 256 docs on shards 0..G-2 (3 rounds), 32g on the last one (2)
    docs = synth.gen_dense(n_docs, dim, seed=17, device=cuda, dtype=torch.bfloat16)
    q16 = ops.f32_to_bf16(synth.gen_dense(nq, dim, seed=8, device=cuda))
    want = ops.dense_search(docs, q16, k)
    plan = shard.ShardPlan(n_docs, world)
    n_ex = lib.b200ret_dense_exchange_rounds(plan.per_shard, world)
    assert n_ex == 2
    shared = sh.LocalExchange.Shared(world)
    exs = [sh.LocalExchange(shared, n_ex, growth, world, cuda) for _ in range(world)]
    parts = [docs[plan.bounds(g)[0]:plan.bounds(g)[1]].contiguous() for g in range(world)]
    got = sh.search_shards_threaded(
        lambda g: ops.dense_search(parts[g], q16, k, doc_id_base=plan.bounds(g)[0], exchange=exs[g]), world, cuda)
    merged = _merge(got, k)
    assert torch.equal(merged[1], want[1]) and torch.equal(merged[0].view(torch.int32), want[0].view(torch.int32))
    assert all(e.rounds_seen == n_ex for e in exs)


def test_exchange_struct_is_validated(cuda):
    rows, cols, vals = synth.gen_sparse_docs(5000, n_terms=300, mean_nnz=10, seed=2, device=cuda)
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, 300, 5000)
    q_off, q_t, q_w = synth.gen_sparse_queries(8, n_terms=300, mean_nnz=6, seed=3, device=cuda)
    shared = sh.LocalExchange.Shared(1)
    bad = sh.LocalExchange(shared, 1, 1, 1, cuda)              # growth < 2
    with pytest.raises(_lib.B200RetError):
        ops.sparse_search(index, q_off, q_t, q_w, 10, 0.0, exchange=bad)
    ok = sh.LocalExchange(sh.LocalExchange.Shared(1), 0, 4, 1, cuda)      # one shard, no exchange round: the plain result
    got = ops.sparse_search(index, q_off, q_t, q_w, 10, 0.0, exchange=ok)
    assert _same(got, ops.sparse_search(index, q_off, q_t, q_w, 10, 0.0))
