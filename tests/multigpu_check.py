"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multigpu_check.py

Every rank searches its doc-range shard (sparse and dense), the per-shard top-k rows are all-gathered over NCCL and merged by
the merge_topk kernel; the result on EVERY rank must be identical (ids and score bits) to the unsharded search of the whole
corpus on one GPU.  Driven by tests/test_multigpu.py when >= 2 GPUs are visible.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scaling_retriever_b200 import ops, shard, synth  # noqa: E402
from scaling_retriever_b200.indexer import DenseFlatIndexer, SparseRetrieval  # noqa: E402
from scaling_retriever_b200.inverted_index import IndexDictOfArray  # noqa: E402
import sharded_helpers as sh  # noqa: E402


def check_overflow_tiers(rank, world, dev):
    """Candidate-list overflow of a sharded search WITH the tau exchange (forced by the B200RET_TEST_EXCHANGE_GROWTH hook: the
    second round covers the rest of the shard) on shards whose scores rise with the doc id (sharded_helpers.rising_score_shard):
    the middle tier (the shard's own bounds, no collectives) finishes query set A, set B needs the safe schedule as well.  The
    pattern repeats per shard, so equal scores sit in different shards and the exchange's strict bound is exercised too."""
    from scaling_retriever_b200 import _lib
    k = 1000
    (l_rows, l_cols, l_vals), n_shard, n_terms, qa, qb = sh.rising_score_shard(dev, k)
    g_rows = torch.cat([l_rows + g * n_shard for g in range(world)])
    full = ops.SparseDeviceIndex.from_coo(g_rows, l_cols.repeat(world), l_vals.repeat(world), n_terms, world * n_shard)
    part = ops.SparseDeviceIndex.from_coo(l_rows, l_cols, l_vals, n_terms, n_shard)
    lo = rank * n_shard
    assert shard.ShardPlan(world * n_shard, world).bounds(rank) == (lo, lo + n_shard)
    os.environ["B200RET_TEST_EXCHANGE_GROWTH"] = "100000"
    try:
        assert _lib.load().b200ret_exchange_growth(world) == 100000
        ex = shard.TauExchange("sparse", world * n_shard, dev)
        assert ex.growth == 100000 and ex.n_exchanges == 1
        for name, (off, t, w), want in (("middle tier", qa, sh.LAUNCHES_MIDDLE_TIER), ("safe tier", qb, sh.LAUNCHES_SAFE_TIER)):
            r_s, r_i, r_c = ops.sparse_search(full, off, t, w, k, 0.0)
            ops.profile_enable(True)
            ops.profile_read(ops.PROF_SPARSE_SCORE)
            s, i, c = ops.sparse_search(part, off, t, w, k, 0.0, doc_id_base=lo, exchange=ex)
            _, launches, _ = ops.profile_read(ops.PROF_SPARSE_SCORE)
            ops.profile_enable(False)
            assert launches == want, (name, launches)
            s, i, c = shard.merge_shards(s, i, k, n_docs_total=world * n_shard)
            assert torch.equal(i, r_i) and torch.equal(c, r_c) and torch.equal(s.view(torch.int32), r_s.view(torch.int32)), name
    finally:
        del os.environ["B200RET_TEST_EXCHANGE_GROWTH"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---- sparse: ops level -------------------------------------------------------------------------------------------
    n_docs, n_terms, n_queries, k = 60011, 3000, 150, 100
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=50, seed=5, device=dev)
    q_off, q_t, q_w = synth.gen_sparse_queries(n_queries, n_terms=n_terms, mean_nnz=12, seed=6, device=dev)
    full = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    ref_s, ref_i, ref_c = ops.sparse_search(full, q_off, q_t, q_w, k, 0.0)
    lo, hi = shard.ShardPlan(n_docs, world).bounds(rank)
    keep = (rows >= lo) & (rows < hi)
    part = ops.SparseDeviceIndex.from_coo((rows[keep] - lo).contiguous(), cols[keep].contiguous(), vals[keep].contiguous(),
                                          n_terms, hi - lo)
    s, i, c = ops.sparse_search(part, q_off, q_t, q_w, k, 0.0, doc_id_base=lo)
    loc_s, loc_i = s, i
    s, i, c = shard.merge_shards(loc_s, loc_i, k, n_docs_total=n_docs)      # packed-key all-to-all + per-slice merge + all-gather
    assert torch.equal(i, ref_i) and torch.equal(c, ref_c) and torch.equal(s.view(torch.int32), ref_s.view(torch.int32)), "sparse ops"
    s, i, c = shard.merge_shards_allgather(loc_s, loc_i, k)                  # the 64-bit-id exchange gives the same rows
    assert torch.equal(i, ref_i) and torch.equal(c, ref_c) and torch.equal(s.view(torch.int32), ref_s.view(torch.int32)), "sparse allgather"
    s, i, c = shard.merge_shards(loc_s, loc_i, k, n_docs_total=1 << 33)      # ids beyond a key's 32 bits take that exchange
    assert torch.equal(i, ref_i), "sparse 64-bit fallback"

    # ---- sparse: tau exchange between the rounds (shard.TauExchange), shards with DIFFERENT round counts --------------------
    from scaling_retriever_b200 import _lib
    growth = _lib.load().b200ret_exchange_growth(world)        # docs scored grow max(4, G + 1) times per round with the exchange
    bd = ops.block_docs()
    n2 = world * 2 * growth * bd + 1  # ranks 0..G-2: 2g blocks + 1 doc (3 rounds); last rank: 2g blocks (2 rounds + 1 extra exchange)
    rows2, cols2, vals2 = synth.gen_sparse_docs(n2, n_terms=n_terms, mean_nnz=20, seed=15, device=dev)
    full2 = ops.SparseDeviceIndex.from_coo(rows2, cols2, vals2, n_terms, n2)
    for kk, thr in ((100, 0.0), (1000, 0.0), (37, 1.5)):
        r_s, r_i, r_c = ops.sparse_search(full2, q_off, q_t, q_w, kk, thr)
        lo2, hi2 = shard.ShardPlan(n2, world).bounds(rank)
        keep2 = (rows2 >= lo2) & (rows2 < hi2)
        part2 = ops.SparseDeviceIndex.from_coo((rows2[keep2] - lo2).contiguous(), cols2[keep2].contiguous(), vals2[keep2].contiguous(),
                                               n_terms, hi2 - lo2)
        ex = shard.TauExchange("sparse", n2, dev)
        assert ex.n_exchanges == 2, ex.n_exchanges
        for _ in range(2):           # the exchange object is reused across searches
            s, i, c = ops.sparse_search(part2, q_off, q_t, q_w, kk, thr, doc_id_base=lo2, exchange=ex)
            plain = ops.sparse_search(part2, q_off, q_t, q_w, kk, thr, doc_id_base=lo2)
            # with the exchanged bound a shard keeps only what can still reach the global top-k: a subset of its plain local rows
            assert bool((c <= plain[2]).all())
            s, i, c = shard.merge_shards(s, i, kk, n_docs_total=n2)
            assert torch.equal(i, r_i) and torch.equal(c, r_c) and torch.equal(s.view(torch.int32), r_s.view(torch.int32)), ("tau exchange", kk)
    del full2, part2, rows2, cols2, vals2
    check_overflow_tiers(rank, world, dev)

    # ---- sparse: class API (SparseRetrieval shards by itself under a process group) ------------------------------------
    index = IndexDictOfArray(index_path=None, dim_voc=n_terms, device=dev)
    index.add_batch_document(rows, cols, vals, n_docs=n_docs)
    retr = SparseRetrieval(torch.nn.Linear(1, 1), {"out_dir": "/tmp"}, n_terms, local,
                           index_d={"index": index, "ids_mapping": {d: f"D{d}" for d in range(n_docs)}})
    assert retr.doc_id_base == lo and retr.device_index.n_docs == hi - lo
    h = retr.search_arrays(q_off.cpu().numpy(), q_t.cpu().numpy(), q_w.cpu().numpy(), k, 0.0)
    assert np.array_equal(h[1], ref_i.cpu().numpy()) and np.array_equal(h[0].view(np.uint32), ref_s.cpu().numpy().view(np.uint32))
    for _ in range(2):     # host_ranks="first": every rank copies its merged query slice into host rows shared with rank 0
        h1 = retr.search_arrays(q_off.cpu().numpy(), q_t.cpu().numpy(), q_w.cpu().numpy(), k, 0.0, host_ranks="first")
        if rank == 0:
            assert np.array_equal(h1[1], ref_i.cpu().numpy()) and np.array_equal(h1[0].view(np.uint32), ref_s.cpu().numpy().view(np.uint32))
            assert np.array_equal(h1[2], ref_c.cpu().numpy())
        else:
            assert h1 == (None, None, None)
    res, _ = retr._sparse_retrieve_multithreaded(synth.queries_to_vecs(q_off, q_t, q_w), list(range(n_queries)), 0.0, k)
    if rank == 0:          # the run lives on the first worker (the rank that writes run.json); the others get an empty run
        assert res["3"][f"D{int(ref_i[3, 0])}"] == float(ref_s[3, 0]) and len(res) == int((ref_c > 0).sum())
    else:
        assert len(res) == 0

    # ---- dense ---------------------------------------------------------------------------------------------------------
    nd, dim, nq, kd = 30007, 256, 70, 200
    docs = synth.gen_dense(nd, dim, seed=7, device=dev, dtype=torch.bfloat16)
    queries = synth.gen_dense(nq, dim, seed=8, device=dev)
    q16 = ops.f32_to_bf16(queries)
    ref_s, ref_i, _ = ops.dense_search(docs, q16, kd)
    lo, hi = shard.ShardPlan(nd, world).bounds(rank)
    s, i, _ = ops.dense_search(docs[lo:hi].contiguous(), q16, kd, doc_id_base=lo)
    s, i, _ = shard.merge_shards(s, i, kd, n_docs_total=nd)
    assert torch.equal(i, ref_i) and torch.equal(s.view(torch.int32), ref_s.view(torch.int32)), "dense ops"
    # dense tau exchange, shards with different round counts (129 vs 128 tiles of 256 docs)
    nd2 = world * 32 * growth * 256 + 1        # 32g + 1 tiles of 256 docs on ranks 0..G-2 (3 rounds), 32g on the last one (2 rounds)
    docs2 = synth.gen_dense(nd2, dim, seed=17, device=dev, dtype=torch.bfloat16)
    r_s, r_i, _ = ops.dense_search(docs2, q16, kd)
    lo2, hi2 = shard.ShardPlan(nd2, world).bounds(rank)
    ex = shard.TauExchange("dense", nd2, dev)
    assert ex.n_exchanges == 2, ex.n_exchanges
    s, i, _ = ops.dense_search(docs2[lo2:hi2].contiguous(), q16, kd, doc_id_base=lo2, exchange=ex)
    s, i, _ = shard.merge_shards(s, i, kd, n_docs_total=nd2)
    assert torch.equal(i, r_i) and torch.equal(s.view(torch.int32), r_s.view(torch.int32)), "dense tau exchange"
    del docs2

    flat = DenseFlatIndexer(device=dev)
    flat.init_index(dim)
    flat.index_data(docs.float().cpu().numpy(), [f"P{j}" for j in range(nd)])
    assert flat.index.shape[0] == hi - lo
    top_ids, top_scores = flat.search_knn(queries.cpu().numpy(), kd)
    if rank == 0:          # like the sparse run: the merged rows are delivered to the first worker
        assert top_ids[5][0] == f"P{int(ref_i[5, 0])}" and np.array_equal(top_scores, ref_s.cpu().numpy())
    else:
        assert len(top_ids) == 0 and top_scores.shape == (0, kd)
    full_s, full_i = flat.search_arrays(queries.cpu().numpy(), kd)          # host_ranks="all": every rank gets every row
    assert np.array_equal(full_i, ref_i.cpu().numpy()) and np.array_equal(full_s, ref_s.cpu().numpy())
    d1 = flat.search_arrays(queries.cpu().numpy(), kd, host_ranks="first")
    if rank == 0:
        assert np.array_equal(d1[1], ref_i.cpu().numpy()) and np.array_equal(d1[0], ref_s.cpu().numpy())
    else:
        assert d1 == (None, None)

    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world_size={world}, sharded sparse + dense results identical to the single-GPU search")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
