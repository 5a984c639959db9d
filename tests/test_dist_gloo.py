"""world_size-2 test of the shard-merge plumbing on CPU (gloo): each rank searches its doc-range shard with the
oracle (standing in for the GPU kernel, which cannot run here), the candidates are all-gathered with
shard.gather_candidates, and the merged result must equal the unsharded search.  The merge itself is checked against
the same total order the merge_topk kernel implements (its GPU parity test is tests/test_merge_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_rows_reference(all_scores, all_ids, k):
    """numpy restatement of merge_topk's contract: [G, Q, k] -> [Q, k] under (score desc, id asc), -1 ids dropped."""
    g, q, _ = all_scores.shape
    out_s = np.full((q, k), -np.inf, dtype=np.float32)
    out_i = np.full((q, k), -1, dtype=np.int64)
    counts = np.zeros(q, dtype=np.int32)
    for qi in range(q):
        s = all_scores[:, qi, :].reshape(-1)
        i = all_ids[:, qi, :].reshape(-1)
        live = i >= 0
        s, i = s[live], i[live]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        out_s[qi, :len(order)], out_i[qi, :len(order)], counts[qi] = s[order], i[order], len(order)
    return out_s, out_i, counts


def pack_keys_reference(scores, ids):
    """numpy restatement of b200ret_pack_keys: (order-preserving fp32 bits << 32) | ~uint32(id); id -1 -> 0."""
    u = np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32)
    hi = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint64)
    lo = (~ids.astype(np.uint32)).astype(np.uint64)
    return np.where(ids >= 0, (hi << np.uint64(32)) | lo, np.uint64(0))


def unpack_keys_reference(keys):
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7fffffff), ~hi)
    live = keys != 0
    scores = np.where(live, u.view(np.float32), np.float32(-np.inf)).astype(np.float32)
    ids = np.where(live, (~keys.astype(np.uint32)).astype(np.int64), -1)
    return scores, ids, live.sum(axis=1).astype(np.int32)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from scaling_retriever_b200 import shard, synth
    n_docs, n_terms, k = 5000, 300, 25
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=12, seed=21)
    off, ids, w = c_oracle.build_csr(rows.numpy(), cols.numpy(), vals.numpy(), n_terms)
    q_off, q_t, q_w = (x.numpy() for x in synth.gen_sparse_queries(17, n_terms=n_terms, mean_nnz=6, seed=22))
    lo, hi = shard.ShardPlan(n_docs, world).bounds(rank)
    s_off, s_ids, s_w = shard.shard_sparse_csr(torch.as_tensor(off), torch.as_tensor(ids), torch.as_tensor(w), lo, hi)
    scores, lids, _ = c_oracle.sparse_search(s_off.numpy(), s_ids.numpy(), s_w.numpy(), hi - lo, q_off, q_t, q_w, k)
    gids = np.where(lids >= 0, lids + lo, -1)
    all_scores, all_ids = shard.gather_candidates(torch.as_tensor(scores), torch.as_tensor(gids))
    assert all_scores.shape == (world, 17, k)
    merged = merge_rows_reference(all_scores.numpy(), all_ids.numpy(), k)
    full = c_oracle.sparse_search(off, ids, w, n_docs, q_off, q_t, q_w, k)
    ok = all(np.array_equal(a, b) for a, b in zip(merged, full))

    # the packed-key exchange (shard.merge_shards): all-to-all of 8-byte keys, per-slice merge, all-gather of the slices.
    # The GPU kernels (pack / merge_keys / unpack) are restated in numpy here; the collectives are the product's.
    nq = len(q_off) - 1
    qs = shard.query_slice(nq, world)
    keys = np.zeros((world * qs, k), dtype=np.uint64)
    keys[:nq] = pack_keys_reference(scores, gids)
    recv = shard.exchange_keys(torch.as_tensor(keys.view(np.int64)), nq)
    assert recv.shape == (world, qs, k)
    mine = np.sort(recv.numpy().view(np.uint64).transpose(1, 0, 2).reshape(qs, world * k), axis=1)[:, ::-1][:, :k]
    got = shard.gather_merged(torch.as_tensor(np.ascontiguousarray(mine).view(np.int64)), nq).numpy().view(np.uint64)
    u_scores, u_ids, u_counts = unpack_keys_reference(got)
    ok = ok and np.array_equal(u_ids, full[1]) and np.array_equal(u_scores.view(np.uint32), full[0].view(np.uint32)) \
        and np.array_equal(u_counts, full[2])

    # tau exchange (shard.TauExchange): the hook the C library calls between rounds is one MIN all-reduce of the published bounds;
    # every rank takes part in the same number of exchanges (the schedule of the LARGEST shard)
    ex = shard.TauExchange("sparse", n_docs, torch.device("cpu"))
    ex.struct(nq, k)
    from scaling_retriever_b200 import _lib
    lib = _lib.load()
    ok = ok and ex.n_exchanges == lib.b200ret_sparse_exchange_rounds(shard.ShardPlan(n_docs, world).per_shard, world)
    ok = ok and ex.growth == max(4, world + 1) == ex._struct.growth
    ok = ok and ex._struct.aux_rank == (k + world - 1) // world and ex._struct.n_exchanges == ex.n_exchanges
    ex.aux.copy_(torch.arange(nq, dtype=torch.float32) + 10.0 * rank)
    ok = ok and ex._on_round(None) == 0 and bool(torch.equal(ex.aux, torch.arange(nq, dtype=torch.float32)))

    # shared host rows (shard.SharedHostRows): every rank writes its merged query slice, the first worker reads all of them
    host = shard.SharedHostRows(nq, k)
    a, b = rank * qs, min(nq, (rank + 1) * qs)
    hs, hi, hc = host.slice_views(rank)
    hs[:b - a].copy_(torch.as_tensor(u_scores[a:b]))
    hi[:b - a].copy_(torch.as_tensor(u_ids[a:b]))
    hc[:b - a].copy_(torch.as_tensor(u_counts[a:b]))
    dist.barrier()
    if rank == 0:
        r_s, r_i, r_c = host.result()
        ok = ok and np.array_equal(r_i, full[1]) and np.array_equal(r_s.view(np.uint32), full[0].view(np.uint32)) \
            and np.array_equal(r_c, full[2])
        ok = ok and not any(n.startswith("b200ret_") for n in os.listdir("/dev/shm"))      # the mapping's file is already unlinked
    dist.barrier()
    np.save(os.path.join(tmp, f"ok_{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_merge_equals_unsharded(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert bool(np.load(tmp_path / f"ok_{r}.npy")[0])


def test_three_rank_exchange_with_ragged_query_slices(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)     # 17 queries over 3 ranks: slices 6, 6, 5 (+1 pad)
    for r in range(3):
        assert bool(np.load(tmp_path / f"ok_{r}.npy")[0])


def test_key_packing_roundtrip_and_order():
    rng = np.random.default_rng(3)
    s = np.concatenate([rng.normal(size=500).astype(np.float32) * 20, [0.0, -0.0, np.inf, -np.inf, 1e-38, -1e-38]]).astype(np.float32)
    i = rng.choice(2 ** 32 - 2, size=len(s), replace=False).astype(np.int64)
    keys = pack_keys_reference(s[None, :], i[None, :])
    back_s, back_i, c = unpack_keys_reference(keys)
    assert np.array_equal(back_i[0], i) and np.array_equal(back_s.view(np.uint32), s[None, :].view(np.uint32)) and c[0] == len(s)
    order = np.argsort(keys[0])[::-1]
    ref = np.lexsort((i, -s.astype(np.float64)))
    nz = s != 0           # +0.0 and -0.0 are distinct keys but equal floats
    assert np.array_equal(order[nz[order]], ref[nz[ref]])


def test_merge_reference_total_order():
    s = np.array([[[3., 1., -np.inf]], [[3., 2., 2.]]], dtype=np.float32)
    i = np.array([[[5, 9, -1]], [[4, 7, 6]]], dtype=np.int64)
    out_s, out_i, c = merge_rows_reference(s, i, 4)
    assert out_i.tolist() == [[4, 5, 6, 7]] and out_s.tolist() == [[3., 3., 2., 2.]] and c.tolist() == [4]


def test_exchange_round_counts_follow_the_geometric_schedule():
    """b200ret_*_exchange_rounds = select launches between the rounds of candidates.cuh's schedule (first round r0 units, then the
    docs seen grow 4x per round) — what every shard must agree on before a sharded search."""
    sys.path.insert(0, ROOT)
    from scaling_retriever_b200 import _lib, ops
    lib = _lib.load()
    bd = ops.block_docs()

    def rounds(n_units, r0, growth=4):
        unit, size, selects = 0, r0, 0
        while unit < n_units:
            end = n_units if n_units - unit <= size else unit + size
            selects += end < n_units
            unit, size = end, end * (growth - 1)
        return selects
    assert [lib.b200ret_exchange_growth(g) for g in (1, 2, 4, 8)] == [4, 4, 5, 9]
    for shards in (1, 2, 8):
        growth = lib.b200ret_exchange_growth(shards)
        for blocks in (0, 1, 2, 3, 8, 9, 32, 33, 309, 2468):
            assert lib.b200ret_sparse_exchange_rounds(blocks * bd, shards) == rounds(blocks, 2, growth), (blocks, shards)
        for tiles in (0, 1, 32, 33, 128, 129, 4317):
            assert lib.b200ret_dense_exchange_rounds(tiles * 256, shards) == rounds(tiles, 32, growth), (tiles, shards)
    assert lib.b200ret_sparse_exchange_rounds(309 * bd, 8) == 3          # 2, 18, 162, 309 blocks: four rounds on a 1/8 shard of 8.8 M docs (five without)
