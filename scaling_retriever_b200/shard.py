"""Doc-range sharding of a retrieval index over the GPUs of one box, and the candidate merge.

The reference searches a single in-RAM index from one process (eval_sparse.py:114 and eval_dense.py:191 assert
world_size == 1).  Here GPU g of G owns the contiguous document rows [g*ceil(N/G), (g+1)*ceil(N/G)) and searches all
queries against its shard.  The per-shard top-k rows are then merged under the same total order (score desc, doc id asc)
the search kernels use — so the sharded result is identical to the 1-GPU result — in three steps (`merge_shards`):

  1. every rank packs its rows into 8-byte keys (score bits | ~global id) and ONE all-to-all hands rank r the rows of the
     query slice [r*ceil(Q/G), (r+1)*ceil(Q/G)) from every shard (each GPU receives Q*k*8 bytes in total, 1/G of what an
     all-gather of all rows lands on it);
  2. rank r merges only its query slice (merge_keys kernel: G*k candidates -> k per query);
  3. one all-gather of the merged slices (again Q*k*8 bytes per GPU) gives every rank the full result.

NCCL over NVLink on GPUs, gloo on the CPU test path.  `merge_shards_allgather` is the older exchange (fp32 scores + int64 ids
all-gathered, every rank merges every query), kept for global ids that do not fit the key's 32 bits.
"""
import os
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import ops


@dataclass(frozen=True)
class ShardPlan:
    n_docs: int
    world_size: int

    @property
    def per_shard(self):
        return (self.n_docs + self.world_size - 1) // self.world_size if self.world_size > 0 else 0

    def bounds(self, rank):
        lo = min(self.n_docs, rank * self.per_shard)
        hi = min(self.n_docs, lo + self.per_shard)
        return lo, hi

    def owner(self, doc):
        return min(doc // self.per_shard, self.world_size - 1) if self.per_shard else 0


def shard_sparse_csr(term_offsets, doc_ids, weights, lo, hi):
    """Restrict a doc-sorted (or any) CSR to doc rows [lo, hi) with LOCAL row ids (torch index plumbing, any device)."""
    keep = (doc_ids >= lo) & (doc_ids < hi)
    n_terms = term_offsets.numel() - 1
    counts = term_offsets[1:] - term_offsets[:-1]
    term_of = torch.repeat_interleave(torch.arange(n_terms, device=doc_ids.device), counts, output_size=doc_ids.numel())
    new_counts = torch.bincount(term_of[keep], minlength=n_terms)
    new_offsets = torch.zeros(n_terms + 1, dtype=torch.int64, device=doc_ids.device)
    new_offsets[1:] = torch.cumsum(new_counts, dim=0)
    return new_offsets, (doc_ids[keep] - lo).to(torch.int32), weights[keep].contiguous()


def gather_candidates(scores, ids, group=None):
    """All-gather per-shard top-k rows: [Q, k] on every rank -> [G, Q, k] on every rank (fixed size, padded rows)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return scores.unsqueeze(0).contiguous(), ids.unsqueeze(0).contiguous()
    q, k = scores.shape
    # concatenated-along-dim-0 output form: accepted by both NCCL and gloo
    all_scores = torch.empty((world * q, k), dtype=scores.dtype, device=scores.device)
    all_ids = torch.empty((world * q, k), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(all_scores, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
    return all_scores.view(world, q, k), all_ids.view(world, q, k)


def merge_shards_allgather(scores, ids, k, group=None):
    """Local top-k rows of this rank's shard -> global top-k rows on every rank: two all-gathers (fp32 scores, int64 ids) and
    a merge of every query on every rank.  64-bit ids; 12 bytes per candidate on the wire."""
    all_scores, all_ids = gather_candidates(scores, ids, group)
    if all_scores.shape[0] == 1:
        counts = (all_ids[0] >= 0).sum(dim=1).to(torch.int32)
        return all_scores[0], all_ids[0], counts
    return ops.merge_topk(all_scores, all_ids, k)


def query_slice(n_queries, world):
    """Queries per rank in the all-to-all exchange (the last slices are padded with empty rows)."""
    return (n_queries + world - 1) // world if world > 0 else n_queries


def exchange_keys(keys, n_queries, group=None):
    """Packed keys [world * qs, k] of this rank's shard (rows >= n_queries zero) -> [world, qs, k]: the rows of THIS rank's
    query slice from every shard (one all-to-all)."""
    world = dist.get_world_size(group)
    qs = keys.shape[0] // world
    recv = torch.empty_like(keys)
    dist.all_to_all_single(recv, keys, group=group)
    return recv.view(world, qs, keys.shape[1])


def gather_merged(merged, n_queries, group=None):
    """Merged keys [qs, k] of this rank's query slice -> [n_queries, k] on every rank (one all-gather)."""
    world = dist.get_world_size(group)
    full = torch.empty((world * merged.shape[0], merged.shape[1]), dtype=merged.dtype, device=merged.device)
    dist.all_gather_into_tensor(full, merged.contiguous(), group=group)
    return full[:n_queries]


KEY_ID_LIMIT = (1 << 32) - 1     # global doc ids a packed key can carry


def merge_shards(scores, ids, k, group=None, n_docs_total=None):
    """Local top-k rows of this rank's shard -> global top-k rows (every rank gets the full result): packed-key all-to-all,
    per-slice merge, all-gather of the merged slices (module docstring).  `n_docs_total` >= 2^32 - 1 (ids that do not fit a
    key) takes the all-gather exchange instead."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or (n_docs_total is not None and n_docs_total >= KEY_ID_LIMIT):
        return merge_shards_allgather(scores, ids, k, group)
    n_queries = scores.shape[0]
    qs = query_slice(n_queries, world)
    keys = torch.empty((world * qs, k), dtype=torch.int64, device=scores.device)
    keys[n_queries:].zero_()
    ops.pack_keys(scores.contiguous(), ids.contiguous(), out=keys[:n_queries])
    merged = ops.merge_keys(exchange_keys(keys, n_queries, group), k)
    return ops.unpack_keys(gather_merged(merged, n_queries, group).contiguous(), k)


class TauExchange:
    """The tau exchange between the rounds of a doc-range sharded search (include/b200ret.h (3b)).

    Alone, a shard can only raise its eligibility bound tau[q] to ITS k-th best score, so every shard emits and selects as many
    candidates per round as a whole corpus would — the per-round cost that kept 8 GPUs at 0.85 of linear.  With the exchange
    each shard publishes, after every round, a score that at least ceil(k/G) of its candidates reach; one MIN all-reduce of that [Q]
    vector (28 KB, NCCL over NVLink, enqueued on the search's stream by the hook below — no host synchronisation) gives a
    bound that at least k documents of the whole corpus reach, and every shard continues with it.  The result is unchanged
    (exactly the global top-k after the merge).  Because the bound follows the documents ALL shards have seen, every shard may
    take larger steps — the docs scored grow max(4, G + 1) times per round instead of 4 times — without more candidates per
    shard and round: 4 rounds instead of 5 on a 1/8 shard of 8.8 M docs.
    `kind`: "sparse" or "dense" (their round schedules differ); `n_docs_total` sizes the LARGEST shard, which fixes how many
    all-reduces every shard takes part in."""

    def __init__(self, kind, n_docs_total, device, group=None):
        from . import _lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.device = device
        lib = _lib.load()
        per_shard = ShardPlan(n_docs_total, self.world).per_shard
        rounds = lib.b200ret_sparse_exchange_rounds if kind == "sparse" else lib.b200ret_dense_exchange_rounds
        self.n_exchanges = int(rounds(int(per_shard), self.world))
        self.growth = int(lib.b200ret_exchange_growth(self.world))
        self.aux = None
        self._struct = None
        self._error = None
        self._hook = _lib.RoundExchange.HOOK(self._on_round)       # kept alive with the object

    def _on_round(self, _user):
        try:
            dist.all_reduce(self.aux, op=dist.ReduceOp.MIN, group=self.group)
            return 0
        except Exception as exc:       # surfaced by the caller: the C side returns an error code
            self._error = exc
            return 1

    def struct(self, n_queries, k):
        import ctypes

        from . import _lib
        if self.aux is None or self.aux.numel() != n_queries:
            self.aux = torch.empty(n_queries, dtype=torch.float32, device=self.device)
        aux_rank = (int(k) + self.world - 1) // self.world
        self._struct = _lib.RoundExchange(aux_rank, self.n_exchanges, self.growth, 0, self.aux.data_ptr(), self._hook, None)
        return ctypes.byref(self._struct)


class SharedHostRows:
    """Result rows [n_queries, k] in HOST memory shared by the ranks of one box (a /dev/shm mapping every rank pins with
    cudaHostRegister): after the per-slice merge each GPU copies its merged query slice over ITS OWN PCIe link straight into
    the first worker's view of the result, instead of gathering all rows on one GPU and pushing 84 MB (6,980 x 1000 rows)
    through a single link.  The file is unlinked as soon as every rank has mapped it."""
    _serial = 0

    def __init__(self, n_queries, k, group=None):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        self.n_queries, self.k, self.qs = n_queries, k, query_slice(n_queries, world)
        rows = world * self.qs
        sizes = [rows * k * 4, rows * k * 8, rows * 4]                      # scores fp32, ids int64, counts int32
        offsets = np.concatenate([[0], np.cumsum([(b + 255) // 256 * 256 for b in sizes])])
        name = [None]
        if rank == 0:
            SharedHostRows._serial += 1
            name[0] = f"/dev/shm/b200ret_{os.getpid()}_{SharedHostRows._serial}"
            with open(name[0], "wb") as f:
                f.truncate(int(offsets[-1]))
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._raw = torch.from_file(name[0], shared=True, size=int(offsets[-1]), dtype=torch.uint8)
        self._registered = False
        if torch.cuda.is_available():
            err = torch.cuda.cudart().cudaHostRegister(self._raw.data_ptr(), self._raw.numel(), 0)
            self._registered = int(err) == 0
        dist.barrier(group)
        if rank == 0:
            os.unlink(name[0])
        self.scores = self._raw[offsets[0]:offsets[0] + sizes[0]].view(torch.float32).view(rows, k)
        self.ids = self._raw[offsets[1]:offsets[1] + sizes[1]].view(torch.int64).view(rows, k)
        self.counts = self._raw[offsets[2]:offsets[2] + sizes[2]].view(torch.int32)

    def slice_views(self, rank):
        a, b = rank * self.qs, (rank + 1) * self.qs
        return self.scores[a:b], self.ids[a:b], self.counts[a:b]

    def result(self):
        q = self.n_queries
        return self.scores[:q].numpy(), self.ids[:q].numpy(), self.counts[:q].numpy()

    def __del__(self):
        try:
            if self._registered:
                torch.cuda.cudart().cudaHostUnregister(self._raw.data_ptr())
        except Exception:
            pass


def merge_shards_to_first_host(scores, ids, k, host, group=None):
    """Sharded search, result wanted on the FIRST worker's host only (the rank that writes run.json): packed-key all-to-all,
    per-slice merge, then every rank copies its merged slice into the shared host rows (`host`: SharedHostRows) over its own
    PCIe link.  Returns (scores, ids, counts) numpy views on the first worker, (None, None, None) elsewhere; bytes copied by
    this rank are returned as the 4th value."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_queries = scores.shape[0]
    qs = host.qs
    keys = torch.empty((world * qs, k), dtype=torch.int64, device=scores.device)
    keys[n_queries:].zero_()
    ops.pack_keys(scores.contiguous(), ids.contiguous(), out=keys[:n_queries])
    merged = ops.merge_keys(exchange_keys(keys, n_queries, group), k)
    m_scores, m_ids, m_counts = ops.unpack_keys(merged, k)
    h_scores, h_ids, h_counts = host.slice_views(rank)
    h_scores.copy_(m_scores, non_blocking=True)
    h_ids.copy_(m_ids, non_blocking=True)
    h_counts.copy_(m_counts, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    dist.barrier(group)                      # every slice has landed in the shared rows
    nbytes = m_scores.numel() * 4 + m_ids.numel() * 8 + m_counts.numel() * 4
    if rank == 0:
        return (*host.result(), nbytes)
    return None, None, None, nbytes
