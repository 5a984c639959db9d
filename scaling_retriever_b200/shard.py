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
from dataclasses import dataclass

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
