"""Doc-range sharding of a retrieval index over the GPUs of one box, and the candidate merge.

The reference searches a single in-RAM index from one process (eval_sparse.py:114 and eval_dense.py:191 assert
world_size == 1).  Here GPU g of G owns the contiguous document rows [g*ceil(N/G), (g+1)*ceil(N/G)), searches all
queries against its shard, and the per-shard top-k rows (k x (score, global id)) are exchanged with ONE all-gather
(NCCL over NVLink on GPUs, gloo on the CPU test path) and merged by the merge_topk kernel under the same total
order (score desc, doc id asc) the search kernels use — so the sharded result is identical to the 1-GPU result.
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


def merge_shards(scores, ids, k, group=None):
    """Local top-k rows of this rank's shard -> global top-k rows (every rank gets the full result)."""
    all_scores, all_ids = gather_candidates(scores, ids, group)
    if all_scores.shape[0] == 1:
        counts = (all_ids[0] >= 0).sum(dim=1).to(torch.int32)
        return all_scores[0], all_ids[0], counts
    return ops.merge_topk(all_scores, all_ids, k)
