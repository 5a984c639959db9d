"""CPU fp32 restatement of faiss.IndexFlatIP.search as the reference calls it. TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: faiss-cpu 1.8.0 (README.md:12 of the reference; conda `pytorch` channel) is a third-party
dependency that is neither vendored in the reference nor installable here, and the reference has no tests or
golden vectors at this boundary.  This file restates the published semantics of IndexFlatIP at the reference's
call sites — DenseFlatIndexer.init_index / index_data / search_knn (scaling_retriever/indexer.py:195-214):
  * index.add(x): keeps a float32 row-major copy of the corpus;
  * index.search(q, k): exact fp32 inner products, the k largest per query sorted DESCENDING with int64 row
    labels; when fewer than k rows exist the tail is padded with label -1 (scores -inf here; faiss pads with the
    lowest float);
  * search_knn then maps labels through index_id_to_db_id (indexer.py:212).
Tie order inside faiss is unspecified; this restatement orders ties by ascending row id like the GPU kernels.
"""
import numpy as np
import torch


def flat_ip_search(corpus, queries, k, block=65536):
    """corpus f32 [N, d], queries f32 [Q, d] -> (scores f32 [Q, k] desc, labels int64 [Q, k])."""
    corpus = torch.as_tensor(np.ascontiguousarray(corpus, dtype=np.float32))
    queries = torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32))
    n, q = corpus.shape[0], queries.shape[0]
    best_s = torch.full((q, 0), -float("inf"))
    best_i = torch.full((q, 0), -1, dtype=torch.int64)
    for lo in range(0, n, block):
        s = queries @ corpus[lo:lo + block].T          # fp32 sgemm block, like faiss's blocked search
        i = torch.arange(lo, lo + s.shape[1], dtype=torch.int64).expand(q, -1)
        best_s = torch.cat([best_s, s], dim=1)
        best_i = torch.cat([best_i, i], dim=1)
        if best_s.shape[1] > k:
            # total order (score desc, id asc): sort ids first (already ascending), stable sort by score
            order = torch.sort(best_s, dim=1, descending=True, stable=True).indices[:, :k]
            best_s = torch.gather(best_s, 1, order)
            best_i = torch.gather(best_i, 1, order)
            # re-sort kept block by id within equal scores is preserved by stability on the next round
    if best_s.shape[1]:
        # kept rows are (score desc, id asc) and a new block's ids are larger, so one stable sort restores the total order
        order = torch.sort(best_s, dim=1, descending=True, stable=True).indices[:, :k]
        best_s = torch.gather(best_s, 1, order)
        best_i = torch.gather(best_i, 1, order)
    if best_s.shape[1] < k:   # fewer than k rows in the index: faiss pads with label -1
        pad = k - best_s.shape[1]
        best_s = torch.cat([best_s, torch.full((q, pad), -float("inf"))], dim=1)
        best_i = torch.cat([best_i, torch.full((q, pad), -1, dtype=torch.int64)], dim=1)
    return best_s.numpy(), best_i.numpy()


def search_knn(corpus, doc_ids, queries, top_docs):
    """Restates DenseFlatIndexer.search_knn (indexer.py:210-214): labels -> external ids."""
    scores, indexes = flat_ip_search(corpus, queries, top_docs)
    top_doc_ids = [[doc_ids[idx] for idx in per_query] for per_query in indexes]
    return top_doc_ids, scores
