"""CPU restatement (numpy) of the reference's sparse index build + scoring + top-k.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package
(scaling_retriever_b200 / scaling_retriever); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and there only as the checker or the timed CPU baseline.

Each function restates one reference function (paths relative to the reference checkout) and is pinned
against golden vectors produced by running the reference's own code (tests/golden/make_golden.py ->
tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
from collections import defaultdict

import numpy as np


class OracleIndex:
    """Dict-of-arrays inverted index: restates IndexDictOfArray (scaling_retriever/utils/inverted_index.py:15-81)
    for the in-memory path (index_path=None)."""

    def __init__(self, dim_voc=None):
        self.n = 0
        self.dim_voc = dim_voc
        self.index_doc_id = defaultdict(list)
        self.index_doc_value = defaultdict(list)

    def add_batch_document(self, row, col, data, n_docs=-1):
        # inverted_index.py:67-76 — n grows by n_docs (or the number of distinct rows), every posting is appended
        # to the list of its term in feed order.
        if n_docs < 0:
            self.n += len(set(np.asarray(row).tolist()))
        else:
            self.n += n_docs
        for doc_id, dim_id, value in zip(np.asarray(row).tolist(), np.asarray(col).tolist(), np.asarray(data).tolist()):
            self.index_doc_id[dim_id].append(doc_id)
            self.index_doc_value[dim_id].append(value)

    def nb_docs(self):
        return self.n

    def finalize(self):
        # indexer.py:298-304 / inverted_index.py:84-88 — lists become np.int32 / np.float32 arrays.
        ids = {k: np.array(v, dtype=np.int32) for k, v in self.index_doc_id.items()}
        vals = {k: np.array(v, dtype=np.float32) for k, v in self.index_doc_value.items()}
        return ids, vals


def build_csr(row, col, data, n_terms):
    """Vectorised equivalent of OracleIndex.add_batch_document + finalize, returned as CSR arrays.

    Appending postings to per-term lists in feed order == a stable sort of the feed by term id.
    Returns (term_offsets int64[n_terms+1], doc_ids int32[nnz], weights float32[nnz]).
    """
    row = np.asarray(row)
    col = np.asarray(col)
    data = np.asarray(data, dtype=np.float32)
    order = np.argsort(col, kind="stable")
    counts = np.bincount(col.astype(np.int64), minlength=n_terms)
    term_offsets = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(counts, out=term_offsets[1:])
    return term_offsets, row[order].astype(np.int32), data[order]


def csr_to_dicts(term_offsets, doc_ids, weights, dim_voc):
    """CSR -> the two dicts SparseRetrieval.__init__ feeds to numba (indexer.py:356-370): every term id in
    range(dim_voc) is present, missing ones as empty arrays."""
    ids, vals = {}, {}
    for t in range(dim_voc):
        a, b = int(term_offsets[t]), int(term_offsets[t + 1])
        ids[t] = doc_ids[a:b]
        vals[t] = weights[a:b]
    return ids, vals


def score_float(index_ids, index_vals, indexes_to_retrieve, query_values, threshold, size_collection):
    """Restates SparseRetrieval.numba_score_float (scaling_retriever/indexer.py:324-344).

    scores[doc] += q * w as fp32 multiply then fp32 add, terms in the given order; doc ids are unique inside one
    posting list (an invariant of the build), so the fancy-index += touches each doc once per term.
    Returns (filtered_indexes int64[h], -scores[filtered] float32[h]).
    """
    scores = np.zeros(size_collection, dtype=np.float32)
    for local_idx, query_float in zip(indexes_to_retrieve, query_values):
        ids = index_ids[int(local_idx)]
        if len(ids) == 0:
            continue
        prod = np.float32(query_float) * index_vals[int(local_idx)].astype(np.float32)   # fp32 multiply
        scores[ids] = scores[ids] + prod                                                # fp32 add
    filtered_indexes = np.argwhere(scores > np.float32(threshold))[:, 0]
    return filtered_indexes, -scores[filtered_indexes]


def select_topk(filtered_indexes, scores, k):
    """Restates SparseRetrieval.select_topk (indexer.py:315-322): unordered top-k of the negated scores."""
    if len(filtered_indexes) > k:
        sorted_ = np.argpartition(scores, k)[:k]
        filtered_indexes, scores = filtered_indexes[sorted_], -scores[sorted_]
    else:
        scores = -scores
    return filtered_indexes, scores


def retrieve(index_ids, index_vals, doc_ids_map, sparse_query_vecs, qids, size_collection, threshold=0.0, topk=1000):
    """Restates SparseRetrieval._sparse_retrieve_multithreaded (indexer.py:405-474) without the thread pool:
    res[str(qid)][str(doc_ids[row])] = float(score); queries with no eligible doc get no key."""
    res = defaultdict(dict)
    stats = defaultdict(float)
    for qid, (col, values) in zip(qids, sparse_query_vecs):
        filtered, neg = score_float(index_ids, index_vals, col, values, threshold, size_collection)
        filtered, sc = select_topk(filtered, neg, topk)
        for id_, s in zip(filtered, sc):
            res[str(qid)][str(doc_ids_map[int(id_)])] = float(s)
        stats["L0_q"] += len(values) / len(qids)
    return res, stats


def topk_sorted(filtered_indexes, neg_scores, k):
    """Total-order top-k used to compare with the GPU rows: (score desc, doc id asc). Not in the reference
    (argpartition leaves ties at the k boundary arbitrary); the GPU resolves them to the lowest doc id."""
    scores = -neg_scores
    order = np.lexsort((filtered_indexes, -scores.astype(np.float64)))[:k]
    return filtered_indexes[order], scores[order]
