"""CPU restatement of the reference's mrr_k / recall_k (scaling_retriever/utils/metrics.py:13-42) with trec_eval's definitions
of recip_rank and recall_<k> written out (pytrec_eval is not installable here: PARITY UNPINNED against pytrec_eval itself;
the definitions are trec_eval's documented ones).  TEST INFRASTRUCTURE ONLY."""


def truncate_run(run, k):
    """utils/metrics.py:13-19."""
    temp_d = {}
    for q_id in run:
        sorted_run = {kk: v for kk, v in sorted(run[q_id].items(), key=lambda item: item[1], reverse=True)}
        temp_d[q_id] = {kk: sorted_run[kk] for kk in list(sorted_run.keys())[:k]}
    return temp_d


def _ranking(docs):
    # trec_eval: score descending, ties by docno descending
    return [d for d, _ in sorted(docs.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)]


def recip_rank(run, qrel):
    out = {}
    for q, docs in run.items():
        if q not in qrel:
            continue
        rr = 0.0
        for rank, d in enumerate(_ranking(docs), 1):
            if qrel[q].get(d, 0) > 0:
                rr = 1.0 / rank
                break
        out[q] = rr
    return out


def recall_at(run, qrel, k):
    out = {}
    for q, docs in run.items():
        if q not in qrel:
            continue
        n_rel = sum(1 for r in qrel[q].values() if r > 0)
        hit = sum(1 for d in _ranking(docs)[:k] if qrel[q].get(d, 0) > 0)
        out[q] = hit / n_rel if n_rel else 0.0
    return out


def mrr_k(run, qrel, k):
    vals = recip_rank(truncate_run(run, k), qrel)
    return sum(vals.values()) / max(1, len(vals))


def recall_k(run, qrel, k):
    vals = recall_at(run, qrel, k)
    return sum(vals.values()) / len(vals)
