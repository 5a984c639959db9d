"""MRR@k / recall@k of a search result computed on the GPU from its [Q, k] arrays.

Quick-regression counterpart of the reference's `mrr_k` / `recall_k` (scaling_retriever/utils/metrics.py:22-42), which
serialise the run to a dict, truncate / sort it in Python and call pytrec_eval (trec_eval's recip_rank / recall_<k>).
Here the rows never leave their arrays: relevance judgements are mapped to row labels once and one kernel
(`b200ret_rank_metrics`) produces the per-query values.  Same definitions:
  * only queries present in BOTH the run and the qrel are evaluated, and the aggregate is their mean;
  * a document is relevant when its judgement is > 0;
  * mrr_k looks at the k best rows only (truncate_run); recall_k divides by ALL relevant documents of the query.
Rows are ranked by (score desc, row id asc); trec_eval breaks score ties by descending docno instead — identical unless two
retrieved documents of a query have exactly equal scores around a relevant one.
"""
import numpy as np
import torch

from . import ops
from .results import LazyRun


def _pack_qrel(run, qrel):
    """(row indices of the evaluated queries, rel_offsets int64, rel_ids int64 ascending per query) for `run` (LazyRun)."""
    label_of = {}
    ext = run.ext
    kind, payload = ("identity", None) if isinstance(ext._src, range) else (None, None)
    if kind != "identity":
        for row, x in enumerate(ext.obj.tolist()):
            if x is not None:
                label_of[str(x)] = row
    rows, offsets, rel = [], [0], []
    for qid, row_list in run._rows.items():
        judged = qrel.get(qid)
        if judged is None:
            continue
        labels = []
        for docid, r in judged.items():
            if r > 0:
                if kind == "identity":
                    lab = int(docid) if str(docid).lstrip("-").isdigit() and 0 <= int(docid) < ext.size else None
                else:
                    lab = label_of.get(str(docid))
                labels.append(-2 - len(labels) if lab is None else lab)   # judged docs outside the collection still count in the denominator
        labels.sort()
        rows.append(row_list[-1])
        rel.extend(labels)
        offsets.append(len(rel))
    return np.asarray(rows, dtype=np.int64), np.asarray(offsets, dtype=np.int64), np.asarray(rel, dtype=np.int64)


def rank_metrics(run, qrel, mrr_cut=10, recall_cuts=(10, 100, 1000), device=None):
    """Per-query metrics of a LazyRun: {"qids": [...], "recip_rank": fp32 [n], "recall": fp32 [n, len(cuts)]} (numpy)."""
    if not isinstance(run, LazyRun):
        raise TypeError("rank_metrics works on the LazyRun a retriever returns (the arrays of the search), not on a dict")
    rows, offsets, rel = _pack_qrel(run, qrel)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if len(rows) == 0:
        return {"qids": [], "recip_rank": np.zeros(0, np.float32), "recall": np.zeros((0, len(recall_cuts)), np.float32)}
    ids = torch.from_numpy(run.ids[rows]).to(dev)
    counts = torch.from_numpy(run.counts[rows]).to(dev)
    rr, recall = ops.rank_metrics(ids, counts, torch.from_numpy(offsets).to(dev), torch.from_numpy(rel).to(dev), mrr_cut, recall_cuts)
    return {"qids": [run.qids[i] for i in rows.tolist()], "recip_rank": rr.cpu().numpy(), "recall": recall.cpu().numpy()}


def mrr_k(run, qrel, k, agg=True):
    """reference utils/metrics.py:22-30."""
    out = rank_metrics(run, qrel, mrr_cut=k, recall_cuts=())
    if agg:
        return float(out["recip_rank"].astype(np.float64).sum() / max(1, len(out["qids"])))
    return {q: {"recip_rank": float(v)} for q, v in zip(out["qids"], out["recip_rank"])}


def recall_k(run, qrel, k, agg=True):
    """reference utils/metrics.py:32-42 (always aggregated there)."""
    out = rank_metrics(run, qrel, mrr_cut=1, recall_cuts=(k,))
    return float(out["recall"][:, 0].astype(np.float64).sum() / len(out["qids"]))
