"""Generate the golden vectors in this directory by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (the reference lives at /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's hot path imports three modules that are not installed here (ujson, faiss, h5py); none of them
is touched by the functions exercised below, so they are stubbed with empty modules before the import.
Functions executed (unmodified, from /root/reference):
  * scaling_retriever.utils.inverted_index.IndexDictOfArray.add_batch_document   (inverted_index.py:67-76)
  * scaling_retriever.indexer.SparseRetrieval.numba_score_float                  (indexer.py:324-344)
  * scaling_retriever.indexer.SparseRetrieval.select_topk                        (indexer.py:315-322)
  * scaling_retriever.indexer.SparseRetrieval.__init__ (index_d path) + _sparse_retrieve_multithreaded (:346-474)
Outputs: sparse_golden.npz (inputs + reference outputs), retrieve_golden.json.
"""
import json
import os
import sys
import tempfile
import types

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    for name in ("ujson", "faiss", "h5py"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["ujson"].dump = json.dump
    sys.modules["ujson"].load = json.load
    # the reference must win over any same-named package on the path
    sys.path[:] = [REFERENCE] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
    import scaling_retriever.indexer as ref_indexer
    import scaling_retriever.utils.inverted_index as ref_inv
    assert ref_indexer.__file__.startswith(REFERENCE), ref_indexer.__file__
    return ref_indexer, ref_inv


def random_docs(rng, n_docs, n_terms, mean_nnz, quantize=None):
    """Row-major COO like torch.nonzero of an encoder output: rows ascending, cols ascending inside a row."""
    import numpy as np
    rows, cols, vals = [], [], []
    # Zipf-ish popularity so some lists are long and many are empty
    p = 1.0 / (np.arange(n_terms) + 5.0)
    p /= p.sum()
    perm = rng.permutation(n_terms)
    for d in range(n_docs):
        nnz = max(1, rng.poisson(mean_nnz))
        t = np.unique(perm[rng.choice(n_terms, size=nnz, p=p)])
        v = np.log1p(rng.exponential(1.0, size=len(t))).astype(np.float32) + np.float32(1e-3)
        if quantize:
            v = (np.ceil(v / quantize) * quantize).astype(np.float32)
        rows.append(np.full(len(t), d, dtype=np.int64))
        cols.append(t.astype(np.int64))
        vals.append(v)
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def to_csr(index, n_terms):
    """Concatenate the reference's per-term arrays (converted exactly like indexer.py:300-304) in term order."""
    import numpy as np
    offs = np.zeros(n_terms + 1, dtype=np.int64)
    ids, vals = [], []
    for t in range(n_terms):
        if t in index.index_doc_id:
            a = np.array(index.index_doc_id[t], dtype=np.int32)
            v = np.array(index.index_doc_value[t], dtype=np.float32)
        else:
            a = np.array([], dtype=np.int32)
            v = np.array([], dtype=np.float32)
        ids.append(a)
        vals.append(v)
        offs[t + 1] = offs[t] + len(a)
    return offs, np.concatenate(ids), np.concatenate(vals)


def main():
    import numpy as np
    import numba
    import torch
    ref_indexer, ref_inv = import_reference()
    out = {}
    rng = np.random.default_rng(20261017)

    # ---- A. index build, single rank, three batches (add_batch_document) --------------------------------
    V_A, N_A = 97, 400
    row, col, val = random_docs(rng, N_A, V_A, 12)
    index = ref_inv.IndexDictOfArray(index_path=None)
    bounds = [0, 150, 151, N_A]          # uneven batches, one of a single document
    for b0, b1 in zip(bounds[:-1], bounds[1:]):
        m = (row >= b0) & (row < b1)
        g_row = row[m] * 1 + 0            # indexer.py:262 with world_size=1, rank 0
        index.add_batch_document(g_row, col[m], val[m], n_docs=b1 - b0)
    offs, ids, vals = to_csr(index, V_A)
    out.update(A_row=row.astype(np.int32), A_col=col.astype(np.int32), A_val=val, A_n_terms=V_A, A_n_docs=index.nb_docs(),
               A_offsets=offs, A_ids=ids, A_vals=vals)

    # ---- B. two ranks, rank-interleaved rows (indexer.py:262), merged like merge_indexes (:145-146) ------
    V_B, N_B = 61, 300
    row, col, val = random_docs(rng, N_B, V_B, 9)
    parts = []
    for rank in range(2):
        idx = ref_inv.IndexDictOfArray(index_path=None)
        m = (row % 2) == rank            # DistributedSampler(shuffle=False): rank r sees docs r, r+2, ...
        local = row[m] // 2
        g_row = local * 2 + rank
        idx.add_batch_document(g_row, col[m], val[m], n_docs=int(m.sum()))
        parts.append(idx)
    merged_ids, merged_vals = {}, {}
    for idx in parts:                    # np.append per term in directory order (rank 0 then rank 1)
        for t in range(V_B):
            a = np.array(idx.index_doc_id[t], dtype=np.int32) if t in idx.index_doc_id else np.array([], dtype=np.int32)
            v = np.array(idx.index_doc_value[t], dtype=np.float32) if t in idx.index_doc_value else np.array([], dtype=np.float32)
            if t not in merged_ids:
                merged_ids[t], merged_vals[t] = a, v
            else:
                merged_ids[t], merged_vals[t] = np.append(merged_ids[t], a), np.append(merged_vals[t], v)
    offs = np.zeros(V_B + 1, dtype=np.int64)
    for t in range(V_B):
        offs[t + 1] = offs[t] + len(merged_ids[t])
    out.update(B_row=row.astype(np.int32), B_col=col.astype(np.int32), B_val=val, B_n_terms=V_B, B_n_docs=N_B,
               B_offsets=offs, B_ids=np.concatenate([merged_ids[t] for t in range(V_B)]),
               B_vals=np.concatenate([merged_vals[t] for t in range(V_B)]))

    # ---- C. scoring + select on a 7000-doc index (spans 3 GPU doc blocks), ties included ------------------
    V_C, N_C = 211, 7000
    row, col, val = random_docs(rng, N_C, V_C - 5, 14, quantize=0.25)   # quantised weights -> exact ties; last 5 lists stay empty
    index = ref_inv.IndexDictOfArray(index_path=None)
    index.add_batch_document(row, col, val, n_docs=N_C)
    offs, ids, vals = to_csr(index, V_C)
    out.update(C_n_terms=V_C, C_n_docs=N_C, C_offsets=offs, C_ids=ids, C_vals=vals)

    ids_dict = numba.typed.Dict()
    vals_dict = numba.typed.Dict()
    for t in range(V_C):
        ids_dict[t] = ids[offs[t]:offs[t + 1]]
        vals_dict[t] = vals[offs[t]:offs[t + 1]]
    empty_terms = [t for t in range(V_C) if offs[t] == offs[t + 1]]
    used_terms = [t for t in range(V_C) if offs[t + 1] > offs[t]]
    queries = []
    for _ in range(9):
        nnz = max(1, rng.poisson(8))
        t = np.sort(rng.choice(used_terms, size=min(nnz, len(used_terms)), replace=False)).astype(np.int32)
        w = (np.ceil(np.log1p(rng.exponential(1.0, size=len(t))) / 0.25) * 0.25 + 0.25).astype(np.float32)
        queries.append((t, w))
    queries.append((np.array([], dtype=np.int32), np.array([], dtype=np.float32)))                     # empty query
    if empty_terms:
        queries.append((np.array(empty_terms[:3], dtype=np.int32), np.ones(len(empty_terms[:3]), dtype=np.float32)))  # only empty lists
    heavy = int(np.argmax(np.diff(offs)))
    queries.append((np.array([heavy], dtype=np.int32), np.array([1.5], dtype=np.float32)))           # one long list
    q_off = np.zeros(len(queries) + 1, dtype=np.int32)
    for i, (t, _) in enumerate(queries):
        q_off[i + 1] = q_off[i] + len(t)
    out.update(C_q_offsets=q_off, C_q_terms=np.concatenate([t for t, _ in queries]).astype(np.int32),
               C_q_weights=np.concatenate([w for _, w in queries]).astype(np.float32))
    thresholds = [0.0, 1.0]
    ks = [10, 100, 1000]
    out["C_thresholds"] = np.array(thresholds, dtype=np.float32)
    out["C_ks"] = np.array(ks, dtype=np.int32)
    for ti, thr in enumerate(thresholds):
        for qi, (t, w) in enumerate(queries):
            filtered, neg = ref_indexer.SparseRetrieval.numba_score_float(ids_dict, vals_dict, t, w, threshold=thr,
                                                                          size_collection=N_C)
            assert filtered.dtype == np.int64 and neg.dtype == np.float32
            out[f"C_t{ti}_q{qi}_filtered"] = filtered
            out[f"C_t{ti}_q{qi}_neg_scores"] = neg
            for k in ks:
                sel_idx, sel_sc = ref_indexer.SparseRetrieval.select_topk(filtered, neg, k=k)
                order = np.lexsort((sel_idx, -sel_sc.astype(np.float64)))
                out[f"C_t{ti}_q{qi}_k{k}_ids"] = sel_idx[order]
                out[f"C_t{ti}_q{qi}_k{k}_scores"] = sel_sc[order].astype(np.float32)

    # ---- D. the whole retrieval call: SparseRetrieval(index_d=...)._sparse_retrieve_multithreaded ----------
    index_d = ref_inv.IndexDictOfArray(index_path=None)
    index_d.add_batch_document(row, col, val, n_docs=N_C)
    for key in list(index_d.index_doc_id.keys()):           # indexer.py:298-304
        index_d.index_doc_id[key] = np.array(index_d.index_doc_id[key], dtype=np.int32)
        index_d.index_doc_value[key] = np.array(index_d.index_doc_value[key], dtype=np.float32)
    ids_mapping = {i: f"D{i * 7}" for i in range(N_C)}       # external ids differ from row ids
    tmp = tempfile.mkdtemp(prefix="golden_out_")
    retr = ref_indexer.SparseRetrieval(model=torch.nn.Linear(1, 1), config={"out_dir": tmp}, dim_voc=V_C, device="cpu",
                                       index_d={"index": index_d, "ids_mapping": ids_mapping})
    qids = [100 + i for i in range(len(queries))]
    res, stats = retr._sparse_retrieve_multithreaded(queries, qids, threshold=0.0, topk=50)
    with open(os.path.join(HERE, "retrieve_golden.json"), "w") as f:
        json.dump({"topk": 50, "threshold": 0.0, "qids": qids, "res": res, "stats": dict(stats),
                   "id_rule": "external id of row i is 'D' + str(7*i)"}, f)

    np.savez_compressed(os.path.join(HERE, "sparse_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "sparse_golden.npz"), "and retrieve_golden.json;",
          "numba", numba.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
