"""Pins the CPU oracle (oracle/) to golden vectors produced by RUNNING THE REFERENCE'S OWN CODE
(tests/golden/make_golden.py): index build, numba_score_float, select_topk and the whole
_sparse_retrieve_multithreaded call.  CPU only; the GPU parity tests then compare the kernels with this oracle."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, sparse_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def feed(golden, case):
    row, col, val = golden[f"{case}_row"], golden[f"{case}_col"], golden[f"{case}_val"]
    if case == "B":   # merged two-rank index: rank 0's postings first, then rank 1's
        order = np.concatenate([np.nonzero(row % 2 == r)[0] for r in range(2)])
        row, col, val = row[order], col[order], val[order]
    return row, col, val


@pytest.mark.parametrize("case", ["A", "B"])
@pytest.mark.parametrize("impl", ["numpy", "c", "python-loop"])
def test_build_matches_reference(golden, case, impl):
    row, col, val = feed(golden, case)
    n_terms = int(golden[f"{case}_n_terms"])
    if impl == "numpy":
        off, ids, w = sparse_oracle.build_csr(row, col, val, n_terms)
    elif impl == "c":
        off, ids, w = c_oracle.build_csr(row, col, val, n_terms)
    else:
        idx = sparse_oracle.OracleIndex()
        idx.add_batch_document(row, col, val, n_docs=int(golden[f"{case}_n_docs"]))
        assert idx.nb_docs() == int(golden[f"{case}_n_docs"])
        d_ids, d_vals = idx.finalize()
        off = np.zeros(n_terms + 1, dtype=np.int64)
        for t in range(n_terms):
            off[t + 1] = off[t] + len(d_ids.get(t, ()))
        ids = np.concatenate([d_ids.get(t, np.zeros(0, np.int32)) for t in range(n_terms)])
        w = np.concatenate([d_vals.get(t, np.zeros(0, np.float32)) for t in range(n_terms)])
    assert np.array_equal(off, golden[f"{case}_offsets"])
    assert np.array_equal(ids, golden[f"{case}_ids"])
    assert np.array_equal(w.view(np.uint32), golden[f"{case}_vals"].view(np.uint32))


def queries(golden):
    q_off, q_t, q_w = golden["C_q_offsets"], golden["C_q_terms"], golden["C_q_weights"]
    return [(q_t[q_off[i]:q_off[i + 1]], q_w[q_off[i]:q_off[i + 1]]) for i in range(len(q_off) - 1)]


def test_score_float_matches_reference_bitwise(golden):
    off, ids, vals = golden["C_offsets"], golden["C_ids"], golden["C_vals"]
    n_docs, n_terms = int(golden["C_n_docs"]), int(golden["C_n_terms"])
    d_ids, d_vals = sparse_oracle.csr_to_dicts(off, ids, vals, n_terms)
    for ti, thr in enumerate(golden["C_thresholds"]):
        for qi, (t, w) in enumerate(queries(golden)):
            f, neg = sparse_oracle.score_float(d_ids, d_vals, t, w, float(thr), n_docs)
            assert f.dtype == np.int64 and neg.dtype == np.float32
            assert np.array_equal(f, golden[f"C_t{ti}_q{qi}_filtered"])
            assert np.array_equal(neg.view(np.uint32), golden[f"C_t{ti}_q{qi}_neg_scores"].view(np.uint32))
            full = c_oracle.sparse_scores(off, ids, vals, n_docs, t, w)
            assert np.array_equal(np.nonzero(full > thr)[0], f)
            assert np.array_equal((-full[f]).view(np.uint32), neg.view(np.uint32))


@pytest.mark.parametrize("k", [10, 100, 1000])
def test_select_topk_matches_reference(golden, k):
    """The reference's argpartition result is an unordered set with arbitrary boundary ties: compare sorted score
    multisets bit-for-bit, and ids strictly above the k-th score as sets."""
    off, ids, vals = golden["C_offsets"], golden["C_ids"], golden["C_vals"]
    n_docs = int(golden["C_n_docs"])
    q_off, q_t, q_w = golden["C_q_offsets"], golden["C_q_terms"], golden["C_q_weights"]
    for ti, thr in enumerate(golden["C_thresholds"]):
        c_scores, c_ids, c_counts = c_oracle.sparse_search(off, ids, vals, n_docs, q_off, q_t, q_w, k, float(thr))
        for qi in range(len(q_off) - 1):
            f, neg = golden[f"C_t{ti}_q{qi}_filtered"], golden[f"C_t{ti}_q{qi}_neg_scores"]
            sel_ids, sel_scores = sparse_oracle.select_topk(f, neg, k)
            ref_ids, ref_scores = golden[f"C_t{ti}_q{qi}_k{k}_ids"], golden[f"C_t{ti}_q{qi}_k{k}_scores"]
            assert len(sel_ids) == len(ref_ids) == int(c_counts[qi])
            assert np.array_equal(np.sort(sel_scores), np.sort(ref_scores))
            c = int(c_counts[qi])
            assert np.array_equal(c_scores[qi, :c].view(np.uint32), ref_scores.view(np.uint32))   # golden is sorted desc
            if c:
                kth = ref_scores[-1]
                assert set(c_ids[qi, :c][c_scores[qi, :c] > kth].tolist()) == set(ref_ids[ref_scores > kth].tolist())
                s_ids, s_scores = sparse_oracle.topk_sorted(f, neg, k)
                assert np.array_equal(s_ids, c_ids[qi, :c]) and np.array_equal(s_scores, c_scores[qi, :c])
            assert np.all(c_ids[qi, c:] == -1)


def test_retrieve_matches_reference_run(golden):
    with open(os.path.join(HERE, "golden", "retrieve_golden.json")) as f:
        run = json.load(f)
    off, ids, vals = golden["C_offsets"], golden["C_ids"], golden["C_vals"]
    n_docs, n_terms = int(golden["C_n_docs"]), int(golden["C_n_terms"])
    d_ids, d_vals = sparse_oracle.csr_to_dicts(off, ids, vals, n_terms)
    mapping = {i: f"D{7 * i}" for i in range(n_docs)}
    res, stats = sparse_oracle.retrieve(d_ids, d_vals, mapping, queries(golden), run["qids"], n_docs, run["threshold"], run["topk"])
    assert set(res) == set(run["res"])
    assert abs(stats["L0_q"] - run["stats"]["L0_q"]) < 1e-9
    for qid, ref_docs in run["res"].items():
        assert sorted(res[qid].values()) == sorted(ref_docs.values())
        kth = min(ref_docs.values())
        assert {d for d, s in res[qid].items() if s > kth} == {d for d, s in ref_docs.items() if s > kth}


def test_dense_oracle_is_exact_topk():
    from oracle import dense_oracle
    rng = np.random.default_rng(0)
    docs = rng.standard_normal((1000, 32)).astype(np.float32)
    qs = rng.standard_normal((7, 32)).astype(np.float32)
    scores, labels = dense_oracle.flat_ip_search(docs, qs, 20, block=128)
    full = qs @ docs.T
    for i in range(7):
        order = np.argsort(-full[i], kind="stable")[:20]
        assert np.array_equal(labels[i], order)
        np.testing.assert_allclose(scores[i], full[i][order], rtol=1e-6)
    scores, labels = dense_oracle.flat_ip_search(docs[:5], qs, 8)
    assert np.all(labels[:, 5:] == -1) and np.all(np.isneginf(scores[:, 5:]))
    assert np.all(np.diff(scores[:, :5], axis=1) <= 0)
    for i in range(7):
        assert np.array_equal(labels[i, :5], np.argsort(-(qs[i] @ docs[:5].T), kind="stable"))


def test_dense_oracle_matches_exact_construction():
    """tests/golden/dense_golden.npz (make_golden_dense.py): inner products exactly representable and tie-free by construction,
    so any correct IndexFlatIP must reproduce the stored ids and scores bit for bit — the restatement does."""
    import os
    from oracle import dense_oracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dense_golden.npz"))
    scores, ids = dense_oracle.flat_ip_search(g["docs"].astype(np.float32), g["queries"], int(g["k"]))
    assert np.array_equal(ids, g["top_ids"]) and np.array_equal(scores.view(np.uint32), g["top_scores"].view(np.uint32))
