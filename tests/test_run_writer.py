"""Result materialisation: LazyRun + the native run.json writer against the reference's own
construction — `res[str(qid)][str(doc_ids[id_])] = float(sc)` (indexer.py:429-430) + `json.dump(res)` (:537-538).
Host-only code of libb200ret.so: runs without a GPU."""
import json
import struct
from collections import defaultdict

import numpy as np
import pytest

from scaling_retriever_b200.results import ExternalIds, LazyRun


def reference_run(qids, ids, scores, counts, doc_ids):
    """The reference's insert loop, verbatim semantics (defaultdict(dict), later insert wins, no key without a hit)."""
    res = defaultdict(dict)
    for i, qid in enumerate(qids):
        c = ids.shape[1] if counts is None else int(counts[i])
        for id_, sc in zip(ids[i, :c], scores[i, :c]):
            res[str(qid)][str(doc_ids[id_])] = float(sc)
    return res


def check_bytes(tmp_path, qids, ids, scores, counts, doc_ids, expect_path):
    run = LazyRun(qids, ids, scores, counts, ExternalIds(doc_ids, len(doc_ids) if not isinstance(doc_ids, dict) else None))
    ref = reference_run(qids, ids, scores, counts, doc_ids)
    assert dict(run) == dict(ref)
    assert list(run) == list(ref)                       # key order = first query with a hit first
    path = tmp_path / "run.json"
    took = run.write_json(str(path))
    assert took == expect_path
    assert path.read_bytes() == json.dumps(ref).encode()
    assert json.loads(path.read_text()) == json.loads(json.dumps(ref))
    return run


SPECIAL = [0.0, -0.0, 1.0, -1.5, 0.1, 1e-4, 1e-5, 9.999e-5, 1e16, 9.999999e15, 1.2345678e16, 3.4028235e38, 1e-38, 1e-45,
           123456.0, 16777216.0, 0.30000001192092896, 2.5e-7, 1e22, 5e-324, float("inf"), float("-inf"), float("nan")]


def test_float_repr_matches_python():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, size=20000, dtype=np.uint64).astype(np.uint32)
    vals = np.concatenate([bits.view(np.float32), np.asarray(SPECIAL, dtype=np.float32),
                           rng.random(5000, dtype=np.float32) * 40, np.exp(rng.normal(size=5000) * 8).astype(np.float32)])
    q = len(vals) // 50
    vals = vals[:q * 50].reshape(q, 50)
    ids = np.tile(np.arange(50, dtype=np.int64), (q, 1))
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        run = LazyRun(range(q), ids, vals, None, ExternalIds(range(50)))
        assert run.write_json(os.path.join(d, "r.json")) == "native"
        got = open(os.path.join(d, "r.json")).read()
    want = json.dumps({str(i): {str(j): float(vals[i, j]) for j in range(50)} for i in range(q)})
    assert got == want


def test_string_ids_escaping_and_missing_keys(tmp_path):
    doc_ids = {0: "D0", 1: 'quo"te', 2: "back\\slash", 3: "tab\tnl\n", 4: "é中\U0001f600", 5: "ctl\x01\x7f", 6: "", 7: "plain-7"}
    ids = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [7, 0, 3, 2], [1, 2, 3, 4]], dtype=np.int64)
    scores = np.array([[3.5, 2.25, 1.1, 0.7], [9.0, 8.5, 1e-7, 3.0], [1, 2, 3, 4], [4, 3, 2, 1]], dtype=np.float32)
    counts = np.array([4, 3, 0, 1], dtype=np.int32)            # query 2 has no hit -> no key
    run = check_bytes(tmp_path, ["qü1", 17, "z", 'k"'], ids, scores, counts, doc_ids, "native")
    assert "z" not in run and len(run) == 3
    with pytest.raises(KeyError):
        run["z"]


def test_int_ids_identity_and_negative_labels(tmp_path):
    ids = np.array([[5, 2, -1], [0, 1, 2]], dtype=np.int64)
    scores = np.array([[0.5, 0.25, float("-inf")], [1.5, 1.25, 1.0]], dtype=np.float32)
    db_ids = [100, 101, 102, 103, 104, 105, 999]                # DenseFlatIndexer.index_id_to_db_id (ints)
    check_bytes(tmp_path, ["a", "b"], ids, scores, None, db_ids, "native")   # -1 -> last id (indexer.py:212)
    check_bytes(tmp_path, ["a", "b"], ids, scores, None, range(7), "native")
    check_bytes(tmp_path, ["a", "b"], ids, scores, None, [str(x) for x in db_ids], "native")


def test_duplicates_take_the_python_path_with_reference_semantics(tmp_path):
    ids = np.array([[0, 1, 2], [2, 1, 0]], dtype=np.int64)
    scores = np.array([[3, 2, 1], [6, 5, 4]], dtype=np.float32)
    check_bytes(tmp_path, ["q", "q"], ids, scores, None, ["a", "b", "c"], "python")          # same qid twice: rows merge
    check_bytes(tmp_path, ["q", "r"], ids, scores, None, ["a", "b", "a"], "python")          # duplicate external ids collapse
    check_bytes(tmp_path, ["q", "r"], ids, scores, None, ["a", "b\x00", "c"], "python")      # NUL inside an id


def test_empty_run(tmp_path):
    run = LazyRun([], np.zeros((0, 10), np.int64), np.zeros((0, 10), np.float32), np.zeros(0, np.int32), ExternalIds(range(4)))
    run.write_json(str(tmp_path / "e.json"))
    assert (tmp_path / "e.json").read_text() == "{}"
    run = LazyRun(["a"], np.zeros((1, 10), np.int64), np.zeros((1, 10), np.float32), np.zeros(1, np.int32), ExternalIds(range(4)))
    run.write_json(str(tmp_path / "e.json"))
    assert (tmp_path / "e.json").read_text() == "{}"


def test_large_run_is_fast_and_identical(tmp_path):
    import time
    rng = np.random.default_rng(1)
    q, k, n = 500, 1000, 200000
    ids = np.stack([rng.choice(n, size=k, replace=False) for _ in range(q)]).astype(np.int64)
    scores = np.sort(rng.random((q, k), dtype=np.float32) * 30, axis=1)[:, ::-1].copy()
    doc_ids = {i: str(7 * i) for i in range(n)}
    ext = ExternalIds(doc_ids, n)
    ext.native()
    run = LazyRun(range(1000, 1000 + q), ids, scores, None, ext)
    t0 = time.perf_counter()
    run.write_json(str(tmp_path / "big.json"))
    native_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = json.dumps(reference_run(range(1000, 1000 + q), ids, scores, None, doc_ids))
    python_s = time.perf_counter() - t0
    assert (tmp_path / "big.json").read_text() == want
    assert native_s < python_s
