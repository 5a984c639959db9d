"""Result materialisation: the [Q, k] arrays of a search -> the reference's `{qid: {docid: score}}` run, lazily.

The reference builds the run with one CPython dict insert per (query, doc) pair under the GIL
(`res[str(qid)][str(doc_ids[id_])] = float(sc)`, indexer.py:429-430; eval_dense.py:229-241) and `json.dump`s it
(indexer.py:537-538): ~10 s for 6,980 x 1000 pairs around a 0.11 s search.  Here the search result stays in its arrays:

* `LazyRun` is a read-only Mapping with the run's dict-of-dicts semantics (keys = str(qid) of the queries that have at
  least one eligible doc, first-seen order; values = {str(external id): float(score)} built on access);
* `LazyRun.write_json(path)` formats run.json straight from the arrays with the host-side writer of libb200ret.so
  (`b200ret_write_run_json`, csrc/run_writer.cpp), byte-identical to `json.dumps(dict_of_dicts)`.
"""
import ctypes
import json
from collections.abc import Mapping, Sequence

import numpy as np

from . import _lib


def owned_copy(arr):
    """A caller-owned copy of a (pinned staging) result array, made by several host threads."""
    arr = np.ascontiguousarray(arr)
    out = np.empty_like(arr)
    if arr.nbytes < (4 << 20):
        out[...] = arr
    else:
        _lib.check(_lib.load().b200ret_host_copy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(arr.ctypes.data), arr.nbytes, 8))
    return out


class ExternalIds:
    """Row label -> external id table (`doc_ids.pkl` dict / list, `index_id_to_db_id` list, or a range), in the forms the
    lazy run and the native writer need.  Everything is built on first use and cached."""

    def __init__(self, doc_ids, size=None):
        self._src = doc_ids
        self.size = int(size) if size is not None else (max(doc_ids) + 1 if isinstance(doc_ids, dict) and doc_ids else len(doc_ids))
        self._obj = None
        self._native = None

    @property
    def obj(self):
        """numpy object array: obj[row] = external id (None for rows without an entry)."""
        if self._obj is None:
            src = self._src
            ext = np.empty(self.size, dtype=object)
            if isinstance(src, dict):
                if src:
                    rows = np.fromiter(src.keys(), dtype=np.int64, count=len(src))
                    vals = np.empty(len(src), dtype=object)
                    vals[:] = list(src.values())
                    ext[rows] = vals
            elif isinstance(src, range) and src == range(self.size):
                ext[:] = np.arange(self.size).tolist()
            else:
                ext[:len(src)] = list(src)
            self._obj = ext
        return self._obj

    def int_table(self):
        """int64[size] when every external id is a Python / numpy integer (MS MARCO pids saved as int64 .npy), else None."""
        if not hasattr(self, "_ints"):
            self._ints = None
            src = self._src
            if isinstance(src, range) and src == range(self.size):
                self._ints = "identity"
            elif not isinstance(src, dict) and self.size and all(type(x) is int for x in (src[0], src[-1], src[self.size // 2])):
                try:
                    ints = np.asarray(src, dtype=np.int64)
                    if ints.shape == (self.size,) and all(type(x) is int for x in src):
                        self._ints = ints
                except (TypeError, ValueError, OverflowError):
                    pass
        return self._ints

    def gather_lists(self, labels):
        """[[external id of label] ...] as nested Python lists (DenseFlatIndexer.search_knn's return, indexer.py:212; negative
        labels index from the end like the reference's list).  Integer tables are gathered as int64 (3x faster than through the
        object array)."""
        ints = self.int_table()
        if isinstance(ints, str):          # identity
            return np.where(labels < 0, labels + self.size, labels).tolist()
        if ints is not None:
            return ints[labels].tolist()
        return self.obj[labels].tolist()

    def native(self):
        """('identity', None) | ('ints', int64[size]) | ('strs', (blob uint8, offsets int64[size+1])) | ('none', why):
        the table as the C writer takes it; 'none' when the writer's preconditions do not hold (duplicate external ids
        would collapse inside a query's dict; a NUL inside an id)."""
        if self._native is None:
            self._native = self._build_native()
        return self._native

    def _build_native(self):
        src = self._src
        if isinstance(src, range) and src == range(self.size):
            return "identity", None
        obj = self.obj
        if self.size == 0:
            return "identity", None
        if all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in obj[:64]):
            try:
                ints = np.asarray(obj.tolist(), dtype=np.int64)
                if ints.shape == (self.size,) and all(isinstance(x, (int, np.integer)) for x in obj[-64:]):
                    if len(np.unique(ints)) != self.size:
                        return "none", "duplicate external ids"
                    return "ints", ints
            except (TypeError, ValueError, OverflowError):
                pass
        strs = [str(x) for x in obj.tolist()]
        if len(set(strs)) != len(strs):
            return "none", "duplicate external ids"
        blob = np.frombuffer(("\x00".join(strs) + "\x00").encode("utf-8"), dtype=np.uint8)
        ends = np.flatnonzero(blob == 0)
        if len(ends) != len(strs):
            return "none", "NUL byte inside an external id"
        offsets = np.zeros(len(strs) + 1, dtype=np.int64)
        offsets[1:] = ends + 1
        return "strs", (blob, offsets)


class IdRows(Sequence):
    """`[[external id ...] per query]` — DenseFlatIndexer.search_knn's first return value (indexer.py:212) — as a read-only
    sequence over the label array: row i is gathered into a Python list when it is indexed or iterated.  Building all
    Q * k Python objects eagerly costs ~1 s for 6,980 x 1000 (CPython object creation), 4x the search itself; the reference's
    callers only iterate the rows once (eval_dense.py:229-241).  Compares equal to the eager list of lists; `.tolist()`
    materialises it."""

    def __init__(self, ext, labels):
        self.ext = ext
        self.labels = labels            # int64 [Q, k], owned

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return self.ext.gather_lists(self.labels[i])

    def tolist(self):
        return self.ext.gather_lists(self.labels)

    def __eq__(self, other):
        if isinstance(other, IdRows):
            other = other.tolist()
        return self.tolist() == other

    def __repr__(self):
        return f"IdRows({len(self)} rows x {self.labels.shape[1] if self.labels.ndim == 2 else 0})"


def _str_blob(strs):
    blob = np.frombuffer(("\x00".join(strs) + "\x00").encode("utf-8"), dtype=np.uint8) if strs else np.zeros(0, np.uint8)
    ends = np.flatnonzero(blob == 0)
    if len(ends) != len(strs):
        return None, None
    offsets = np.zeros(len(strs) + 1, dtype=np.int64)
    offsets[1:] = ends + 1
    return blob, offsets


class LazyRun(Mapping):
    """`{str(qid): {str(external id): float(score)}}` over the result arrays of one search.

    ids int64 [Q, k] row labels, scores fp32 [Q, k], counts int32 [Q] (live entries per row; None = all k); the arrays are
    owned by the run (callers pass copies of reusable staging buffers).  Semantics of the reference's defaultdict(dict)
    after its insert loop: a query without a live entry has no key; the same qid twice merges both rows (later wins)."""

    def __init__(self, qids, ids, scores, counts, ext):
        self.qids = [str(q) for q in qids]
        self.ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.scores = np.ascontiguousarray(scores, dtype=np.float32)
        self.counts = (np.full(len(self.qids), self.ids.shape[1] if self.ids.ndim == 2 else 0, dtype=np.int32)
                       if counts is None else np.ascontiguousarray(counts, dtype=np.int32))
        assert self.ids.shape == self.scores.shape and len(self.qids) == self.ids.shape[0] == len(self.counts)
        self.ext = ext if isinstance(ext, ExternalIds) else ExternalIds(ext)
        self._rows = {}
        for i in np.flatnonzero(self.counts > 0).tolist():
            self._rows.setdefault(self.qids[i], []).append(i)

    # ---- Mapping ------------------------------------------------------------------------------------------------
    def __getitem__(self, qid):
        docs = {}
        ext = self.ext.obj
        for i in self._rows[qid]:
            c = int(self.counts[i])
            docs.update(zip(map(str, ext[self.ids[i, :c]].tolist()), self.scores[i, :c].astype(float).tolist()))
        return docs

    def __iter__(self):
        return iter(self._rows)

    def __len__(self):
        return len(self._rows)

    def __contains__(self, qid):
        return qid in self._rows

    def to_dict(self):
        """The eager dict of dicts the reference returns."""
        return {qid: self[qid] for qid in self._rows}

    # ---- run.json -----------------------------------------------------------------------------------------------
    def native_writer_eligible(self):
        if any(len(r) != 1 for r in self._rows.values()):
            return False, "a query id occurs more than once"
        kind, payload = self.ext.native()
        if kind == "none":
            return False, payload
        return True, kind

    def write_json(self, path):
        """Write run.json; returns 'native' or 'python' (the path taken).  Same bytes either way."""
        ok, _ = self.native_writer_eligible()
        if ok:
            q_blob, q_off = _str_blob(self.qids)
            ok = q_blob is not None
        if not ok:
            with open(path, "w") as handler:
                handler.write(json.dumps(self.to_dict()))
            return "python"
        kind, payload = self.ext.native()
        P = ctypes.c_void_p
        blob_p = off_p = ints_p = None
        if kind == "strs":
            blob_p, off_p = P(payload[0].ctypes.data), P(payload[1].ctypes.data)
        elif kind == "ints":
            ints_p = P(payload.ctypes.data)
        written = ctypes.c_int64(0)
        k = self.ids.shape[1] if self.ids.ndim == 2 else 0
        _lib.check(_lib.load().b200ret_write_run_json(
            str(path).encode(), P(self.ids.ctypes.data), P(self.scores.ctypes.data), P(self.counts.ctypes.data),
            len(self.qids), k, P(q_blob.ctypes.data), P(q_off.ctypes.data), blob_p, off_p, ints_p,
            self.ext.size, ctypes.byref(written)))
        return "native"
