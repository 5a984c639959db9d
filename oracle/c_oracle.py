"""ctypes binding of oracle/liboracle.so (the C + OpenMP restatement in sparse_oracle.c). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "sparse_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _DIR, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        p = ctypes.c_void_p
        lib.oracle_max_threads.restype = ctypes.c_int
        lib.oracle_sparse_search.restype = ctypes.c_int
        lib.oracle_sparse_search.argtypes = [p, p, p, ctypes.c_int32, p, p, p, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_float, p, p, p, ctypes.c_int32]
        lib.oracle_sparse_scores.restype = None
        lib.oracle_sparse_scores.argtypes = [p, p, p, ctypes.c_int32, p, p, ctypes.c_int32, p]
        lib.oracle_build_csr.restype = ctypes.c_int
        lib.oracle_build_csr.argtypes = [p, p, p, ctypes.c_int64, ctypes.c_int32, p, p, p]
        _lib = lib
    return _lib


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data


def max_threads():
    return int(_load().oracle_max_threads())


def sparse_search(term_offsets, doc_ids, weights, n_docs, q_offsets, q_terms, q_weights, k, threshold=0.0, n_threads=0):
    """Top-k rows sorted by (score desc, doc id asc): (scores f32[Q,k], ids i64[Q,k], counts i32[Q])."""
    lib = _load()
    to, p_to = _c(term_offsets, np.int64)
    di, p_di = _c(doc_ids, np.int32)
    w, p_w = _c(weights, np.float32)
    qo, p_qo = _c(q_offsets, np.int32)
    qt, p_qt = _c(q_terms, np.int32)
    qw, p_qw = _c(q_weights, np.float32)
    nq = len(qo) - 1
    out_scores = np.empty((nq, k), dtype=np.float32)
    out_ids = np.empty((nq, k), dtype=np.int64)
    out_counts = np.empty(nq, dtype=np.int32)
    rc = lib.oracle_sparse_search(p_to, p_di, p_w, int(n_docs), p_qo, p_qt, p_qw, nq, int(k), float(threshold),
                                  out_scores.ctypes.data, out_ids.ctypes.data, out_counts.ctypes.data, int(n_threads))
    if rc != 0:
        raise MemoryError("oracle_sparse_search: allocation failed")
    return out_scores, out_ids, out_counts


def sparse_scores(term_offsets, doc_ids, weights, n_docs, q_terms, q_weights):
    """Dense fp32 score vector of one query (bit-exactness checks)."""
    lib = _load()
    to, p_to = _c(term_offsets, np.int64)
    di, p_di = _c(doc_ids, np.int32)
    w, p_w = _c(weights, np.float32)
    qt, p_qt = _c(q_terms, np.int32)
    qw, p_qw = _c(q_weights, np.float32)
    scores = np.empty(int(n_docs), dtype=np.float32)
    lib.oracle_sparse_scores(p_to, p_di, p_w, int(n_docs), p_qt, p_qw, len(qt), scores.ctypes.data)
    return scores


def build_csr(rows, cols, vals, n_terms):
    lib = _load()
    r, p_r = _c(rows, np.int32)
    c, p_c = _c(cols, np.int32)
    v, p_v = _c(vals, np.float32)
    nnz = len(r)
    term_offsets = np.empty(n_terms + 1, dtype=np.int64)
    doc_ids = np.empty(nnz, dtype=np.int32)
    weights = np.empty(nnz, dtype=np.float32)
    if lib.oracle_build_csr(p_r, p_c, p_v, nnz, int(n_terms), term_offsets.ctypes.data, doc_ids.ctypes.data,
                            weights.ctypes.data) != 0:
        raise MemoryError("oracle_build_csr: allocation failed")
    return term_offsets, doc_ids, weights
