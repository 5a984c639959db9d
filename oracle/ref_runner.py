"""Runs the REFERENCE'S OWN sparse retrieval code (unmodified) as a timed CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference is Python + numba and lives at /root/reference in the build container only.  `copy_reference()` (called by
`__graft_entry__.build()` there) copies the four files its sparse path needs to `baseline/_ref/` — git-ignored, never part of
the repository's history, but shipped to the GPU box with the snapshot like the built .so files:

    scaling_retriever/indexer.py                        SparseRetrieval (numba_score_float :324-344, select_topk :315-322,
                                                        _sparse_retrieve_multithreaded :405-474)
    scaling_retriever/utils/inverted_index.py           IndexDictOfArray
    scaling_retriever/utils/utils.py                    is_first_worker, ...
    scaling_retriever/modeling/losses/regulariaztion.py L0

`load()` imports that copy under stubs for the three absent third-party modules (ujson, faiss, h5py — none is touched by
the functions used here; the recipe of tests/golden/make_golden.py), without disturbing this repo's own `scaling_retriever`
overlay package in sys.modules.  Only bench.py's cpu_baseline leg and tests may use this module.
"""
import json
import os
import shutil
import sys
import tempfile
import types

FILES = ("scaling_retriever/indexer.py", "scaling_retriever/utils/inverted_index.py", "scaling_retriever/utils/utils.py",
         "scaling_retriever/modeling/losses/regulariaztion.py")


def copy_reference(src_root="/root/reference", dst_root=None):
    """Copy the reference files listed above into baseline/_ref/ (run-time copy, git-ignored).  Returns the number copied."""
    dst_root = dst_root or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    n = 0
    for rel in FILES:
        src = os.path.join(src_root, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


class Runner:
    def __init__(self, indexer_mod, inv_mod):
        self.indexer = indexer_mod
        self.inv = inv_mod

    def numba_threads(self):
        import numba
        return int(numba.get_num_threads())

    def make_retriever(self, term_offsets, doc_ids, weights, n_docs, n_terms):
        """The reference's SparseRetrieval over the given CSR: the per-term dict entries are numpy VIEWS of the CSR arrays
        (what IndexDictOfArray would hold after loading, without the 69-minute add_batch_document build)."""
        import numpy as np
        import torch
        index = self.inv.IndexDictOfArray(index_path=None, dim_voc=n_terms)
        off = np.asarray(term_offsets)
        ids = np.ascontiguousarray(doc_ids, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        index.index_doc_id = {t: ids[off[t]:off[t + 1]] for t in range(n_terms)}
        index.index_doc_value = {t: w[off[t]:off[t + 1]] for t in range(n_terms)}
        index.n = int(n_docs)
        out_dir = tempfile.mkdtemp(prefix="ref_run_")
        return self.indexer.SparseRetrieval(model=torch.nn.Linear(1, 1), config={"out_dir": out_dir}, dim_voc=n_terms,
                                            device="cpu", index_d={"index": index, "ids_mapping": range(int(n_docs))})

    def retrieve(self, retriever, sparse_query_vecs, qids, topk, threshold=0.0):
        return retriever._sparse_retrieve_multithreaded(sparse_query_vecs, qids, threshold=threshold, topk=topk)


def load(ref_root, threads=None):
    """Import the reference copy under `ref_root`; raises when it (or numba) is absent."""
    if not os.path.exists(os.path.join(ref_root, FILES[0])):
        raise FileNotFoundError(f"no reference copy under {ref_root} (made by __graft_entry__.build() where /root/reference exists)")
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    if threads:
        os.environ.setdefault("NUMBA_NUM_THREADS", str(int(threads)))
    import numba  # noqa: F401  (ImportError -> caller reports "unavailable")
    saved_modules = {k: v for k, v in sys.modules.items() if k == "scaling_retriever" or k.startswith("scaling_retriever.")}
    saved_stubs = {k: sys.modules.get(k) for k in ("ujson", "faiss", "h5py")}
    saved_path = list(sys.path)
    try:
        for k in saved_modules:
            del sys.modules[k]
        for name in ("ujson", "faiss", "h5py"):
            if saved_stubs[name] is None:
                sys.modules[name] = types.ModuleType(name)
        if saved_stubs["ujson"] is None:
            sys.modules["ujson"].dump = json.dump
            sys.modules["ujson"].load = json.load
        sys.path[:] = [ref_root] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
        import scaling_retriever.indexer as ref_indexer
        import scaling_retriever.utils.inverted_index as ref_inv
        assert os.path.abspath(ref_indexer.__file__).startswith(os.path.abspath(ref_root)), ref_indexer.__file__
        if threads:
            try:
                numba.set_num_threads(min(int(threads), numba.config.NUMBA_NUM_THREADS))
            except Exception:
                pass
        return Runner(ref_indexer, ref_inv)
    finally:
        for k in [k for k in sys.modules if k == "scaling_retriever" or k.startswith("scaling_retriever.")]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
        for name, mod in saved_stubs.items():
            if mod is None:
                sys.modules.pop(name, None)
        sys.path[:] = saved_path
