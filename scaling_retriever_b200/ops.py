"""Torch-tensor wrappers over the C ABI (include/b200ret.h).

PyTorch is plumbing here: it owns device memory (caching allocator) and the current stream; the compute is
the hand-written sm_100a kernels in csrc/.  Every wrapper validates device/dtype/contiguity and raises on
CPU tensors — there is no CPU fallback.
"""
from dataclasses import dataclass

import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(name, t, dtype):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {t.device}; the b200ret kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


def block_docs():
    """Documents per warp-private score tile compiled into the sparse search kernel."""
    return int(_lib.load().b200ret_sparse_block_docs())


def device_info():
    import ctypes
    sm, major, minor, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    _lib.check(_lib.load().b200ret_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor), ctypes.byref(smem)))
    return {"sm_count": sm.value, "cc": (major.value, minor.value), "smem_optin_bytes": smem.value}


PROF_SPARSE_SCORE, PROF_SPARSE_SELECT, PROF_DENSE_GEMM, PROF_CSR_SORT = 0, 1, 2, 3


def profile_enable(on=True):
    """Bracket the dominant kernels with CUDA events on the launching stream (bench.py's roofline leg)."""
    _lib.check(_lib.load().b200ret_profile_enable(int(bool(on))))


def profile_read(kind):
    """(summed ms, timed launches of `kind`, all kernel launches of the library) since the last read; syncs the device."""
    import ctypes
    ms, n, total = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.load().b200ret_profile_read(int(kind), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(total)))
    return ms.value, n.value, total.value


def csr_build(rows, cols, vals, n_terms, n_docs, sort_docs=False):
    """COO postings in feed order -> (term_offsets int64[V+1], doc_ids int32[nnz], weights fp32[nnz]).

    GPU replacement for IndexDictOfArray.add_batch_document + the ndarray conversion in save()
    (reference scaling_retriever/utils/inverted_index.py:67-76, :84-88).
    """
    lib = _lib.load()
    _check_cuda("rows", rows, torch.int32)
    _check_cuda("cols", cols, torch.int32)
    _check_cuda("vals", vals, torch.float32)
    nnz = rows.numel()
    if cols.numel() != nnz or vals.numel() != nnz:
        raise ValueError("rows, cols, vals must have the same length")
    dev = rows.device
    with torch.cuda.device(dev):
        term_offsets = torch.empty(n_terms + 1, dtype=torch.int64, device=dev)
        doc_ids = torch.empty(nnz, dtype=torch.int32, device=dev)
        weights = torch.empty(nnz, dtype=torch.float32, device=dev)
        ws_bytes = lib.b200ret_csr_build_workspace_bytes(nnz, n_terms, n_docs, int(sort_docs))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.b200ret_csr_build(_ptr(rows), _ptr(cols), _ptr(vals), nnz, n_terms, n_docs, int(sort_docs),
                                         _ptr(term_offsets), _ptr(doc_ids), _ptr(weights), _ptr(ws), ws_bytes, _stream()))
    return term_offsets, doc_ids, weights


def block_table_build(term_offsets, doc_ids, n_docs, blk_docs=None):
    """Doc-block skip table (uint32 positions stored in an int32 tensor of shape [n_terms, n_blocks+1])."""
    lib = _lib.load()
    _check_cuda("term_offsets", term_offsets, torch.int64)
    _check_cuda("doc_ids", doc_ids, torch.int32)
    blk_docs = block_docs() if blk_docs is None else blk_docs
    n_terms = term_offsets.numel() - 1
    n_blocks = (n_docs + blk_docs - 1) // blk_docs
    dev = term_offsets.device
    with torch.cuda.device(dev):
        table = torch.empty((n_terms, n_blocks + 1), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.b200ret_block_table_build(_ptr(term_offsets), _ptr(doc_ids), doc_ids.numel(), n_terms, n_docs, blk_docs,
                                                 _ptr(table), _ptr(status), _stream()))
    return table


@dataclass
class SparseDeviceIndex:
    """A doc-sorted CSR posting-list index resident in HBM: the canonical CSR (term_offsets, doc_ids, weights — bit-identical
    to the reference's per-term arrays), the doc-block skip table, and the search-side posting array the kernel streams
    ({doc id, weight} interleaved as 8-byte elements at the CSR positions, every (term, doc block) slice bank-ordered)."""
    term_offsets: torch.Tensor   # int64 [n_terms + 1]
    doc_ids: torch.Tensor        # int32 [nnz], ascending inside each term (None after release_canonical)
    weights: torch.Tensor        # fp32 [nnz] (None after release_canonical)
    table: torch.Tensor          # int32 storage of uint32 [n_terms, n_blocks + 1]
    postings: torch.Tensor       # int32 [nnz, 2]: column 0 doc id, column 1 fp32 weight bits — or, in the opt-in compressed
    #                              format, int32 [nnz]: fp16 weight << 16 | block-local doc id (weight_format == "fp16")
    n_terms: int
    n_docs: int
    block_docs: int
    weight_format: str = "fp32"

    @property
    def nnz(self):
        return self.postings.shape[0]

    @property
    def device(self):
        return self.postings.device

    def release_canonical(self):
        """Drop the canonical doc_ids/weights arrays (8 B/posting) once nothing needs to export them."""
        self.doc_ids = None
        self.weights = None

    def csr_arrays(self):
        """(term_offsets, doc_ids, weights) with lists in search order (a per-slice permutation of the canonical CSR)."""
        p = self.postings
        if self.weight_format == "fp16":
            raise NotImplementedError("csr_arrays of a compressed (fp16) index: doc ids are block-local there")
        return self.term_offsets, p[:, 0].contiguous(), p[:, 1].contiguous().view(torch.float32)

    @classmethod
    def from_csr(cls, term_offsets, doc_ids, weights, n_docs, bank_order=True, weight_format="fp32"):
        """Wrap a doc-sorted CSR: build the skip table and the search-side posting array (the CSR itself is only read).
        `weight_format="fp16"` builds the opt-in compressed array (4 bytes per posting: fp16 weight + 16-bit block-local doc id);
        searches over it score with the fp16-ROUNDED weights (see b200ret_sparse_layout_f16) — not the parity format."""
        if weight_format not in ("fp32", "fp16"):
            raise ValueError(f"weight_format must be 'fp32' or 'fp16', got {weight_format!r}")
        table = block_table_build(term_offsets, doc_ids, n_docs)
        n_terms = term_offsets.numel() - 1
        nnz = doc_ids.numel()
        lib = _lib.load()
        with torch.cuda.device(doc_ids.device):
            if weight_format == "fp32":
                postings = torch.empty((nnz, 2), dtype=torch.int32, device=doc_ids.device)
                _lib.check(lib.b200ret_sparse_layout(_ptr(table), _ptr(doc_ids), _ptr(weights), nnz, n_terms, int(n_docs),
                                                     block_docs(), int(bool(bank_order)), _ptr(postings), _stream()))
            else:
                postings = torch.empty(nnz, dtype=torch.int32, device=doc_ids.device)
                _lib.check(lib.b200ret_sparse_layout_f16(_ptr(table), _ptr(doc_ids), _ptr(weights), nnz, n_terms, int(n_docs),
                                                         block_docs(), int(bool(bank_order)), _ptr(postings), _stream()))
        return cls(term_offsets, doc_ids, weights, table, postings, n_terms, int(n_docs), block_docs(), weight_format)

    @classmethod
    def from_coo(cls, rows, cols, vals, n_terms, n_docs, weight_format="fp32"):
        """Build from COO postings in any order (lists come out ascending in doc id)."""
        term_offsets, doc_ids, weights = csr_build(rows, cols, vals, n_terms, n_docs, sort_docs=True)
        return cls.from_csr(term_offsets, doc_ids, weights, n_docs, weight_format=weight_format)


def sparse_search(index, q_offsets, q_terms, q_weights, k, threshold=0.0, doc_id_base=0, exchange=None):
    """Score a CSR-packed query batch against `index`; return (scores fp32[Q,k], ids int64[Q,k], counts int32[Q]).

    Rows are sorted by (score desc, doc id asc); slots past counts[q] hold (-inf, -1).  `exchange` (shard.TauExchange): this
    index is one doc-range shard of a sharded search whose shards exchange their bounds between the rounds.
    """
    lib = _lib.load()
    _check_cuda("q_offsets", q_offsets, torch.int32)
    _check_cuda("q_terms", q_terms, torch.int32)
    _check_cuda("q_weights", q_weights, torch.float32)
    n_queries = q_offsets.numel() - 1
    dev = index.device
    with torch.cuda.device(dev):
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        out_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)
        ws_bytes = lib.b200ret_sparse_search_workspace_bytes(n_queries, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        args = (_ptr(index.table), _ptr(index.postings), index.n_terms, index.n_docs, index.block_docs,
                _ptr(q_offsets), _ptr(q_terms), _ptr(q_weights), n_queries, k, float(threshold), int(doc_id_base),
                _ptr(out_scores), _ptr(out_ids), _ptr(out_counts), _ptr(ws), ws_bytes, _stream())
        if exchange is not None:
            if index.weight_format != "fp32":
                raise ValueError("the tau exchange of a sharded search is implemented for the fp32 posting format only")
            _lib.check(lib.b200ret_sparse_search_sharded(*args, exchange.struct(n_queries, k)))
        else:
            entry = lib.b200ret_sparse_search_f16 if index.weight_format == "fp16" else lib.b200ret_sparse_search
            _lib.check(entry(*args))
    return out_scores, out_ids, out_counts


def sparse_scores(index, q_offsets, q_terms, q_weights):
    """Full fp32 score vectors [Q, n_docs] (the `scores` array of numba_score_float, indexer.py:332-341)."""
    lib = _lib.load()
    _check_cuda("q_offsets", q_offsets, torch.int32)
    _check_cuda("q_terms", q_terms, torch.int32)
    _check_cuda("q_weights", q_weights, torch.float32)
    n_queries = q_offsets.numel() - 1
    dev = index.device
    n_blocks = (index.n_docs + index.block_docs - 1) // index.block_docs
    with torch.cuda.device(dev):
        out = torch.zeros((n_queries, n_blocks * index.block_docs), dtype=torch.float32, device=dev)
        ws = torch.empty(256, dtype=torch.uint8, device=dev)
        entry = lib.b200ret_sparse_scores_f16 if index.weight_format == "fp16" else lib.b200ret_sparse_scores
        _lib.check(entry(
            _ptr(index.table), _ptr(index.postings), index.n_terms, index.n_docs, index.block_docs,
            _ptr(q_offsets), _ptr(q_terms), _ptr(q_weights), n_queries, _ptr(out), _ptr(ws), 256, _stream()))
    return out[:, :index.n_docs]


def f32_to_bf16(src):
    lib = _lib.load()
    _check_cuda("src", src, torch.float32)
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(lib.b200ret_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()))
    return dst


def dense_search(corpus, queries, k, doc_id_base=0, exchange=None):
    """Exact top-k of queries . corpus^T (bf16 inputs, fp32 accumulate); rows sorted descending.  `exchange`: see sparse_search."""
    lib = _lib.load()
    _check_cuda("corpus", corpus, torch.bfloat16)
    _check_cuda("queries", queries, torch.bfloat16)
    n_docs, dim = corpus.shape
    n_queries = queries.shape[0]
    if queries.shape[1] != dim:
        raise ValueError(f"dim mismatch: corpus {dim}, queries {queries.shape[1]}")
    dev = corpus.device
    with torch.cuda.device(dev):
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        out_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)
        ws_bytes = lib.b200ret_dense_search_workspace_bytes(n_queries, n_docs, dim, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        args = (_ptr(corpus), _ptr(queries), n_docs, n_queries, dim, k, int(doc_id_base),
                _ptr(out_scores), _ptr(out_ids), _ptr(out_counts), _ptr(ws), ws_bytes, _stream())
        if exchange is not None:
            _lib.check(lib.b200ret_dense_search_sharded(*args, exchange.struct(n_queries, k)))
        else:
            _lib.check(lib.b200ret_dense_search(*args))
    return out_scores, out_ids, out_counts


def merge_topk(scores, ids, k):
    """Merge per-shard rows [G, Q, k] into the global top-k [Q, k] (same total order as the search kernels)."""
    lib = _lib.load()
    _check_cuda("scores", scores, torch.float32)
    _check_cuda("ids", ids, torch.int64)
    n_shards, n_queries, kk = scores.shape
    if kk != k or ids.shape != scores.shape:
        raise ValueError("scores/ids must both be [n_shards, n_queries, k]")
    dev = scores.device
    per_pass = max(2, merge_max_shards(k))
    while n_shards > per_pass:     # large G * k: merge groups of shards first (group order keeps shard order)
        parts = [merge_topk(scores[a:a + per_pass].contiguous(), ids[a:a + per_pass].contiguous(), k)[:2]
                 for a in range(0, n_shards, per_pass)]
        scores, ids = torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts])
        n_shards = scores.shape[0]
    with torch.cuda.device(dev):
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        out_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)
        _lib.check(lib.b200ret_merge_topk(_ptr(scores), _ptr(ids), n_shards, n_queries, k,
                                          _ptr(out_scores), _ptr(out_ids), _ptr(out_counts), _stream()))
    return out_scores, out_ids, out_counts


def merge_max_shards(k):
    """Shards one merge pass can take for this k (shared-memory bound of the merge kernels)."""
    return int(_lib.load().b200ret_merge_max_shards(int(k)))


def pack_keys(scores, ids, out=None):
    """(scores fp32, global ids int64; id -1 = padding) -> packed 64-bit keys (int64 storage), same shape."""
    lib = _lib.load()
    _check_cuda("scores", scores, torch.float32)
    _check_cuda("ids", ids, torch.int64)
    if scores.shape != ids.shape:
        raise ValueError("scores/ids shapes differ")
    keys = torch.empty(scores.shape, dtype=torch.int64, device=scores.device) if out is None else out
    with torch.cuda.device(scores.device):
        _lib.check(lib.b200ret_pack_keys(_ptr(scores), _ptr(ids), scores.numel(), _ptr(keys), _stream()))
    return keys


def merge_keys(keys, k):
    """Packed keys [G, Q, k] (rows sorted descending, zero padded) -> [Q, k] best keys per query, sorted descending.
    More shards than one pass takes (large G * k) are merged in several passes."""
    lib = _lib.load()
    _check_cuda("keys", keys, torch.int64)
    n_shards, n_queries, kk = keys.shape
    if kk != k:
        raise ValueError("keys must be [n_shards, n_queries, k]")
    per_pass = max(2, merge_max_shards(k))
    with torch.cuda.device(keys.device):
        while True:
            g = keys.shape[0]
            if g <= per_pass:
                out = torch.empty((n_queries, k), dtype=torch.int64, device=keys.device)
                _lib.check(lib.b200ret_merge_keys(_ptr(keys), g, n_queries, k, _ptr(out), _stream()))
                return out
            parts = []
            for a in range(0, g, per_pass):
                grp = keys[a:a + per_pass].contiguous()
                out = torch.empty((n_queries, k), dtype=torch.int64, device=keys.device)
                _lib.check(lib.b200ret_merge_keys(_ptr(grp), grp.shape[0], n_queries, k, _ptr(out), _stream()))
                parts.append(out)
            keys = torch.stack(parts)


def unpack_keys(keys, k):
    """Packed keys [Q, k] -> (scores fp32 [Q,k], ids int64 [Q,k], counts int32 [Q]) with (-inf, -1) padding."""
    lib = _lib.load()
    _check_cuda("keys", keys, torch.int64)
    n_queries = keys.shape[0]
    dev = keys.device
    with torch.cuda.device(dev):
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        out_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)
        _lib.check(lib.b200ret_unpack_keys(_ptr(keys), n_queries, k, _ptr(out_scores), _ptr(out_ids), _ptr(out_counts), _stream()))
    return out_scores, out_ids, out_counts


def term_scores(pred, codes):
    """doc_scores[b, n] = sum_l pred[b, codes[n, l]] as a full fp32 [Q, N] matrix (TermEncoderRetriever.get_doc_scores)."""
    lib = _lib.load()
    _check_cuda("pred", pred, torch.float32)
    _check_cuda("codes", codes, torch.int32)
    n_queries, n_vocab = pred.shape
    n_docs, code_len = codes.shape
    with torch.cuda.device(pred.device):
        out = torch.empty((n_queries, n_docs), dtype=torch.float32, device=pred.device)
        _lib.check(lib.b200ret_term_scores(_ptr(pred), _ptr(codes), n_queries, n_vocab, n_docs, code_len, _ptr(out), _stream()))
    return out


def term_search(pred, codes, k):
    """Exact top-k of the gather-sum scores without materialising them: (scores fp32 [Q,k] desc, rows int64 [Q,k], counts)."""
    lib = _lib.load()
    _check_cuda("pred", pred, torch.float32)
    _check_cuda("codes", codes, torch.int32)
    n_queries, n_vocab = pred.shape
    n_docs, code_len = codes.shape
    dev = pred.device
    with torch.cuda.device(dev):
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        out_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)
        ws_bytes = lib.b200ret_term_search_workspace_bytes(n_queries, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.b200ret_term_search(_ptr(pred), _ptr(codes), n_queries, n_vocab, n_docs, code_len, k,
                                           _ptr(out_scores), _ptr(out_ids), _ptr(out_counts), _ptr(ws), ws_bytes, _stream()))
    return out_scores, out_ids, out_counts


def rank_metrics(ids, counts, rel_offsets, rel_ids, mrr_cut=10, recall_cuts=(10, 100, 1000)):
    """Per-query (reciprocal rank @mrr_cut fp32 [Q], recall @cuts fp32 [Q, len(cuts)]) of result rows `ids` [Q, k] against
    CSR-packed relevant row labels (rel_offsets int64 [Q+1], rel_ids int64 ascending per query)."""
    import ctypes
    lib = _lib.load()
    _check_cuda("ids", ids, torch.int64)
    _check_cuda("rel_offsets", rel_offsets, torch.int64)
    _check_cuda("rel_ids", rel_ids, torch.int64)
    if counts is not None:
        _check_cuda("counts", counts, torch.int32)
    n_queries, k = ids.shape
    cuts = (ctypes.c_int32 * max(len(recall_cuts), 1))(*[int(c) for c in recall_cuts])
    dev = ids.device
    with torch.cuda.device(dev):
        rr = torch.empty(n_queries, dtype=torch.float32, device=dev)
        recall = torch.empty((n_queries, len(recall_cuts)), dtype=torch.float32, device=dev)
        _lib.check(lib.b200ret_rank_metrics(_ptr(ids), _ptr(counts), n_queries, k, _ptr(rel_offsets), _ptr(rel_ids), int(mrr_cut),
                                            cuts, len(recall_cuts), _ptr(rr), _ptr(recall), _stream()))
    return rr, recall
