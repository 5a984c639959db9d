"""Seeded synthetic MS-MARCO-shaped inputs (BASELINE.md §3 / SURVEY.md §8d), generated with torch on any device.

Sparse: Zipf-Mandelbrot term popularity p_r ∝ 1/(r + 1 + 200) over the vocabulary, ranks scattered over term ids by
a fixed permutation; document length ~ Poisson(200) >= 1, query length ~ Poisson(40) >= 1; terms unique and ascending
inside a vector; weights log1p(Exp(1)) + 1e-3 as fp32 (SPLADE-shaped, strictly positive).
Documents are generated in fixed chunks of CHUNK_DOCS whose random streams depend only on (seed, chunk id), so a
doc-range shard [lo, hi) of the corpus is identical no matter how many ranks generate it.
The output is the row-major COO stream torch.nonzero would give SparseIndexer.index (indexer.py:259-262):
rows ascending, cols ascending inside a row.
"""
import torch

LLAMA3_VOCAB = 128256
MSMARCO_DOCS = 8841823
MSMARCO_DEV_QUERIES = 6980
CHUNK_DOCS = 65536


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def term_cdf(n_terms, shift=200.0, device="cpu"):
    r = torch.arange(n_terms, dtype=torch.float64, device=device)
    p = 1.0 / (r + 1.0 + shift)
    cdf = torch.cumsum(p / p.sum(), dim=0)
    cdf[-1] = 1.0
    return cdf


def term_permutation(n_terms, device="cpu"):
    # fixed scatter of popularity ranks over term ids (same for corpus and queries); generated on CPU for stability
    g = torch.Generator()
    g.manual_seed(977)
    return torch.randperm(n_terms, generator=g).to(device)


def _sparse_vectors(n_vec, first_id, n_terms, mean_nnz, cdf, perm, g, device):
    """n_vec sparse vectors -> (vec ids, term ids, weights), sorted by (vec, term), terms unique per vector."""
    lam = torch.full((n_vec,), float(mean_nnz), dtype=torch.float32, device=device)
    lengths = torch.poisson(lam, generator=g).clamp_(min=1).to(torch.int64)
    total = int(lengths.sum().item())
    vec = torch.repeat_interleave(torch.arange(n_vec, dtype=torch.int64, device=device), lengths, output_size=total)
    u = torch.rand(total, dtype=torch.float64, device=device, generator=g)
    rank = torch.searchsorted(cdf, u).clamp_(max=n_terms - 1)
    term = perm[rank]
    key = torch.unique(vec * n_terms + term)          # sorted, duplicates (same term twice in a vector) dropped
    vec = torch.div(key, n_terms, rounding_mode="floor")
    term = key - vec * n_terms
    w = torch.empty(key.numel(), dtype=torch.float32, device=device).exponential_(1.0, generator=g)
    w = torch.log1p(w) + 1e-3
    return (vec + first_id), term, w


def gen_sparse_docs(n_docs, n_terms=LLAMA3_VOCAB, mean_nnz=200, seed=1234, device="cpu", doc_lo=0, doc_hi=None):
    """COO postings of docs [doc_lo, doc_hi) of an n_docs corpus: (rows int32, cols int32, vals fp32)."""
    doc_hi = n_docs if doc_hi is None else doc_hi
    cdf = term_cdf(n_terms, device=device)
    perm = term_permutation(n_terms, device=device)
    rows, cols, vals = [], [], []
    for chunk in range(doc_lo // CHUNK_DOCS, (max(doc_hi, doc_lo + 1) - 1) // CHUNK_DOCS + 1):
        c_lo, c_hi = chunk * CHUNK_DOCS, min(n_docs, (chunk + 1) * CHUNK_DOCS)
        if c_hi <= c_lo or doc_hi <= doc_lo:
            break
        g = _gen(seed * 1000003 + chunk, device)
        r, c, v = _sparse_vectors(c_hi - c_lo, c_lo, n_terms, mean_nnz, cdf, perm, g, device)
        keep = (r >= doc_lo) & (r < doc_hi)
        if not bool(keep.all()):
            r, c, v = r[keep], c[keep], v[keep]
        rows.append(r.to(torch.int32))
        cols.append(c.to(torch.int32))
        vals.append(v)
    if not rows:
        e = torch.empty(0, dtype=torch.int32, device=device)
        return e, e.clone(), torch.empty(0, dtype=torch.float32, device=device)
    return torch.cat(rows), torch.cat(cols), torch.cat(vals)


def gen_sparse_queries(n_queries, n_terms=LLAMA3_VOCAB, mean_nnz=40, seed=4321, device="cpu"):
    """CSR-packed queries: (q_offsets int32[Q+1], q_terms int32, q_weights fp32), terms ascending per query."""
    cdf = term_cdf(n_terms, device=device)
    perm = term_permutation(n_terms, device=device)
    g = _gen(seed, device)
    qid, term, w = _sparse_vectors(n_queries, 0, n_terms, mean_nnz, cdf, perm, g, device)
    counts = torch.bincount(qid, minlength=n_queries)
    q_offsets = torch.zeros(n_queries + 1, dtype=torch.int64, device=device)
    q_offsets[1:] = torch.cumsum(counts, dim=0)
    return q_offsets.to(torch.int32), term.to(torch.int32), w


def queries_to_vecs(q_offsets, q_terms, q_weights):
    """CSR-packed queries -> the reference's list of (col int32 ndarray, values fp32 ndarray) (indexer.py:400-401)."""
    off = q_offsets.cpu().numpy()
    t = q_terms.cpu().numpy()
    w = q_weights.cpu().numpy()
    return [(t[off[i]:off[i + 1]], w[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def gen_dense(n, dim, seed, device="cpu", dtype=torch.float32, row_lo=0, row_hi=None, chunk=CHUNK_DOCS):
    """L2-normalised Gaussian rows [row_lo, row_hi) of an n-row matrix (chunk-seeded like the sparse corpus)."""
    row_hi = n if row_hi is None else row_hi
    out = torch.empty((max(row_hi - row_lo, 0), dim), dtype=dtype, device=device)
    for c in range(row_lo // chunk, (max(row_hi, row_lo + 1) - 1) // chunk + 1):
        c_lo, c_hi = c * chunk, min(n, (c + 1) * chunk)
        if c_hi <= c_lo or row_hi <= row_lo:
            break
        g = _gen(seed * 1000003 + c, device)
        x = torch.randn((c_hi - c_lo, dim), dtype=torch.float32, device=device, generator=g)
        x = torch.nn.functional.normalize(x, dim=1)
        a, b = max(c_lo, row_lo), min(c_hi, row_hi)
        out[a - row_lo:b - row_lo] = x[a - c_lo:b - c_lo].to(dtype)
    return out


def sparse_algorithmic_bytes(term_offsets, q_terms, n_queries, k, weight_bytes=4):
    """sum over (query, term) of df(term) * (4 + w)  +  query bytes  +  result bytes."""
    df = (term_offsets[1:] - term_offsets[:-1])
    postings = int(df[q_terms.long()].sum().item())
    return postings * (4 + weight_bytes) + int(q_terms.numel()) * 8 + n_queries * k * 8, postings
