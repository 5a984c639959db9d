"""Dense golden fixture from an EXACT construction (no faiss needed, none is installable here):

    python tests/golden/make_golden_dense.py        -> tests/golden/dense_golden.npz

faiss.IndexFlatIP.search (reference scaling_retriever/indexer.py:210-214) returns the exact top-k inner products.  The vectors
below are built so that every inner product is exactly representable whatever the arithmetic (fp32 sgemm blocks in faiss, bf16
inputs with fp32 accumulation on tensor cores here) and no two documents tie for a query:
  * dims 0..d-3: entries in {-1, 0, 1}  -> integer part of the score, |.| <= d - 2 < 2^7;
  * dims d-2, d-1: the doc stores (n mod 256, n div 256) (integers < 256: exact in bf16), the query 2^-16 and 2^-8
    -> a fraction n / 2^16 unique to the document (n < 2^16), 16 fractional bits.
Every score (and every partial sum of it, in any order) is a multiple of 2^-16 below 2^7: it fits fp32's 24-bit significand,
so ANY correct IndexFlatIP implementation must return exactly the ids and scores stored here (computed in int64), bit for bit."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(86420)
    n, d, nq, k = 12000, 64, 48, 100
    docs = rng.integers(-1, 2, size=(n, d)).astype(np.int16)
    queries = rng.integers(-1, 2, size=(nq, d)).astype(np.float64)
    ids = np.arange(n)
    docs[:, d - 2] = ids % 256
    docs[:, d - 1] = ids // 256
    queries[:, d - 2] = 2.0 ** -16
    queries[:, d - 1] = 2.0 ** -8
    # exact scores in units of 2^-16 (int64)
    q_int = np.rint(queries * 65536).astype(np.int64)
    units = q_int @ docs.astype(np.int64).T                                   # [nq, n]
    assert all(len(np.unique(row)) == n for row in units)                     # no ties by construction
    order = np.argsort(-units, axis=1, kind="stable")[:, :k]
    scores = (np.take_along_axis(units, order, axis=1) / 65536.0).astype(np.float32)
    assert np.array_equal((scores.astype(np.float64) * 65536).astype(np.int64), np.take_along_axis(units, order, axis=1))
    np.savez_compressed(os.path.join(HERE, "dense_golden.npz"), docs=docs.astype(np.int16), queries=queries.astype(np.float32),
                        top_ids=order.astype(np.int64), top_scores=scores, k=np.int64(k))
    print("wrote dense_golden.npz", docs.shape, queries.shape, order.shape)


if __name__ == "__main__":
    main()
