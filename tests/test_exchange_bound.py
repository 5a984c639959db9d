"""The argument behind the tau exchange of a sharded search (include/b200ret.h (3b), shard.TauExchange), checked on the CPU:
if every shard publishes a score that at least m = ceil(k / G) of its documents reach (-inf when it has fewer), the MIN over
the shards is reached by at least k documents of the corpus, so every member of the global top-k scores >= it and survives
the strict test against the largest float below it — ties included."""
import numpy as np
from hypothesis import given, settings, strategies as st


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 8), st.integers(1, 40), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_min_of_shard_bounds_keeps_the_global_topk(n_shards, k, seed, coarse):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, 60, size=n_shards)
    scores = [(np.round(rng.standard_normal(n), 1) if coarse else rng.standard_normal(n)).astype(np.float32) for n in sizes]   # coarse: many ties
    m = -(-k // n_shards)
    published = [np.sort(s)[::-1][m - 1] if s.size >= m else np.float32(-np.inf) for s in scores]
    # a shard may publish any LOWER bound of its m-th best score (the kernel publishes a histogram bucket's lower edge)
    published = [np.float32(p - abs(rng.standard_normal()) * 0.05) if np.isfinite(p) and rng.random() < 0.5 else p for p in published]
    bound = np.float32(min(published))
    everything = np.concatenate(scores) if scores else np.empty(0, np.float32)
    if np.isfinite(bound):
        assert (everything >= bound).sum() >= k
    tau = np.nextafter(bound, np.float32(-np.inf)) if np.isfinite(bound) else bound
    topk = np.sort(everything)[::-1][:k]
    assert (topk > tau).all()          # nothing of the global top-k is filtered by `score > tau`
