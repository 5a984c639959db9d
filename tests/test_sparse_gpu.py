"""GPU parity tests of the sparse path: every check calls the CUDA kernels through the C ABI (ops.* -> ctypes ->
libb200ret.so) and compares with (a) golden vectors produced by the reference's own code (tests/golden) and (b) the
CPU oracle (oracle/) on the same seeded inputs.  Integer/index work is compared bit-exactly; fp32 scores are
compared bit-exactly too (same multiply/add order as the reference), which is stricter than the 1e-5 the spec allows.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, sparse_oracle
from scaling_retriever_b200 import ops, synth

pytestmark = pytest.mark.gpu


def dev_i32(a, cuda):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(cuda)


def dev_f32(a, cuda):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)


def build_index(off, ids, vals, n_docs, cuda):
    return ops.SparseDeviceIndex.from_csr(torch.as_tensor(off).to(cuda), dev_i32(ids, cuda), dev_f32(vals, cuda), n_docs)


# ---------------------------------------------------------------------------------------------------- CSR build

@pytest.mark.parametrize("case", ["A", "B"])
def test_csr_build_matches_reference_golden(golden, cuda, case):
    """add_batch_document + ndarray conversion of the REFERENCE (golden) vs the GPU radix-sort build: bit-exact."""
    row, col, val = golden[f"{case}_row"], golden[f"{case}_col"], golden[f"{case}_val"]
    if case == "B":   # feed order of the merged two-rank index: rank 0's postings, then rank 1's (merge_indexes)
        order = np.concatenate([np.nonzero(row % 2 == r)[0] for r in range(2)])
        row, col, val = row[order], col[order], val[order]
    n_terms, n_docs = int(golden[f"{case}_n_terms"]), int(golden[f"{case}_n_docs"])
    off, ids, w = ops.csr_build(dev_i32(row, cuda), dev_i32(col, cuda), dev_f32(val, cuda), n_terms, n_docs)
    assert np.array_equal(off.cpu().numpy(), golden[f"{case}_offsets"])
    assert np.array_equal(ids.cpu().numpy(), golden[f"{case}_ids"])
    assert np.array_equal(w.cpu().numpy().view(np.uint32), golden[f"{case}_vals"].view(np.uint32))


def test_csr_build_sort_docs_orders_unsorted_feed(golden, cuda):
    row, col, val = golden["B_row"], golden["B_col"], golden["B_val"]
    order = np.concatenate([np.nonzero(row % 2 == r)[0] for r in range(2)])
    n_terms, n_docs = int(golden["B_n_terms"]), int(golden["B_n_docs"])
    off, ids, w = ops.csr_build(dev_i32(row[order], cuda), dev_i32(col[order], cuda), dev_f32(val[order], cuda), n_terms, n_docs,
                                sort_docs=True)
    # row-major feed is already (doc asc) inside each term -> the doc-sorted build equals the oracle on the row-major feed
    o_off, o_ids, o_w = sparse_oracle.build_csr(row, col, val, n_terms)
    assert np.array_equal(off.cpu().numpy(), o_off)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert np.array_equal(w.cpu().numpy().view(np.uint32), o_w.view(np.uint32))


@pytest.mark.parametrize("n_docs,n_terms,mean_nnz", [(1, 7, 3), (300, 5000, 40), (20000, 128256, 200), (70000, 1000, 30)])
def test_csr_build_matches_oracle_random(cuda, n_docs, n_terms, mean_nnz):
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=mean_nnz, seed=7, device=cuda)
    off, ids, w = ops.csr_build(rows, cols, vals, n_terms, n_docs)
    o_off, o_ids, o_w = c_oracle.build_csr(rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy(), n_terms)
    assert np.array_equal(off.cpu().numpy(), o_off)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert np.array_equal(w.cpu().numpy().view(np.uint32), o_w.view(np.uint32))


def test_csr_build_shuffled_feed_is_stable(cuda):
    """Arbitrary feed order (not row-major): the build must keep feed order inside every term (stable)."""
    rows, cols, vals = synth.gen_sparse_docs(5000, n_terms=300, mean_nnz=20, seed=11, device="cpu")
    perm = torch.randperm(rows.numel(), generator=torch.Generator().manual_seed(3))
    rows, cols, vals = rows[perm], cols[perm], vals[perm]
    off, ids, w = ops.csr_build(rows.to(cuda), cols.to(cuda), vals.to(cuda), 300, 5000)
    o_off, o_ids, o_w = c_oracle.build_csr(rows.numpy(), cols.numpy(), vals.numpy(), 300)
    assert np.array_equal(off.cpu().numpy(), o_off)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert np.array_equal(w.cpu().numpy(), o_w)


def test_csr_build_empty(cuda):
    e = torch.empty(0, dtype=torch.int32, device=cuda)
    off, ids, w = ops.csr_build(e, e.clone(), torch.empty(0, dtype=torch.float32, device=cuda), 17, 0)
    assert off.cpu().tolist() == [0] * 18 and ids.numel() == 0 and w.numel() == 0


def test_block_table_matches_searchsorted(golden, cuda):
    off, ids = golden["C_offsets"], golden["C_ids"]
    n_docs = int(golden["C_n_docs"])
    index = build_index(off, ids, golden["C_vals"], n_docs, cuda)
    bd = index.block_docs
    n_blocks = (n_docs + bd - 1) // bd
    table = index.table.cpu().numpy().view(np.uint32)
    assert table.shape == (len(off) - 1, n_blocks + 1)
    bounds = np.arange(n_blocks + 1) * bd
    for t in range(len(off) - 1):
        lst = ids[off[t]:off[t + 1]]
        expect = off[t] + np.searchsorted(lst, bounds, side="left")
        expect[-1] = off[t + 1]
        assert np.array_equal(table[t], expect.astype(np.uint32)), t


def test_block_table_rejects_unsorted_lists(golden, cuda):
    from scaling_retriever_b200._lib import B200RetError
    with pytest.raises(B200RetError) as err:
        build_index(golden["B_offsets"], golden["B_ids"], golden["B_vals"], int(golden["B_n_docs"]), cuda)
    assert err.value.code == -4


def test_posting_layout_permutes_inside_slices_only(golden, cuda):
    """b200ret_sparse_layout: the search-side {doc, weight} array keeps every (term, doc block) slice's multiset of postings
    at the CSR positions, leaves the canonical CSR untouched, and with bank_order a window of 32 consecutive postings holds
    at most two more than the even share ceil(32*c/len) of any shared-memory bank (doc % 32, c = postings of that bank)."""
    rows, cols, vals = synth.gen_sparse_docs(40000, n_terms=300, mean_nnz=40, seed=13, device=cuda)
    off, ids, w = ops.csr_build(rows, cols, vals, 300, 40000)
    a_ids, a_w = ids.cpu().numpy(), w.cpu().numpy()
    index = ops.SparseDeviceIndex.from_csr(off, ids, w, 40000)
    assert index.doc_ids.data_ptr() == ids.data_ptr() and np.array_equal(ids.cpu().numpy(), a_ids)   # canonical CSR only read
    table = index.table.cpu().numpy().view(np.uint32)
    _, p_ids, p_w = index.csr_arrays()
    b_ids, b_w = p_ids.cpu().numpy(), p_w.cpu().numpy()
    worst = 0
    for t in range(0, 300, 7):
        for b in range(table.shape[1] - 1):
            lo, hi = int(table[t, b]), int(table[t, b + 1])
            if hi - lo == 0:
                continue
            oa, ob = np.argsort(a_ids[lo:hi]), np.argsort(b_ids[lo:hi])
            assert np.array_equal(a_ids[lo:hi][oa], b_ids[lo:hi][ob]) and np.array_equal(a_w[lo:hi][oa], b_w[lo:hi][ob])
            share = np.ceil(32.0 * np.bincount(b_ids[lo:hi] & 31, minlength=32) / (hi - lo))
            for s0 in range(0, hi - lo, 32):
                got = np.bincount(b_ids[lo + s0:min(hi, lo + s0 + 32)] & 31, minlength=32)
                worst = max(worst, int((got - share).max()))
    assert worst <= 2   # integer quantile keys: at most two more than the even share
    plain = ops.SparseDeviceIndex.from_csr(off, ids, w, 40000, bank_order=False)      # interleave only: CSR order kept
    _, q_ids, q_w = plain.csr_arrays()
    assert torch.equal(q_ids, ids) and torch.equal(q_w, w)


# ---------------------------------------------------------------------------------------------------- scoring

def golden_queries(golden):
    return golden["C_q_offsets"], golden["C_q_terms"], golden["C_q_weights"]


def test_scores_bit_exact_vs_reference_golden(golden, cuda):
    """Full score vectors vs numba_score_float's (filtered, -scores) at threshold 0: identical rows, identical bits."""
    n_docs = int(golden["C_n_docs"])
    index = build_index(golden["C_offsets"], golden["C_ids"], golden["C_vals"], n_docs, cuda)
    q_off, q_t, q_w = golden_queries(golden)
    scores = ops.sparse_scores(index, dev_i32(q_off, cuda), dev_i32(q_t, cuda), dev_f32(q_w, cuda)).cpu().numpy()
    assert scores.shape == (len(q_off) - 1, n_docs)
    for ti, thr in enumerate(golden["C_thresholds"]):
        for qi in range(len(q_off) - 1):
            filtered = np.nonzero(scores[qi] > thr)[0]
            assert np.array_equal(filtered, golden[f"C_t{ti}_q{qi}_filtered"])
            assert np.array_equal((-scores[qi][filtered]).view(np.uint32), golden[f"C_t{ti}_q{qi}_neg_scores"].view(np.uint32))


def check_topk_against_reference(ids_row, scores_row, count, ref_filtered, ref_neg_scores, k):
    """The reference returns an unordered top-k set with arbitrary ties at the boundary (argpartition).  Rule: counts
    equal; our scores (sorted desc) equal the reference's sorted scores bit-for-bit; every id we return has exactly
    the reference score; ids strictly above the k-th score are identical sets."""
    ref_scores = -ref_neg_scores
    expect = min(k, len(ref_filtered))
    assert count == expect
    assert np.all(ids_row[count:] == -1) and np.all(np.isneginf(scores_row[count:]))
    if count == 0:
        return
    ref_sorted = np.sort(ref_scores)[::-1][:count]
    assert np.array_equal(scores_row[:count].view(np.uint32), ref_sorted.view(np.uint32))
    lookup = dict(zip(ref_filtered.tolist(), ref_scores.tolist()))
    for d, s in zip(ids_row[:count].tolist(), scores_row[:count].tolist()):
        assert lookup[d] == s
    kth = ref_sorted[-1]
    above_ref = set(ref_filtered[ref_scores > kth].tolist())
    above_ours = set(ids_row[:count][scores_row[:count] > kth].tolist())
    assert above_ref == above_ours
    # our deterministic tie rule: among docs tied at the k-th score the lowest row ids are kept, rows sorted (desc, id asc)
    tied = np.sort(ref_filtered[ref_scores == kth])[:count - len(above_ref)]
    assert set(tied.tolist()) == set(ids_row[:count][scores_row[:count] == kth].tolist())
    order = np.lexsort((ids_row[:count], -scores_row[:count].astype(np.float64)))
    assert np.array_equal(order, np.arange(count))


@pytest.mark.parametrize("k", [10, 100, 1000])
def test_search_matches_reference_golden(golden, cuda, k):
    n_docs = int(golden["C_n_docs"])
    index = build_index(golden["C_offsets"], golden["C_ids"], golden["C_vals"], n_docs, cuda)
    q_off, q_t, q_w = golden_queries(golden)
    for ti, thr in enumerate(golden["C_thresholds"]):
        scores, ids, counts = ops.sparse_search(index, dev_i32(q_off, cuda), dev_i32(q_t, cuda), dev_f32(q_w, cuda), k, float(thr))
        scores, ids, counts = scores.cpu().numpy(), ids.cpu().numpy(), counts.cpu().numpy()
        for qi in range(len(q_off) - 1):
            check_topk_against_reference(ids[qi], scores[qi], int(counts[qi]), golden[f"C_t{ti}_q{qi}_filtered"],
                                         golden[f"C_t{ti}_q{qi}_neg_scores"], k)
            # and against the reference's own select_topk output (as a sorted set of scores)
            assert np.array_equal(np.sort(golden[f"C_t{ti}_q{qi}_k{k}_scores"])[::-1].view(np.uint32),
                                  scores[qi][:counts[qi]].view(np.uint32))


@pytest.mark.parametrize("n_docs,n_queries,k", [(100000, 1000, 1000), (30000, 64, 10), (3072 * 5 + 1, 33, 100)])
def test_search_matches_oracle_synthetic(cuda, n_docs, n_queries, k):
    """BASELINE config 1 shape (100k docs x Llama-3 vocab, 1k queries, top-1000) and ragged sizes vs the C oracle:
    ids and scores identical (same total order on both sides)."""
    rows, cols, vals = synth.gen_sparse_docs(n_docs, device=cuda)
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, synth.LLAMA3_VOCAB, n_docs)
    q_off, q_t, q_w = synth.gen_sparse_queries(n_queries, device=cuda)
    scores, ids, counts = ops.sparse_search(index, q_off, q_t, q_w, k, 0.0)
    o_scores, o_ids, o_counts = c_oracle.sparse_search(index.term_offsets.cpu().numpy(), index.doc_ids.cpu().numpy(),
                                                       index.weights.cpu().numpy(), n_docs, q_off.cpu().numpy(),
                                                       q_t.cpu().numpy(), q_w.cpu().numpy(), k, 0.0)
    assert np.array_equal(counts.cpu().numpy(), o_counts)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert np.array_equal(scores.cpu().numpy().view(np.uint32), o_scores.view(np.uint32))


def test_search_long_query_and_negative_threshold(cuda):
    """> 32 terms per query (two term groups in the kernel) and threshold < 0 (zero-score docs become eligible,
    like np.zeros(N) > threshold in the reference)."""
    n_docs, n_terms = 10000, 400
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=25, seed=5, device=cuda)
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    q_off, q_t, q_w = synth.gen_sparse_queries(40, n_terms=n_terms, mean_nnz=70, seed=9, device=cuda)
    assert int((q_off[1:] - q_off[:-1]).max()) > 32
    for thr, k in [(0.0, 100), (-1.0, 4096), (2.5, 50)]:
        scores, ids, counts = ops.sparse_search(index, q_off, q_t, q_w, k, thr)
        o_scores, o_ids, o_counts = c_oracle.sparse_search(index.term_offsets.cpu().numpy(), index.doc_ids.cpu().numpy(),
                                                           index.weights.cpu().numpy(), n_docs, q_off.cpu().numpy(),
                                                           q_t.cpu().numpy(), q_w.cpu().numpy(), k, thr)
        assert np.array_equal(counts.cpu().numpy(), o_counts)
        assert np.array_equal(ids.cpu().numpy(), o_ids)
        assert np.array_equal(scores.cpu().numpy().view(np.uint32), o_scores.view(np.uint32))


def test_search_ragged_query_mix_exercises_pipeline_control(cuda):
    """Runs of empty queries, one-term queries and 150-term queries (five term groups) over a corpus whose postings sit in a
    few doc blocks only: items without any posting directly behind each other, groups with only empty slices, warps that
    get no item at all — every control path of the score kernel's never-draining load pipeline, vs the C oracle."""
    bd = ops.block_docs()
    n_docs, n_terms = bd * 9 + 17, 600
    g = torch.Generator(device="cpu").manual_seed(77)
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=30, seed=21, device=cuda)
    keep = ((rows // bd) % 3 != 1) | (cols % 7 == 0)             # thin out every third doc block
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    lens = [0, 0, 0, 1, 150, 0, 1, 1, 0, 40, 0, 0, 97, 33, 32, 0, 64, 65, 0, 0, 2, 150, 0]
    q_terms, q_off = [], [0]
    for n in lens:
        t = torch.randperm(n_terms, generator=g)[:n].sort().values
        q_terms.append(t)
        q_off.append(q_off[-1] + n)
    q_t = torch.cat(q_terms).to(torch.int32).to(cuda)
    q_w = (torch.rand(q_t.numel(), generator=g) + 0.1).to(torch.float32).to(cuda)
    q_off = torch.tensor(q_off, dtype=torch.int32, device=cuda)
    for thr, k in [(0.0, 64), (-0.5, 1000)]:
        scores, ids, counts = ops.sparse_search(index, q_off, q_t, q_w, k, thr)
        o_scores, o_ids, o_counts = c_oracle.sparse_search(index.term_offsets.cpu().numpy(), index.doc_ids.cpu().numpy(),
                                                           index.weights.cpu().numpy(), n_docs, q_off.cpu().numpy(),
                                                           q_t.cpu().numpy(), q_w.cpu().numpy(), k, thr)
        assert np.array_equal(counts.cpu().numpy(), o_counts)
        assert np.array_equal(ids.cpu().numpy(), o_ids)
        assert np.array_equal(scores.cpu().numpy().view(np.uint32), o_scores.view(np.uint32))


def test_search_overflow_falls_back_to_safe_schedule(cuda):
    """Adversarial doc order: scores increase with the row id, so every later doc beats the running k-th score and
    the candidate lists overflow in the doubling rounds; the safe re-run must still return the exact top-k."""
    n_docs, n_terms, k = 3072 * 40, 8, 10
    rows = torch.arange(n_docs, dtype=torch.int32, device=cuda)
    cols = torch.zeros(n_docs, dtype=torch.int32, device=cuda)
    vals = (torch.arange(n_docs, dtype=torch.float32, device=cuda) + 1.0) / n_docs
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    q_off = torch.tensor([0, 1, 1, 2], dtype=torch.int32, device=cuda)      # middle query is empty
    q_t = torch.tensor([0, 0], dtype=torch.int32, device=cuda)
    q_w = torch.tensor([1.0, 2.0], dtype=torch.float32, device=cuda)
    scores, ids, counts = ops.sparse_search(index, q_off, q_t, q_w, k, 0.0)
    assert counts.cpu().tolist() == [k, 0, k]
    expect = np.arange(n_docs - 1, n_docs - 1 - k, -1)
    assert np.array_equal(ids[0].cpu().numpy(), expect) and np.array_equal(ids[2].cpu().numpy(), expect)
    assert np.array_equal(scores[2].cpu().numpy(), (2.0 * vals[expect.copy()].cpu().numpy()).astype(np.float32))


def test_search_doc_id_base_and_empty_inputs(cuda):
    n_docs, n_terms = 500, 50
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=5, seed=2, device=cuda)
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    q_off, q_t, q_w = synth.gen_sparse_queries(5, n_terms=n_terms, mean_nnz=4, seed=3, device=cuda)
    s0, i0, c0 = ops.sparse_search(index, q_off, q_t, q_w, 20, 0.0)
    s1, i1, c1 = ops.sparse_search(index, q_off, q_t, q_w, 20, 0.0, doc_id_base=1000)
    assert torch.equal(s0, s1) and torch.equal(c0, c1)
    assert torch.equal(torch.where(i0 >= 0, i0 + 1000, i0), i1)
    e_off = torch.zeros(1, dtype=torch.int32, device=cuda)
    s, i, c = ops.sparse_search(index, e_off, q_t[:0], q_w[:0], 20, 0.0)
    assert s.shape == (0, 20) and c.numel() == 0


def test_ops_reject_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.csr_build(torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), torch.zeros(1), 4, 1)


def test_search_item_range_chunking(cuda, monkeypatch):
    """The kernel's item counter is 32 bits wide; launch_score cuts the doc-block range into launches of < 2^31 items.  The
    test hook lowers that limit so that every round is cut into many launches: results must not change."""
    n_docs, n_terms, k = ops.block_docs() * 23 + 5, 2000, 100
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=20, seed=13, device=cuda)
    index = ops.SparseDeviceIndex.from_coo(rows, cols, vals, n_terms, n_docs)
    q_off, q_t, q_w = synth.gen_sparse_queries(50, n_terms=n_terms, mean_nnz=20, seed=14, device=cuda)
    ref = ops.sparse_search(index, q_off, q_t, q_w, k, 0.0)
    monkeypatch.setenv("B200RET_TEST_MAX_ITEMS", "120")          # 50 queries -> 2 doc blocks per launch
    cut = ops.sparse_search(index, q_off, q_t, q_w, k, 0.0)
    for a, b in zip(ref, cut):
        assert torch.equal(a, b)
    dense_ref = ops.sparse_scores(index, q_off, q_t, q_w)
    monkeypatch.delenv("B200RET_TEST_MAX_ITEMS")
    assert torch.equal(dense_ref, ops.sparse_scores(index, q_off, q_t, q_w))


@pytest.mark.parametrize("n_docs,n_terms,nq,k", [(30011, 2500, 120, 100), (ops.block_docs() * 3 + 1, 300, 33, 1000)])
def test_fp16_weight_index_matches_oracle_on_rounded_weights(cuda, n_docs, n_terms, nq, k):
    """The opt-in compressed posting format (north_star allows fp16 weights): 4-byte postings {fp16 weight,
    block-local doc id}.  Its parity mode is the oracle run on the fp16-ROUNDED weights: all N scores and the top-k rows must be
    bit-identical to that; against the fp32-weight scores the error stays within fp16 rounding (2^-11 relative per term)."""
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=40, seed=31, device=cuda)
    q_off, q_t, q_w = synth.gen_sparse_queries(nq, n_terms=n_terms, mean_nnz=15, seed=32, device=cuda)
    off, ids, w = ops.csr_build(rows, cols, vals, n_terms, n_docs)
    index16 = ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs, weight_format="fp16")
    assert index16.postings.shape == (ids.numel(),) and index16.postings.element_size() == 4      # 4 bytes per posting
    index32 = ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs)
    w16 = w.cpu().numpy().astype(np.float16).astype(np.float32)                                    # round-to-nearest-even, like the kernel
    h_off, h_ids = off.cpu().numpy(), ids.cpu().numpy()
    hq_off, hq_t, hq_w = q_off.cpu().numpy(), q_t.cpu().numpy(), q_w.cpu().numpy()

    full = ops.sparse_scores(index16, q_off, q_t, q_w).cpu().numpy()
    for qi in (0, nq // 2, nq - 1):
        ref = c_oracle.sparse_scores(h_off, h_ids, w16, n_docs, hq_t[hq_off[qi]:hq_off[qi + 1]], hq_w[hq_off[qi]:hq_off[qi + 1]])
        assert np.array_equal(full[qi].view(np.uint32), ref.view(np.uint32))
    s16, i16, c16 = (x.cpu().numpy() for x in ops.sparse_search(index16, q_off, q_t, q_w, k, 0.0))
    o_s, o_i, o_c = c_oracle.sparse_search(h_off, h_ids, w16, n_docs, hq_off, hq_t, hq_w, k)
    assert np.array_equal(c16, o_c) and np.array_equal(i16, o_i) and np.array_equal(s16.view(np.uint32), o_s.view(np.uint32))

    full32 = ops.sparse_scores(index32, q_off, q_t, q_w).cpu().numpy()
    np.testing.assert_allclose(full, full32, rtol=2.0 ** -10, atol=1e-6)
    with pytest.raises(ValueError):
        ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs, weight_format="int8")


def test_search_large_k_capacity(cuda):
    """k = 4096 (B200RET_MAX_K): the candidate capacity is k + 5 k there (with k + 2 blocks almost every query
    overflowed every round and the whole batch was re-run under the safe schedule) — results vs the oracle, and no query may
    need the safe re-run on exchangeable data (the launch count equals one geometric schedule)."""
    n_docs, n_terms, nq, k = 90000, 3000, 24, 4096
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, mean_nnz=60, seed=41, device=cuda)
    q_off, q_t, q_w = synth.gen_sparse_queries(nq, n_terms=n_terms, mean_nnz=25, seed=42, device=cuda)
    off, ids, w = ops.csr_build(rows, cols, vals, n_terms, n_docs)
    index = ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs)
    ops.profile_enable(True)
    ops.profile_read(ops.PROF_SPARSE_SCORE)
    s, i, c = ops.sparse_search(index, q_off, q_t, q_w, k, 0.0)
    _, score_launches, _ = ops.profile_read(ops.PROF_SPARSE_SCORE)
    ops.profile_enable(False)
    o_s, o_i, o_c = c_oracle.sparse_search(off.cpu().numpy(), ids.cpu().numpy(), w.cpu().numpy(), n_docs, q_off.cpu().numpy(),
                                           q_t.cpu().numpy(), q_w.cpu().numpy(), k)
    assert np.array_equal(c.cpu().numpy(), o_c) and np.array_equal(i.cpu().numpy(), o_i)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), o_s.view(np.uint32))
    n_blocks = (n_docs + ops.block_docs() - 1) // ops.block_docs()
    rounds, seen, size = 0, 0, 2
    while seen < n_blocks:                       # the geometric schedule of candidates.cuh (2 blocks, then 3x the docs seen)
        seen = n_blocks if n_blocks - seen <= size else seen + size
        size = seen * 3
        rounds += 1
    assert score_launches == rounds, (score_launches, rounds)


def test_search_property_random_small_cases(cuda):
    """Property test over many tiny random cases (SURVEY §8c list): empty queries, terms with empty lists, k > hits, positive and
    negative thresholds, doc counts that are no multiple of any tile, k from 1 to 300 — every row must equal the oracle's."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 10 ** 6), n_docs=st.integers(1, 9000), n_terms=st.integers(1, 400), nq=st.integers(1, 12),
           k=st.integers(1, 300), threshold=st.sampled_from([0.0, 0.0, 0.7, 2.5, -1.0]))
    def check(seed, n_docs, n_terms, nq, k, threshold):
        rng = np.random.default_rng(seed)
        nnz = int(rng.integers(0, 6 * n_docs + 1))
        pairs = np.unique(np.stack([rng.integers(0, n_docs, nnz), rng.integers(0, n_terms, nnz)], axis=1), axis=0) if nnz else np.zeros((0, 2), np.int64)
        rows, cols = pairs[:, 0].astype(np.int32), pairs[:, 1].astype(np.int32)
        vals = (rng.random(len(rows)) * 2 + 1e-3).astype(np.float32)
        q_terms, q_off = [], [0]
        for _ in range(nq):
            t = np.sort(rng.choice(n_terms, size=int(rng.integers(0, min(n_terms, 20) + 1)), replace=False))
            q_terms.append(t)
            q_off.append(q_off[-1] + len(t))
        q_t = np.concatenate(q_terms).astype(np.int32) if q_off[-1] else np.zeros(0, np.int32)
        q_w = (rng.random(len(q_t)) * 2 + 1e-3).astype(np.float32)
        q_off = np.asarray(q_off, dtype=np.int32)
        index = ops.SparseDeviceIndex.from_coo(torch.as_tensor(rows).to(cuda), torch.as_tensor(cols).to(cuda), torch.as_tensor(vals).to(cuda),
                                               n_terms, n_docs)
        s, i, c = ops.sparse_search(index, torch.as_tensor(q_off).to(cuda), torch.as_tensor(q_t).to(cuda), torch.as_tensor(q_w).to(cuda),
                                    k, threshold)
        o_off, o_ids, o_w = c_oracle.build_csr(rows, cols, vals, n_terms)
        o_s, o_i, o_c = c_oracle.sparse_search(o_off, o_ids, o_w, n_docs, q_off, q_t, q_w, k, threshold)
        assert np.array_equal(c.cpu().numpy(), o_c)
        assert np.array_equal(i.cpu().numpy(), o_i)
        assert np.array_equal(s.cpu().numpy().view(np.uint32), o_s.view(np.uint32))

    check()
