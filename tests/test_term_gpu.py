"""GPU parity of the gather-sum path (TermEncoderRetriever.get_doc_scores + torch.topk, reference
indexer.py:621-641, :688) through the C ABI (b200ret_term_scores / b200ret_term_search) against the reference's own outputs
(tests/golden/term_golden.npz) and the oracle on seeded inputs; then the class API with a fake encoder."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import term_oracle
from scaling_retriever.indexer import TermEncoderRetriever
from scaling_retriever_b200 import ops

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def term_golden():
    return np.load(os.path.join(HERE, "golden", "term_golden.npz"))


@pytest.mark.parametrize("case", ["X", "S", "R"])
def test_term_scores_and_search_match_reference_golden(term_golden, cuda, case):
    pred = torch.as_tensor(term_golden[f"{case}_pred"]).to(cuda)
    codes = torch.as_tensor(term_golden[f"{case}_codes"]).to(cuda)
    k = int(term_golden[f"{case}_k"])
    ref_scores, ref_top = term_golden[f"{case}_scores"], term_golden[f"{case}_top_scores"]
    scores = ops.term_scores(pred, codes).cpu().numpy()
    top_s, top_i, counts = (x.cpu().numpy() for x in ops.term_search(pred, codes, k))
    assert (counts == k).all()
    if case == "R":     # random fp32 scores: summation order differs from torch's -> tolerance of the spec (1e-5 relative)
        np.testing.assert_allclose(scores, ref_scores, rtol=1e-5)
        np.testing.assert_allclose(top_s, ref_top, rtol=1e-5)
        assert np.array_equal(top_i, term_golden["R_top_idx"]) or np.allclose(np.take_along_axis(ref_scores, top_i, axis=1), ref_top, rtol=1e-5)
    else:               # exactly representable sums: bit-exact against the reference's own output
        assert np.array_equal(scores.view(np.uint32), ref_scores.view(np.uint32))
        assert np.array_equal(top_s, ref_top)
        o_s, o_i = term_oracle.topk(ref_scores, k)
        assert np.array_equal(top_i, o_i)              # ties resolve to the lowest doc row (torch leaves them arbitrary)
    assert np.array_equal(np.take_along_axis(scores, top_i, axis=1), top_s)     # search and full-matrix entry points agree bitwise


@pytest.mark.parametrize("bz,vocab,n,length,k", [(1, 500, 100000, 16, 1000), (300, 32000, 30000, 32, 100), (7, 70000, 20011, 64, 10),
                                                 (2, 128256, 5000, 128, 4096)])
def test_term_search_vs_oracle_shapes(cuda, bz, vocab, n, length, k):
    """few queries (doc range split over CTAs), many queries (one CTA per query), a vocabulary beyond shared memory (table
    gathered from L1/L2), ragged doc count, large k — all against the oracle (bitwise: same sequential fp32 sum)."""
    rng = np.random.default_rng(bz * 31 + length)
    pred = np.log1p(rng.exponential(1.0, size=(bz, vocab))).astype(np.float32)
    codes = rng.integers(0, vocab, size=(n, length)).astype(np.int32)
    o_scores = term_oracle.get_doc_scores(pred, codes)
    o_s, o_i = term_oracle.topk(o_scores, k)
    top_s, top_i, counts = (x.cpu().numpy() for x in ops.term_search(torch.as_tensor(pred).to(cuda), torch.as_tensor(codes).to(cuda), k))
    assert np.array_equal(top_s.view(np.uint32), o_s.view(np.uint32)) and np.array_equal(top_i, o_i) and (counts == k).all()


class FakeTermEncoder(torch.nn.Module):
    def __init__(self, table):
        super().__init__()
        self.base_model = torch.nn.Linear(1, 1)
        self.base_model.device = table.device
        self.register_buffer("table", table)

    def lex_encode(self, input_ids):
        return self.table[input_ids], None          # the reference accepts (preds, None) tuples (:663-667)


def test_term_encoder_retriever_class_api(term_golden, cuda, tmp_path):
    pred, codes = term_golden["X_pred"], term_golden["X_codes"]
    k = int(term_golden["X_k"])
    docid_to_smtids = {f"doc{7 * i}": codes[i].tolist() for i in range(len(codes))}
    qids = [f"q{i}" for i in range(len(pred))]
    loader = [{"input_ids": torch.arange(i, min(i + 4, len(pred))), "queries": qids[i:i + 4]} for i in range(0, len(pred), 4)]
    retr = TermEncoderRetriever(FakeTermEncoder(torch.as_tensor(pred).to(cuda)), args=None)
    out_dir = str(tmp_path / "out")
    run = retr.retrieve(loader, docid_to_smtids, k, out_dir)
    full = retr.get_doc_scores(torch.as_tensor(pred).to(cuda), torch.as_tensor(codes.astype(np.int64)).to(cuda)).cpu().numpy()
    assert np.array_equal(full.view(np.uint32), term_golden["X_scores"].view(np.uint32))
    with open(os.path.join(out_dir, "run.json")) as f:
        text = f.read()
    assert text == json.dumps(run.to_dict())
    disk = json.loads(text)
    for b, qid in enumerate(qids):      # the reference's run: qid -> {docids[idx]: score} for torch.topk's (scores, idxes)
        want = {}
        for s, idx in zip(term_golden["X_top_scores"][b].tolist(), term_golden["X_top_idx"][b].tolist()):
            want[f"doc{7 * idx}"] = s
        assert sorted(disk[qid].values(), reverse=True) == sorted(want.values(), reverse=True)
        kth = min(want.values())
        assert {d for d, s in disk[qid].items() if s > kth} == {d for d, s in want.items() if s > kth}
    with pytest.raises(RuntimeError, match="out of range"):
        retr.retrieve(loader, docid_to_smtids, len(codes) + 1, out_dir)
