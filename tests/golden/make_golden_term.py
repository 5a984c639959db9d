"""Golden vectors for the gather-sum path, produced by RUNNING THE REFERENCE'S OWN CODE (build container only):

    python tests/golden/make_golden_term.py        -> tests/golden/term_golden.npz

Executed unmodified from /root/reference: TermEncoderRetriever.get_doc_scores (scaling_retriever/indexer.py:621-641) and the
torch.topk call of TermEncoderRetriever.retrieve (:688).  Case X uses scores that are multiples of 2^-8 in [0, 8): every sum of
<= 128 of them is exact in fp32, so the outputs do not depend on the summation order (bit-exact parity); case R uses random
fp32 scores (parity within 1e-5 relative).  ujson / faiss / h5py are stubbed like in make_golden.py (not touched here)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ref_indexer, _ = import_reference()
    retr = ref_indexer.TermEncoderRetriever(torch.nn.Linear(1, 1), args=None)
    rng = np.random.default_rng(20261017)
    out = {}
    for name, (bz, vocab, n, length, k, exact) in {"X": (9, 3000, 5000, 32, 50, True), "R": (5, 40000, 3000, 16, 100, False),
                                                  "S": (3, 700, 64, 128, 64, True)}.items():
        if exact:
            pred = (rng.integers(0, 2048, size=(bz, vocab)) / 256.0).astype(np.float32)
        else:
            pred = np.log1p(rng.exponential(1.0, size=(bz, vocab))).astype(np.float32)
        codes = rng.integers(0, vocab, size=(n, length)).astype(np.int64)
        scores = retr.get_doc_scores(torch.from_numpy(pred), torch.from_numpy(codes))
        top_scores, top_idx = torch.topk(scores, k=k, dim=-1)
        out.update({f"{name}_pred": pred, f"{name}_codes": codes.astype(np.int32), f"{name}_scores": scores.numpy(),
                    f"{name}_top_scores": top_scores.numpy(), f"{name}_top_idx": top_idx.numpy(), f"{name}_k": np.int64(k)})
    np.savez_compressed(os.path.join(HERE, "term_golden.npz"), **out)
    print("wrote term_golden.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
