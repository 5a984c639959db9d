"""Query-batch x corpus-size sweep (BASELINE.json configs[4]): QPS and roofline fraction of the sparse and the dense path.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.jsonl        (one GPU; one JSON line per point)

Sparse points are reported against the HBM roof on algorithmic bytes (8 B per (query, term) posting).  Dense points are
reported against the bf16 tensor peak (2*Q*N*d FLOP) AND against the HBM roof (N*d*2 corpus bytes per pass): below a few
hundred queries the GEMM is bound by streaming the corpus once, not by the tensor cores.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scaling_retriever_b200 import ops, synth  # noqa: E402

K = 1000


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
    hbm, tensor = peaks["hbm_gbs"], peaks["bf16_tflops"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    corpus_sizes = [1_000_000, 2_000_000] if args.quick else [1_000_000, 2_000_000, 5_000_000, synth.MSMARCO_DOCS]
    batches = [1, 8, 64, 512, 4096]

    q_off_all, q_t_all, q_w_all = synth.gen_sparse_queries(4096, device=dev)
    sparse_sizes = corpus_sizes if args.quick else corpus_sizes + [20_000_000]
    for n_docs in sparse_sizes:
        # one search-side index per doc range of <= 10 M docs: the kernels address postings with 32-bit positions, so the
        # 20 M-doc corpus (4.0 G postings) is held as two consecutive ranges that are searched one after the other and merged
        # (what IndexDictOfArray.device_shards does for a loaded index)
        n_parts = (n_docs + 9_999_999) // 10_000_000
        bounds = [(i * n_docs // n_parts, (i + 1) * n_docs // n_parts) for i in range(n_parts)]
        parts, df = [], None
        for lo, hi in bounds:
            rows, cols, vals = synth.gen_sparse_docs(n_docs, device=dev, doc_lo=lo, doc_hi=hi)
            rows -= lo
            off, ids, w = ops.csr_build(rows, cols, vals, synth.LLAMA3_VOCAB, hi - lo)
            del rows, cols, vals
            index = ops.SparseDeviceIndex.from_csr(off, ids, w, hi - lo)
            index.release_canonical()
            del ids, w
            torch.cuda.empty_cache()
            parts.append((index, lo))
            df = off if df is None else df + off
        for b in batches:
            q_off = q_off_all[:b + 1].contiguous()
            nq_terms = int(q_off[-1].item())
            q_t, q_w = q_t_all[:nq_terms].contiguous(), q_w_all[:nq_terms].contiguous()

            def search():
                out = [ops.sparse_search(index, q_off, q_t, q_w, K, 0.0, doc_id_base=lo) for index, lo in parts]
                if len(out) == 1:
                    return out[0]
                return ops.merge_topk(torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]), K)
            ms = timed(search, 3 if b >= 512 else 5)
            algo, postings = synth.sparse_algorithmic_bytes(df, q_t, b, K)
            print(json.dumps({"path": "sparse", "n_docs": n_docs, "batch": b, "ms": ms, "qps": b / ms * 1e3,
                              "algorithmic_gbs": algo / ms / 1e6, "frac_hbm": algo / ms / 1e6 / hbm,
                              "postings_per_query": postings / b, "doc_ranges": n_parts}), flush=True)
        del parts, index
        torch.cuda.empty_cache()

    dense_sizes = corpus_sizes if args.quick else corpus_sizes + [20_000_000]
    for dim in (2048,) if args.quick else (2048, 4096):
        q32 = synth.gen_dense(4096, dim, seed=4321, device=dev)
        q16_all = ops.f32_to_bf16(q32)
        for n_docs in dense_sizes:
            if n_docs * dim * 2 > 150e9:
                continue
            corpus = synth.gen_dense(n_docs, dim, seed=1234, device=dev, dtype=torch.bfloat16)
            for b in batches:
                q16 = q16_all[:b].contiguous()
                ms = timed(lambda: ops.dense_search(corpus, q16, K), 3)
                flops, bytes_ = 2.0 * b * n_docs * dim, n_docs * dim * 2.0
                print(json.dumps({"path": "dense", "dim": dim, "n_docs": n_docs, "batch": b, "ms": ms, "qps": b / ms * 1e3,
                                  "tflops": flops / ms / 1e9, "frac_tensor": flops / ms / 1e9 / tensor,
                                  "corpus_gbs": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / hbm,
                                  "bound": "tensor" if b >= 512 else "hbm"}), flush=True)
            del corpus
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
