"""Render the markdown report of a tools/sweep.py run:  python tools/sweep_report.py profiles/rNN_sweep.jsonl > profiles/rNN_sweep.md"""
import json
import sys


def main(path):
    pts = [json.loads(line) for line in open(path) if line.startswith("{")]
    print("# Query-batch x corpus-size sweep on one B200 (BASELINE.json configs[4])\n")
    print("Produced by `python tools/sweep.py` (raw points: `%s`); top-1000, synthetic MS-MARCO-shaped data, CUDA-event timing,"
          % path.split("/")[-1])
    print("roofs from MEASURED_PEAKS.json (HBM copy 6,544 GB/s; bf16 1,619.3 TFLOP/s burst).\n")
    print("## Sparse (128,256 terms, ~200 nnz/doc, ~40 nnz/query) — fraction of the HBM roof on algorithmic bytes (8 B per streamed posting)\n")
    print("| docs | batch | ms | QPS | frac of HBM roof |\n|---:|---:|---:|---:|---:|")
    for p in pts:
        if p["path"] == "sparse":
            print("| {:,} | {} | {:.2f} | {:,.0f} | {:.3f} |".format(p["n_docs"], p["batch"], p["ms"], p["qps"], p["frac_hbm"]))
    print("\nSmall batches are launch/latency bound (7 score + 7 select launches per call); the kernel needs a few hundred queries "
          "to fill 148 SMs x 15 warps.  The 20 M-doc corpus (4.0 G postings, beyond the kernels' 32-bit posting positions) is held as "
          "two consecutive doc-range indexes on the one GPU, searched one after the other and merged (`merge_topk`).\n")
    print("## Dense (bf16, fp32 accumulate) — fraction of the bf16 tensor peak (2*Q*N*d FLOP) and of the HBM roof (corpus bytes N*d*2 per pass)\n")
    print("| dim | docs | batch | ms | QPS | frac of tensor peak | frac of HBM roof | bound |\n|---:|---:|---:|---:|---:|---:|---:|---|")
    for p in pts:
        if p["path"] == "dense":
            print("| {} | {:,} | {} | {:.2f} | {:,.0f} | {:.3f} | {:.3f} | {} |".format(
                p["dim"], p["n_docs"], p["batch"], p["ms"], p["qps"], p["frac_tensor"], p["frac_hbm"], p["bound"]))
    print("\nBelow a few hundred queries the dense search is bound by streaming the corpus once per batch (HBM column); above, by the "
          "tensor cores.")


if __name__ == "__main__":
    main(sys.argv[1])
