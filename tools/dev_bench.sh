#!/bin/bash
# dev helper (run under gpurun): GPU sparse tests + full-scale bench summary + optional ncu of the score kernel
tag=$1
timeout 300 python -m pytest tests/test_sparse_gpu.py tests/test_api_gpu.py -m gpu -q -x --deselect tests/test_api_gpu.py::test_dense_store_embs_then_search_knn 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
r=d["roofline"]
print("$tag", "qps=%.0f ms/step=%.1f e2e=%.0f frac=%.3f score_share=%.3f select_ms=%.2f launches=%d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["score_kernel_share_of_step"], r["select_kernels_ms_per_step"], d["gpu_launches"]))
PY
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_score -s 16 -c 1 -o gpurun_out/prof_score_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
fi
