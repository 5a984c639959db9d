#!/bin/bash
# dev helper (run under gpurun): sparse GPU tests + bench of every tuning build under variants/ ; BENCH_ARGS selects the workload
BENCH_ARGS=${BENCH_ARGS:---workload sparse}
for lib in variants/*.so; do
  tag=$(basename $lib .so)
  if [ -z "$SKIP_TESTS" ]; then
    B200RET_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_sparse_gpu.py -m gpu -q -x 2>&1 | tail -1
  fi
  B200RET_LIB=$PWD/$lib timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.json") if l.startswith("{")][-1]); r=d["roofline"]
    share=r.get("score_kernel_share_of_step", r.get("gemm_kernel_share_of_step", 0))
    print("$tag", "qps=%.0f ms/step=%.2f e2e=%.0f frac=%.4f kernel_ms=%.2f select_ms=%.2f launches=%d clocks=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], share*d["ms_per_step"], r["select_kernels_ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
except Exception as e: print("$tag failed", e)
PY
done
