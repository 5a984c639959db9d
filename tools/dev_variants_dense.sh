#!/bin/bash
# dev helper (run under gpurun): dense GPU tests + dense bench (d = 2048 and, with DIMS="2048 4096", d = 4096) of every build under variants/
for lib in variants/*.so; do
  tag=$(basename $lib .so)
  B200RET_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_dense_gpu.py -m gpu -q -x 2>&1 | tail -1
  for dim in ${DIMS:-2048}; do
    B200RET_LIB=$PWD/$lib timeout 400 python bench.py --workload dense --dim $dim --steps 8 --warmup 3 --no-cpu-baseline ${NDOCS:+--n-docs $NDOCS} > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$tag d=$dim", "qps=%.0f ms/step=%.2f frac=%.4f sust=%.4f gemm_ms=%.2f select_ms=%.2f clocks=%s %s" % (d["value"], d["ms_per_step"], r["frac"], r["frac_of_sustained_peak"], r["gemm_kernel_share_of_step"]*d["ms_per_step"], r["select_kernels_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("$tag failed", e)
PY
  done
done
