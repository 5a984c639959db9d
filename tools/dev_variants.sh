#!/bin/bash
# dev helper (run under gpurun): bench every tuning build under variants/
for lib in variants/*.so; do
  tag=$(basename $lib .so)
  B200RET_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json")); r=d["roofline"]
    print("$tag", "qps=%.0f ms/step=%.1f e2e=%.0f frac=%.3f score_ms=%.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["score_kernel_share_of_step"]*d["ms_per_step"]))
except Exception as e: print("$tag failed", e)
PY
done
