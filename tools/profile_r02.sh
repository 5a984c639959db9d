#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU): ncu launch lists of the bench command + one --set full capture of each
# dominant kernel.  Numbers printed by runs under ncu are never bench values.
set -x
O=gpurun_out
# (1) launch lists: per-launch duration + DRAM bytes of every kernel of the library over a short default bench run
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1
# (2) full captures: the largest round of the sparse score kernel (7 launches per step: skip 3 warm-up steps + 5 rounds), and a
#     full-size dense GEMM launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_score -s 26 -c 1 -o $O/r02_prof_score \
  python bench.py --workload sparse --steps 1 --warmup 3 --no-cpu-baseline > $O/r02_ncu_score.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_search -s 200 -c 1 -o $O/r02_prof_dense \
  python bench.py --workload dense --steps 1 --warmup 3 --no-cpu-baseline > $O/r02_ncu_dense.log 2>&1
ls -la $O/r02_*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 44 -c 3 -o $O/r02_prof_select \
  python bench.py --workload sparse --steps 1 --warmup 3 --no-cpu-baseline > $O/r02_ncu_select.log 2>&1
ls -la $O/r02_*
