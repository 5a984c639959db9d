#!/bin/bash
# ncu launch list of the library's own kernels over a short default bench run (sparse + dense), with DRAM bytes per launch
O=gpurun_out
K='regex:sparse_score|select_kernel|cand_|dense_search|merge_|pack_keys|unpack_keys|sort_|block_table|posting_layout|term_offsets|f32_to_bf16'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 1500 --csv \
  --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sort_scatter -c 1 -o $O/r02_prof_sort \
  python bench.py --workload sparse --steps 1 --warmup 3 --no-cpu-baseline > $O/r02_ncu_sort.log 2>&1
ls -la $O/r02_launches.csv $O/r02_prof_sort.ncu-rep
