#!/bin/bash
# Run under gpurun: compute-sanitizer memcheck over the GPU suites and racecheck over the shared-memory heavy kernels
# (score kernel pipeline control — both posting formats —, radix scatter, posting layout, select with its bulk-copy load, merge / key
# merge, term gather, the sharded search's tau exchange: by-product bound of the radix select + tau raise, run by host threads).
# Round-1 / round-2 result: 0 errors / 0 hazards.  Output: gpurun_out/r02_sanitize.log
set -u
O=gpurun_out/r02_sanitize.log
: > $O
echo "== memcheck ==" | tee -a $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sparse_gpu.py tests/test_dense_gpu.py \
  tests/test_merge_gpu.py tests/test_term_gpu.py tests/test_metrics_gpu.py tests/test_sharded_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee -a $O
echo "memcheck exit: ${PIPESTATUS[0]}" | tee -a $O
echo "== racecheck ==" | tee -a $O
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_sparse_gpu.py \
  tests/test_merge_gpu.py tests/test_term_gpu.py tests/test_sharded_gpu.py -m gpu -q -x \
  -k "ragged or golden or merge or csr_build_matches_oracle or posting_layout or fp16 or term_scores_and_search or (threads and sparse)" \
  > gpurun_out/r02_racecheck_full.log 2>&1
rc=$?
grep -m 12 -A6 "Race reported\|hazard detected" gpurun_out/r02_racecheck_full.log | head -60 | tee -a $O    # first hazards, if any
tail -4 gpurun_out/r02_racecheck_full.log | tee -a $O
echo "racecheck exit: $rc" | tee -a $O
exit 0
