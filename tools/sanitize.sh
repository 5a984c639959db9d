#!/bin/bash
# Run under gpurun: compute-sanitizer memcheck over the GPU suites and racecheck over the shared-memory heavy kernels
# (score kernel pipeline control, radix scatter, posting layout, merge).  Round-1 result: 0 errors / 0 hazards.
set -u
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sparse_gpu.py tests/test_dense_gpu.py tests/test_merge_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_sparse_gpu.py tests/test_merge_gpu.py -m gpu -q -x \
  -k "ragged or golden or merge or csr_build_matches_oracle or posting_layout" 2>&1 | tail -3
