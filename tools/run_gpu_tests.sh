#!/bin/bash
# Run every GPU test file in its own process (a CUDA fault in one must not poison the others).
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_*.py; do
  echo "=== $f" | tee -a gpurun_out/tests.log
  timeout 600 python -m pytest "$f" -q -m gpu --timeout 300 2>&1 | tail -150 | tee -a gpurun_out/tests.log
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
