#!/bin/bash
mkdir -p gpurun_out
python tools/probe_large_backward.py > gpurun_out/r2_probe_large_backward.log 2>&1; echo "largebwd rc=$?"
python -m pytest tests/test_gpu_direct.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -3
cat gpurun_out/r2_probe_large_backward.log
