#!/bin/bash
# round-2 GPU call AC (1 GPU): tensor-core backward with the fused plumbing kernels: tests, then the probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_tensor.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_ac_tests_bwd.txt
timeout 500 python tools/probe_tensor_backward.py 2>&1 | tee gpurun_out/r2_probe_tensor_backward.txt
