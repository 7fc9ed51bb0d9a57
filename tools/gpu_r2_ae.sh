#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_tensor.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_ae_tests_bwd.txt
timeout 500 python tools/probe_tensor_backward.py 4096,1280000,2048,1000,1 1024,1280000,2048,1000,0 4096,160000,2048,1000,1 2>&1 | tee gpurun_out/r2_probe_tensor_backward_b.txt
