#!/bin/bash
# round-2 GPU call AF (1 GPU): whole GPU suite on the ABI-3 build, tensor-backward probe
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_af_tests.txt
timeout 500 python tools/probe_tensor_backward.py 2>&1 | tee gpurun_out/r2_probe_tensor_backward.txt
