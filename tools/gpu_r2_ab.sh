#!/bin/bash
# round-2 GPU call AB (1 GPU): tensor-core backward tests, the forward / scale suites on the changed K1, throughput probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_tensor.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_ab_tests_bwd.txt
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py tests/test_gpu_aux.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_ab_tests_fwd.txt
timeout 300 python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,2048,1000 512,1280000,2048,1000 | cut -c1-200 | tee gpurun_out/r2_ab_probe.txt
