#!/bin/bash
# round-2 GPU call M (1 GPU): top-k probe after the gather rewrite + the tests that use it
mkdir -p gpurun_out
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?"
python tools/probe_topk.py 1280000 512 256 20 >> gpurun_out/r2_probe_topk.log 2>&1; echo "topk512 rc=$?"
python -m pytest tests/test_gpu_aux.py tests/test_gpu_nwnet.py tests/test_gpu_forward.py -q -m gpu 2>&1 | tail -3
cat gpurun_out/r2_probe_topk.log
