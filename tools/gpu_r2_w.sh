#!/bin/bash
# round-2 GPU call W (1 GPU): forward/scale/aux tests + throughput probe
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py tests/test_gpu_aux.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 4096,1280000,512,10000 | cut -c1-130
