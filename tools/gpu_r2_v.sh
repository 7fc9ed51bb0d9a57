#!/bin/bash
# round-2 GPU call V (1 GPU): ncu --set full of the 4-set epilogue kernel at d=256 (source-level stall attribution)
python tools/probe_perf.py 4096,1280000,256,1000 > gpurun_out/r2_probe_d256.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 2 -c 1 \
    -o gpurun_out/r2_prof_k1_d256_quad python tools/probe_perf.py 4096,1280000,256,1000 > gpurun_out/r2_ncu_d256.log 2>&1; echo "ncu256 rc=$?"
