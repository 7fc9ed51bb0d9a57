#!/bin/bash
# round-2 GPU call L (1 GPU): ncu of the top-k refinement kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:topk_refine -s 1 -c 1 \
    -o gpurun_out/r2_prof_topk_refine python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_ncu_topk.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2_ncu_topk.log
