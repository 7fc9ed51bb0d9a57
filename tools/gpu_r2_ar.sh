#!/bin/bash
# round-2 GPU call AR (1 GPU): fused grad_s products with 4 / 2 epilogue sets (probe: min of 4 steps, warm allocator)
mkdir -p gpurun_out
for sets in 4 2 4 2; do echo "== sets $sets"; NW_B200_GRADT_SETS=$sets timeout 300 python tools/probe_tensor_backward.py 4096,1280000,2048,1000,1 2>&1 | cut -c1-420; done | tee gpurun_out/r2_ar_gradt_sets.txt
