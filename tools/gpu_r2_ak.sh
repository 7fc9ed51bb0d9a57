#!/bin/bash
# round-2 GPU call AK (1 GPU): host-side stage timer inside the forward call (NW_B200_TRACE_HOST=1)
mkdir -p gpurun_out
for i in 1 2; do
NW_B200_TRACE_HOST=1 timeout 600 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 1 > gpurun_out/r2_ak_bench.json 2> gpurun_out/r2_ak_bench.err; echo "rc=$?"
grep "host trace" gpurun_out/r2_ak_bench.err | head -30
done
