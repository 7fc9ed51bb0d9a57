#!/bin/bash
# round-2 final GPU call (1 GPU): what the driver runs at round end (GPU suite, smoke, both bench arms) on the final
# code, plus the config benchmarks, the throughput probe and the ncu launch list of the bench command
bash tools/gpu_r2_x.sh
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_x_bench_reference.json 2> gpurun_out/r2_x_bench_reference.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r2_x_bench_reference.json
bash tools/gpu_r2_k.sh 2>&1 | tail -12
