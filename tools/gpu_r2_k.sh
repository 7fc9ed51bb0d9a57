#!/bin/bash
# round-2 GPU call K (1 GPU): ncu launch list of the bench command (our kernels only; the synthetic-data generation
# launches hundreds of torch kernels before them)
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-aux --sustained-seconds 0.05"
$CMD > gpurun_out/r2_plain_short_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"nw_forward_kernel|rows_to_bf16_kernel|fill_kernel|logp_kernel|logp_rows_kernel|merge_side_kernel|lse_merge|labels_to_i32|class_offsets|column_" \
    -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launchlist rc=$?"
python tools/ncu_summary.py launches gpurun_out/r2_bench_launches.csv
