#!/bin/bash
mkdir -p gpurun_out
python tools/bench_configs.py --resnet > gpurun_out/r2_bench_configs.log 2>&1; echo "configs rc=$?"
cp gpurun_out/bench_configs.json gpurun_out/r2_bench_configs.json
python tools/sweep_batch.py > gpurun_out/r2_batch_sweep.log 2>&1; echo "sweep rc=$?"
tail -15 gpurun_out/r2_batch_sweep.log
