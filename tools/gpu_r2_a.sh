#!/bin/bash
# round-2 GPU call A: tests, bench (both arms), sanitizer
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_a_smi.txt 2>&1
free -g | head -2 >> gpurun_out/r2_a_smi.txt; nproc >> gpurun_out/r2_a_smi.txt
bash tools/run_gpu_tests.sh > gpurun_out/r2_a_tests_full.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2_a_status.txt
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ); echo "bench rc=$?" | tee -a gpurun_out/r2_a_status.txt
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ); echo "ref rc=$?" | tee -a gpurun_out/r2_a_status.txt
bash tools/run_sanitizer.sh > gpurun_out/r2_a_sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/r2_a_status.txt
tail -3 gpurun_out/tests.log; tail -c 600 gpurun_out/r2_bench_n1.err; tail -8 gpurun_out/r2_a_sanitizer.log
