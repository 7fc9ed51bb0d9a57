#!/bin/bash
# round-2 GPU call C (1 GPU): tests, top-k probe, d=512 probe + one ncu --set full of it, config benchmarks
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash tools/run_gpu_tests.sh > gpurun_out/r2_c_tests_full.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2_c_status.txt
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?" | tee -a gpurun_out/r2_c_status.txt
python tools/probe_topk.py 1280000 512 256 20 >> gpurun_out/r2_probe_topk.log 2>&1; echo "topk512 rc=$?" | tee -a gpurun_out/r2_c_status.txt
python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 > gpurun_out/r2_probe_d512.log 2>&1; echo "probe512 rc=$?" | tee -a gpurun_out/r2_c_status.txt
python tools/bench_configs.py --resnet > gpurun_out/r2_bench_configs.log 2>&1; echo "configs rc=$?" | tee -a gpurun_out/r2_c_status.txt
cp gpurun_out/bench_configs.json gpurun_out/r2_bench_configs.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 2 -c 1 \
    -o gpurun_out/r2_prof_k1_d512 python tools/probe_perf.py 4096,1280000,512,1000 > gpurun_out/r2_ncu_d512.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/r2_c_status.txt
grep -E "passed|failed|===" gpurun_out/r2_c_tests_full.log | tail -24
cat gpurun_out/r2_probe_topk.log gpurun_out/r2_probe_d512.log | tail -12
