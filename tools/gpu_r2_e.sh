#!/bin/bash
# round-2 GPU call E (1 GPU): tests, sqrt vs rsqrt epilogue A/B, top-k breakdown, k-means timing, large-N backward probe
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash tools/run_gpu_tests.sh > gpurun_out/r2_e_tests_full.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2_e_status.txt
for lib in libnw_sm100_sqrt.so libnw_sm100.so libnw_sm100_sqrt.so libnw_sm100.so; do
  echo "== $lib" >> gpurun_out/r2_probe_rsqrt.log
  NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 >> gpurun_out/r2_probe_rsqrt.log 2>&1
done
echo "rsqrt A/B rc=$?" | tee -a gpurun_out/r2_e_status.txt
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?" | tee -a gpurun_out/r2_e_status.txt
python tools/bench_configs.py --cfg4 > gpurun_out/r2_cfg4.log 2>&1; echo "cfg4 rc=$?" | tee -a gpurun_out/r2_e_status.txt
python tools/probe_large_backward.py > gpurun_out/r2_probe_large_backward.log 2>&1; echo "largebwd rc=$?" | tee -a gpurun_out/r2_e_status.txt
grep -E "passed|failed|===" gpurun_out/r2_e_tests_full.log | tail -24
cat gpurun_out/r2_probe_rsqrt.log gpurun_out/r2_probe_topk.log gpurun_out/r2_probe_large_backward.log | tail -30
tail -4 gpurun_out/r2_cfg4.log
