#!/bin/bash
# round-2 GPU call D (1 GPU): tests, top-k probe, epilogue-set A/B at d=512/1024, config-4 k-means timing
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash tools/run_gpu_tests.sh > gpurun_out/r2_d_tests_full.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2_d_status.txt
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?" | tee -a gpurun_out/r2_d_status.txt
python tools/probe_topk.py 1280000 512 256 20 >> gpurun_out/r2_probe_topk.log 2>&1; echo "topk512 rc=$?" | tee -a gpurun_out/r2_d_status.txt
for sets in 2 4; do
  echo "== NW_B200_EPI_SETS=$sets" >> gpurun_out/r2_probe_sets.log
  NW_B200_EPI_SETS=$sets python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,256,1000 >> gpurun_out/r2_probe_sets.log 2>&1; echo "sets$sets rc=$?" | tee -a gpurun_out/r2_d_status.txt
done
echo "== default" >> gpurun_out/r2_probe_sets.log
python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 >> gpurun_out/r2_probe_sets.log 2>&1
python tools/bench_configs.py --cfg4 > gpurun_out/r2_cfg4.log 2>&1; echo "cfg4 rc=$?" | tee -a gpurun_out/r2_d_status.txt
grep -E "passed|failed|===" gpurun_out/r2_d_tests_full.log | tail -24
cat gpurun_out/r2_probe_topk.log gpurun_out/r2_probe_sets.log | tail -24
tail -5 gpurun_out/r2_cfg4.log
