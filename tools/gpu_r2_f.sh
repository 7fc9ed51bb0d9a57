#!/bin/bash
# round-2 GPU call F (1 GPU): tests, lazy-max epilogue A/B against the previous build, top-k probe, bench, ncu captures
mkdir -p gpurun_out
rm -f gpurun_out/tests.log gpurun_out/r2_probe_lazy.log
bash tools/run_gpu_tests.sh > gpurun_out/r2_f_tests_full.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2_f_status.txt
for lib in libnw_sm100_prev.so libnw_sm100.so libnw_sm100_prev.so libnw_sm100.so; do
  echo "== $lib" >> gpurun_out/r2_probe_lazy.log
  NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 512,1280000,2048,1000 >> gpurun_out/r2_probe_lazy.log 2>&1
done
echo "A/B rc=$?" | tee -a gpurun_out/r2_f_status.txt
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?" | tee -a gpurun_out/r2_f_status.txt
python bench.py --no-cpu-baseline --no-aux > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; echo "bench rc=$?" | tee -a gpurun_out/r2_f_status.txt
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 2 -c 1 \
    -o gpurun_out/r2_prof_k1_d512_lazy python tools/probe_perf.py 4096,1280000,512,1000 > gpurun_out/r2_ncu_d512.log 2>&1; echo "ncu512 rc=$?" | tee -a gpurun_out/r2_f_status.txt
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 2 -c 1 \
    -o gpurun_out/r2_prof_k1_d2048 python tools/probe_perf.py 4096,1280000,2048,1000 > gpurun_out/r2_ncu_d2048.log 2>&1; echo "ncu2048 rc=$?" | tee -a gpurun_out/r2_f_status.txt
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 2 -c 1 \
    -o gpurun_out/r2_prof_k1_b512 python tools/probe_perf.py 512,1280000,2048,1000 > gpurun_out/r2_ncu_b512.log 2>&1; echo "ncub512 rc=$?" | tee -a gpurun_out/r2_f_status.txt
grep -E "passed|failed|===" gpurun_out/r2_f_tests_full.log | tail -24
cat gpurun_out/r2_probe_lazy.log gpurun_out/r2_probe_topk.log | tail -36
