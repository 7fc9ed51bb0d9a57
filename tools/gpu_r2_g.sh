#!/bin/bash
# round-2 GPU call G (1 GPU): MUFU throughput microbenchmark; sqrt vs x*rsqrt(x) in the lean epilogue
mkdir -p gpurun_out
tools/bin/probe_mufu > gpurun_out/r2_probe_mufu.log 2>&1; echo "mufu rc=$?" | tee gpurun_out/r2_g_status.txt
rm -f gpurun_out/r2_probe_rsqrt2.log
for lib in libnw_sm100.so libnw_sm100_rsqrt.so libnw_sm100.so libnw_sm100_rsqrt.so; do
  echo "== $lib" >> gpurun_out/r2_probe_rsqrt2.log
  NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 >> gpurun_out/r2_probe_rsqrt2.log 2>&1
done
cat gpurun_out/r2_probe_mufu.log; awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' gpurun_out/r2_probe_rsqrt2.log
