#!/bin/bash
# round-2 GPU call H (1 GPU): mainloop-only speed (epilogue skipped) and 2 vs 4 epilogue sets with the lean epilogue
mkdir -p gpurun_out
rm -f gpurun_out/r2_probe_mainloop.log
for cfg in "2 0" "4 0" "2 1" "4 1" "2 0" "4 0"; do
  set -- $cfg
  echo "== NW_B200_EPI_SETS=$1 NW_B200_DEBUG_SKIP_EPI=$2" >> gpurun_out/r2_probe_mainloop.log
  NW_B200_EPI_SETS=$1 NW_B200_DEBUG_SKIP_EPI=$2 python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 >> gpurun_out/r2_probe_mainloop.log 2>&1
done
echo "== default sets, skip=1, d=2048" >> gpurun_out/r2_probe_mainloop.log
NW_B200_DEBUG_SKIP_EPI=1 python tools/probe_perf.py 4096,1280000,2048,1000 >> gpurun_out/r2_probe_mainloop.log 2>&1
awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' gpurun_out/r2_probe_mainloop.log
