#!/bin/bash
# round-2 GPU call AS (1 GPU): L2 prefetch of the support boxes ahead of the TMA ring (NW_B200_L2_PREFETCH), with the
# 6-stage (shipped) and the 4-stage ring: does it give the 4-stage ring's DRAM traffic at the 6-stage ring's pipe utilisation?
mkdir -p gpurun_out
P=$PWD/nwhead_b200
run() {  # lib, prefetch
  NW_B200_LIB=$P/$1 NW_B200_L2_PREFETCH=$2 timeout 300 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 2 > gpurun_out/r2_as_bench.json 2> gpurun_out/r2_as_bench.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_as_bench.json").read().strip().splitlines()[-1])
s=l["sustained"]
print("$1 pf=$2", "value",round(l["value"]),"sust",round(s["value"]),"e2e",round(l["e2e"]["value"]),"MHz",round(s["sm_mhz_in_kernel"]["median"]),"pipe",round(s["tensor_pipe_busy_at_that_clock"],3), "check", l["check"]["passed"])
PY
}
for rep in 1 2; do
run libnw_sm100.so 0; run libnw_sm100.so 8; run libnw_sm100_st4.so 0; run libnw_sm100_st4.so 8; run libnw_sm100_st4.so -8; run libnw_sm100_st4.so 16
done 2>&1 | tee gpurun_out/r2_as_ab.txt
for cfg in "libnw_sm100.so 8" "libnw_sm100_st4.so 8" "libnw_sm100_st4.so -8"; do set -- $cfg
  NW_B200_LIB=$P/$1 NW_B200_L2_PREFETCH=$2 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
     --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_as_ncu.csv python tools/probe_perf.py 4096,1280000,2048,1000 > /dev/null 2>&1
  echo "== ncu $1 pf=$2"; grep -v "^==" gpurun_out/r2_as_ncu.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print('   ', ' | '.join(r[h.index('Metric Name')].split('.')[0]+' '+r[h.index('Metric Value')] for r in rows[1:]))
"
done 2>&1 | tee gpurun_out/r2_as_ncu.txt
