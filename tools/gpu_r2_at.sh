#!/bin/bash
# round-2 GPU call AT (1 GPU): per-tile producer gate (NW_B200_TILE_GATE=1): parity, DRAM bytes, sustained A/B
mkdir -p gpurun_out
NW_B200_TILE_GATE=1 timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -3
for gate in 0 1; do
  NW_B200_TILE_GATE=$gate ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
     --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_at_ncu.csv python tools/probe_perf.py 4096,1280000,2048,1000 > /dev/null 2>&1
  echo "== ncu gate=$gate"; grep -v "^==" gpurun_out/r2_at_ncu.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print('   ', ' | '.join(r[h.index('Metric Name')].split('.')[0]+' '+r[h.index('Metric Value')] for r in rows[1:]))
"
done 2>&1 | tee gpurun_out/r2_at_ncu.txt
for rep in 1 2; do for gate in 0 1; do
  NW_B200_TILE_GATE=$gate timeout 300 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 2 > gpurun_out/r2_at_bench.json 2> gpurun_out/r2_at_bench.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_at_bench.json").read().strip().splitlines()[-1])
s=l["sustained"]
print("gate=$gate", "value",round(l["value"]),"sust",round(s["value"]),"e2e",round(l["e2e"]["value"]),"MHz",round(s["sm_mhz_in_kernel"]["median"]),"pipe",round(s["tensor_pipe_busy_at_that_clock"],3), "check", l["check"]["passed"])
PY
done; done 2>&1 | tee gpurun_out/r2_at_ab.txt
