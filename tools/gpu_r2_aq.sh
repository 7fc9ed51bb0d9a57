#!/bin/bash
# round-2 GPU call AQ (1 GPU): k-block rotation between the query groups that share a support tile (NW_B200_KROT=1/0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py tests/test_gpu_backward_tensor.py tests/test_gpu_aux.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_aq_tests.txt
for krot in 1 0; do
  NW_B200_KROT=$krot ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_aq_ncu.csv python tools/probe_perf.py 4096,1280000,2048,1000 > gpurun_out/r2_aq_probe.log 2>&1
  echo "== ncu KROT=$krot"; grep -v "^==" gpurun_out/r2_aq_ncu.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print('   ', ' | '.join(r[h.index('Metric Name')].split('.')[0]+' '+r[h.index('Metric Value')] for r in rows[1:]))
"
done 2>&1 | tee gpurun_out/r2_aq_ncu.txt
for rep in 1 2; do for krot in 1 0; do
  NW_B200_KROT=$krot timeout 300 python bench.py --no-cpu-baseline --no-aux > gpurun_out/r2_aq_bench.json 2> gpurun_out/r2_aq_bench.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_aq_bench.json").read().strip().splitlines()[-1])
s=l["sustained"]
print("KROT=$krot", "value",round(l["value"]),"sust",round(s["value"]),"e2e",round(l["e2e"]["value"]),"MHz",round(s["sm_mhz_in_kernel"]["median"]),"pipe",round(s["tensor_pipe_busy_at_that_clock"],3),"W",s["clocks"]["power_w"], "check", l["check"]["passed"], l["check"]["prob_err_vs_fp64"])
PY
done; done 2>&1 | tee gpurun_out/r2_aq_ab.txt
NW_B200_KROT=1 timeout 200 python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 512,1280000,2048,1000 | cut -c1-150
NW_B200_KROT=0 timeout 200 python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,1024,1000 512,1280000,2048,1000 | cut -c1-150
