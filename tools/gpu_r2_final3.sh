#!/bin/bash
# round-2 final GPU call (1 GPU), tile gate on: what the driver runs at round end (GPU suite, smoke, both bench arms),
# the ncu launch list of the bench command, and the DRAM traffic of one K1 launch for profiles/roofline_traffic.json
bash tools/gpu_r2_am.sh
bash tools/gpu_r2_k.sh 2>&1 | tail -14
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
   --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_final_traffic.csv python tools/probe_perf.py 4096,1280000,2048,1000 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_final_traffic.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]: print('   ', r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')])
" | tee gpurun_out/r2_final_traffic.txt
