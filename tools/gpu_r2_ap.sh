#!/bin/bash
# round-2 GPU call AP (1 GPU): do the bank re-reads depend on the k-block stride (N * 128 B = 2^18 * 625 at N = 1.28M)?
mkdir -p gpurun_out
for shape in 4096,1280000,2048,1000 4096,1281000,2048,1000 4096,1280128,2048,1 4096,1310720,2048,1024; do
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_ap_ncu.csv python tools/probe_perf.py $shape > gpurun_out/r2_ap_probe.log 2>&1
  echo "== $shape"; grep -v "^==" gpurun_out/r2_ap_ncu.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print('   ', ' | '.join(r[h.index('Metric Name')].split('.')[0]+' '+r[h.index('Metric Value')] for r in rows[1:]))
"
done 2>&1 | tee gpurun_out/r2_ap_ncu.txt
