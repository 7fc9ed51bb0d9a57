#!/bin/bash
# round-2 GPU call AN (1 GPU): does the TMA ring depth (6 / 4 / 3 stages) change the bank re-reads from HBM?
# (the 4-stage coefficient emit read 1.5x its bank under ncu, the 6-stage forward reads 2.8x)
mkdir -p gpurun_out
P=$PWD/nwhead_b200
for rep in 1 2; do for lib in libnw_sm100.so libnw_sm100_st4.so libnw_sm100_st3.so; do
  echo "== $lib"; NW_B200_LIB=$P/$lib timeout 200 python tools/probe_perf.py 4096,1280000,2048,1000 512,1280000,2048,1000 | cut -c1-150
done; done 2>&1 | tee gpurun_out/r2_an_probe.txt
for lib in libnw_sm100.so libnw_sm100_st4.so libnw_sm100_st3.so; do
  NW_B200_LIB=$P/$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_an_ncu_$lib.csv python tools/probe_perf.py 4096,1280000,2048,1000 > /dev/null 2>&1
  echo "== $lib"; grep -v "^==" gpurun_out/r2_an_ncu_$lib.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]: print('   ', r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')])
"
done 2>&1 | tee gpurun_out/r2_an_ncu.txt
