#!/bin/bash
# scheduling experiments on K1's DRAM re-reads: chunk count and L2 policy of the support tiles (dram bytes + duration per launch)
mkdir -p gpurun_out
rm -f gpurun_out/r2_sched.log
run() {
  echo "== $1" >> gpurun_out/r2_sched.log
  env $1 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:nw_forward_kernel -s 2 -c 2 --csv \
      --log-file gpurun_out/sched_tmp.csv python tools/probe_perf.py 4096,1280000,2048,1000 > gpurun_out/sched_tmp.log 2>&1
  grep -E "dram__bytes_read|gpu__time_duration|hit_rate" gpurun_out/sched_tmp.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' >> gpurun_out/r2_sched.log
  grep "TFLOP" gpurun_out/sched_tmp.log | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' >> gpurun_out/r2_sched.log
}
run "NW_X=0"
run "NW_B200_STAGGER_NS=500"
run "NW_B200_STAGGER_NS=2000"
run "NW_B200_STAGGER_NS=8000"
cat gpurun_out/r2_sched.log
for st in 0 2000 8000 0 2000; do echo "== stagger $st (no ncu)"; NW_B200_STAGGER_NS=$st python tools/probe_perf.py 4096,1280000,2048,1000 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18,$19}'; done
