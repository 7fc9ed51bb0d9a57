#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_sched2.log
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, getattr(p,'persisting_l2_cache_max_size',None), getattr(p,'access_policy_max_window_size',None))
PY
run() {
  echo "== $1" >> gpurun_out/r2_sched2.log
  env $1 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:nw_forward_kernel -s 2 -c 1 --csv \
      --log-file gpurun_out/sched_tmp.csv python tools/probe_perf.py 4096,1280000,2048,1000 > gpurun_out/sched_tmp.log 2>&1
  grep -E "dram__bytes_read|gpu__time_duration|hit_rate" gpurun_out/sched_tmp.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' >> gpurun_out/r2_sched2.log
}
run "NW_B200_PERSIST_L2_MB=0"
run "NW_B200_PERSIST_L2_MB=32"
run "NW_B200_PERSIST_L2_MB=64"
run "NW_B200_PERSIST_L2_MB=96"
run "NW_B200_PERSIST_L2_MB=64 NW_B200_S_KEEP_MIN_GROUPS=99"
run "NW_B200_PERSIST_L2_MB=96 NW_B200_S_KEEP_MIN_GROUPS=99"
cat gpurun_out/r2_sched2.log
for cfg in "NW_X=1" "NW_B200_PERSIST_L2_MB=64" "NW_B200_PERSIST_L2_MB=64 NW_B200_S_KEEP_MIN_GROUPS=99" "NW_X=1" "NW_B200_PERSIST_L2_MB=96"; do echo "== $cfg (no ncu)"; env $cfg python tools/probe_perf.py 4096,1280000,2048,1000 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18,$19}'; done
