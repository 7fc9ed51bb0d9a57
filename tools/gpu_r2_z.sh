#!/bin/bash
# round-2 GPU call Z (1 GPU): whole GPU suite + config-1 head call (event-timed, graph replay, per-kernel list under ncu)
timeout 1200 python -m pytest tests/ -q -x -m gpu 2>&1 | tail -2
python tools/probe_cfg1_launches.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_cfg1_launches.csv python tools/probe_cfg1_launches.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_cfg1_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size') if 'Grid Size' in h else None; bi=h.index('Block Size') if 'Block Size' in h else None
for r in rows[-6:]:
    print(r[vi].rjust(10), (r[gi] if gi else ''), (r[bi] if bi else ''), r[ki][:110])
PY
python tools/probe_perf.py 4096,1280000,512,1000 4096,1280000,2048,1000 | cut -c1-130
