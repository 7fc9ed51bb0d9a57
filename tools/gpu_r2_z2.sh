#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py tests/test_gpu_nwnet.py tests/test_gpu_dropin.py tests/test_gpu_aux.py -q -x -m gpu 2>&1 | tail -2
python tools/probe_cfg4_launches.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_cfg4_launches.csv python tools/probe_cfg4_launches.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_cfg4_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[-5:]:
    print(r[vi].rjust(10), r[gi], r[bi], r[ki][:110])
PY
python tools/probe_cfg1_launches.py
