#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"scores_kernel|aggregate|coef_kernel|grad_q|grad_s" -c 60 --csv --log-file gpurun_out/r2_large_backward_launches.csv python tools/probe_large_backward.py > gpurun_out/r2_ncu_lb.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_large_backward_launches.csv') if not l.startswith('=='))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); ii=hdr.index('ID')
seen={}
for r in rows[1:]:
    seen.setdefault(r[ii],{'k':r[ki].split('(')[0][-40:]})[r[mi]]=r[vi]+' '+r[ui]
for i,v in list(seen.items())[-24:]: print(i, v)
PY
