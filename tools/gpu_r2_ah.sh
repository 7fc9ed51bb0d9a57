#!/bin/bash
# round-2 GPU call AH (1 GPU): bench with the per-step trace, three times (is the slow step reproducible? where is it?)
mkdir -p gpurun_out
for i in 1 2 3; do
  NW_BENCH_TRACE=gpurun_out/r2_ah_trace_$i.json timeout 600 python bench.py --no-cpu-baseline --no-aux > gpurun_out/r2_ah_bench_$i.json 2> gpurun_out/r2_ah_bench_$i.err; echo "bench $i rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_ah_trace_$i.json"))
l=json.loads(open("gpurun_out/r2_ah_bench_$i.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"warmup",l["warmup"],"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]))
r=d["regions"]
print(" warm-up kernel ms:", " ".join("%.1f"%x for x in r[0]["kernel_ms_per_step"]))
print(" timed kernel ms  :", " ".join("%.1f"%x for x in r[1]["kernel_ms_per_step"]))
print(" timed host submit:", " ".join("%.1f"%x for x in r[1]["host_submit_ms"]))
sus=r[2]["kernel_ms_per_step"]; big=[(i,round(x,1)) for i,x in enumerate(sus) if x>1.3*sorted(sus)[len(sus)//2]]
print(" sustained: steps", len(sus), "median %.2f"%sorted(sus)[len(sus)//2], "outliers", big[:20])
PY
done
nproc; uptime
