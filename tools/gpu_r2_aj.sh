#!/bin/bash
# round-2 GPU call AJ (1 GPU): host stamps of the resident step, clock probe on (default), three runs
mkdir -p gpurun_out
for mode in on on on; do
NW_BENCH_CLOCK_PROBE=$mode NW_BENCH_TRACE=gpurun_out/r2_aj_trace.json timeout 600 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 1 > gpurun_out/r2_aj_bench.json 2> gpurun_out/r2_aj_bench.err; echo "probe=$mode rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_aj_trace.json"))
l=json.loads(open("gpurun_out/r2_aj_bench.json").read().strip().splitlines()[-1])
print("  value",round(l["value"]),"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]), "regions", [r["steps"] for r in d["regions"]], "minmax", l["roofline"]["kernel_ms_min_max"])
for i,s in enumerate(d["stamps"]):
    if max(s) > 0.3 and i > 1: print("  step", i, " ".join("%.2f"%x for x in s))
PY
done
