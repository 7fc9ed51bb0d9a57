#!/bin/bash
# round-2 GPU call AL (1 GPU): the default bench line (what the driver runs) on the final bench.py, twice
mkdir -p gpurun_out
for i in 1 2; do
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_al_bench_$i.json 2> gpurun_out/r2_al_bench_$i.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_al_bench_$i.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"warmup",l["warmup"],"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]),"frac",round(l["roofline"]["frac"],3), "minmax", [round(x,1) for x in l["roofline"]["kernel_ms_min_max"]], "check", l["check"]["passed"], "cpu", round(l["cpu_baseline"]["value"],3))
print({k: (v.get("ms") or v.get("centroid_ms") or v.get("wall_us") or v) for k,v in l["aux"].items()})
PY
done
