#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_ag_bench.json 2> gpurun_out/r2_ag_bench.err; echo "bench rc=$?"
echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_ag_bench.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"warmup",l["warmup"],"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]),"frac",round(l["roofline"]["frac"],3), "minmax", l["roofline"]["kernel_ms_min_max"], "check", l["check"]["passed"])
print(json.dumps(l["aux"].get("large_support_backward"))[:900])
print({k: (v.get("ms") or v.get("centroid_ms") or v.get("wall_us")) for k,v in l["aux"].items()})
PY
