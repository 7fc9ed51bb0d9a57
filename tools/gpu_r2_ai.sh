#!/bin/bash
# round-2 GPU call AI (1 GPU): does the in-process NVML sampler stall the second step of a region?
mkdir -p gpurun_out
for mode in on on on; do
  NW_BENCH_SAMPLER=$mode NW_BENCH_TRACE=gpurun_out/r2_ai_trace.json timeout 600 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 1 > gpurun_out/r2_ai_bench.json 2> gpurun_out/r2_ai_bench.err; echo "sampler=$mode rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_ai_trace.json"))
l=json.loads(open("gpurun_out/r2_ai_bench.json").read().strip().splitlines()[-1])
print("  value",round(l["value"]),"ms",round(l["ms_per_step"],2),"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]))
for r in d["regions"][1:]:
    print("  host submit:", " ".join("%.1f"%x for x in r["host_submit_ms"][:6]), "| kernel:", " ".join("%.1f"%x for x in r["kernel_ms_per_step"][:6]))
PY
done
