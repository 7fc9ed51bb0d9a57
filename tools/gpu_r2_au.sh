#!/bin/bash
# round-2 GPU call AU (1 GPU): tile gate on by default: whole GPU suite, probe A/B over the other shapes, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_au_tests.txt
for gate in 1 0 1 0; do echo "== gate $gate"; NW_B200_TILE_GATE=$gate timeout 200 python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 512,1280000,2048,1000 1024,1280000,2048,1000 4096,1280000,512,10000 4096,160000,2048,125 | cut -c1-132; done 2>&1 | tee gpurun_out/r2_au_probe.txt
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_au_bench.json 2> gpurun_out/r2_au_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_au_bench.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]),"frac",round(l["roofline"]["frac"],3), "minmax", [round(x,1) for x in l["roofline"]["kernel_ms_min_max"]], "check", l["check"], "MHz", round(l["sustained"]["sm_mhz_in_kernel"]["median"]))
print({k: (v.get("ms") or v.get("centroid_ms") or v.get("wall_us")) for k,v in l["aux"].items()}, l["aux"]["large_support_backward"].get("both_gradients",{}).get("ms"))
PY
