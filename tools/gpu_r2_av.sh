#!/bin/bash
# round-2 GPU call AV (1 GPU): tile gate (shape-restricted default) vs off under the power cap, same box; then the suite
mkdir -p gpurun_out
for rep in 1 2; do for gate in 1 0; do
  NW_B200_TILE_GATE=$gate timeout 300 python bench.py --no-cpu-baseline --no-aux --sustained-seconds 2 > gpurun_out/r2_av_bench.json 2> gpurun_out/r2_av_bench.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_av_bench.json").read().strip().splitlines()[-1])
s=l["sustained"]
print("gate=$gate", "value",round(l["value"]),"sust",round(s["value"]),"e2e",round(l["e2e"]["value"]),"MHz",round(s["sm_mhz_in_kernel"]["median"]),"pipe",round(s["tensor_pipe_busy_at_that_clock"],3), "W", s["clocks"]["power_w"], "check", l["check"]["passed"])
PY
done; done 2>&1 | tee gpurun_out/r2_av_ab.txt
for gate in 1 0; do echo "== gate $gate"; NW_B200_TILE_GATE=$gate timeout 200 python tools/probe_perf.py 4096,1280000,512,10000 4096,160000,2048,125 4096,1280000,1024,1000 | cut -c1-132; done
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_av_tests.txt
