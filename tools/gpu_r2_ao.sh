#!/bin/bash
# round-2 GPU call AO (1 GPU): TMA ring depth 6 / 5 / 4 under the power cap (bench's >= 3 s sustained arm), same box
mkdir -p gpurun_out
P=$PWD/nwhead_b200
for rep in 1 2; do for lib in libnw_sm100.so libnw_sm100_st5.so libnw_sm100_st4.so; do
  NW_B200_LIB=$P/$lib timeout 300 python bench.py --no-cpu-baseline --no-aux > gpurun_out/r2_ao_bench.json 2> gpurun_out/r2_ao_bench.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_ao_bench.json").read().strip().splitlines()[-1])
s=l["sustained"]
print("$lib", "value",round(l["value"]),"sust",round(s["value"]),"e2e",round(l["e2e"]["value"]),"MHz",round(s["sm_mhz_in_kernel"]["median"]),"pipe",round(s["tensor_pipe_busy_at_that_clock"],3),"W",s["clocks"]["power_w"])
PY
done; done 2>&1 | tee gpurun_out/r2_ao_ab.txt
