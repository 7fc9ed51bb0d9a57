#!/bin/bash
# round-2 GPU call AM (1 GPU): what the driver runs at round end — GPU suite, smoke, both bench arms — on the final code
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_am_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_am_bench.json 2> gpurun_out/r2_am_bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/r2_am_bench.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_am_bench.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"warmup",l["warmup"],"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]),"frac",round(l["roofline"]["frac"],3), "minmax", [round(x,1) for x in l["roofline"]["kernel_ms_min_max"]], "check", l["check"]["passed"], "cpu", round(l["cpu_baseline"]["value"],3))
print(json.dumps(l["aux"]["large_support_backward"])[:1500])
PY
SECONDS=0; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_am_bench_reference.json 2> gpurun_out/r2_am_bench_reference.err; echo "reference arm rc=$? wall ${SECONDS}s"; cut -c1-300 gpurun_out/r2_am_bench_reference.json
