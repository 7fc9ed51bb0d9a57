#!/bin/bash
# round-2 GPU call AA (1 GPU): is the slow 20-step window of box 6 (19.8 ms/step at 484 W, sustained 15.0 ms) a ramp
# from idle?  Per-step kernel times + NVML samples of two bench runs (no extra warm-up / 1.5 s warm-up), then the suite.
mkdir -p gpurun_out
for ws in 0 1.5; do
  NW_BENCH_TRACE=gpurun_out/r2_aa_trace_$ws.json timeout 600 python bench.py --no-cpu-baseline --no-aux --warmup-seconds $ws \
     > gpurun_out/r2_aa_bench_$ws.json 2> gpurun_out/r2_aa_bench_$ws.err; echo "bench ws=$ws rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_aa_trace_$ws.json"))
l=json.loads(open("gpurun_out/r2_aa_bench_$ws.json").read().strip().splitlines()[-1])
print("value",round(l["value"]),"ms",round(l["ms_per_step"],2),"warmup",l["warmup"],"sust",round(l["sustained"]["value"]),"e2e",round(l["e2e"]["value"]),"pw",l["clocks"]["power_w"])
for r in d["regions"][:2]:
    print(r.get("what","timed"), r["steps"], " ".join("%.1f"%x for x in r["kernel_ms_per_step"][:40]))
w=d["regions"][1]["window"]
rows=[r for r in d["nvml"] if w[0]-0.5<=r[0]<=w[1]+0.2]
print("nvml (t-w0 ms, MHz, W):", " ".join("%d:%d:%d"%((r[0]-w[0])*1e3,r[1],r[2]) for r in rows[::4]))
PY
done
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_aa_tests.txt
