#!/bin/bash
# round-2 GPU call X (1 GPU): the whole GPU suite as the driver runs it, smoke, the bench (both arms), the config
# benchmarks and the throughput probe, all on the epilogue of this commit
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_x_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_x_bench.json 2> gpurun_out/r2_x_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_x_bench.json
timeout 300 python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 512,1280000,2048,1000 4096,1280000,512,10000 | cut -c1-200 | tee gpurun_out/r2_x_probe.txt
timeout 900 python tools/bench_configs.py > gpurun_out/r2_x_bench_configs.log 2>&1; echo "configs rc=$?"
