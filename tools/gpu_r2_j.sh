#!/bin/bash
# round-2 GPU call J (1 GPU): what the driver runs (single-process gpu tests, smoke, bench both arms) + launch list + clock probes
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2_j_pytest_single_process.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/r2_j_status.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_j_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2_j_status.txt
( time python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err ); echo "bench rc=$?" | tee -a gpurun_out/r2_j_status.txt
rm -f gpurun_out/r2_probe_clock.log
for cfg in "0" "1"; do
  echo "== NW_B200_DEBUG_SKIP_EPI=$cfg" >> gpurun_out/r2_probe_clock.log
  NW_B200_DEBUG_SKIP_EPI=$cfg python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,1024,1000 4096,1280000,2048,1000 >> gpurun_out/r2_probe_clock.log 2>&1
done
rm -f gpurun_out/r2_probe_poly.log
for lib in libnw_sm100.so libnw_sm100_poly4.so libnw_sm100_poly3.so libnw_sm100_poly2.so libnw_sm100.so libnw_sm100_poly3.so; do
  echo "== $lib" >> gpurun_out/r2_probe_poly.log
  NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 >> gpurun_out/r2_probe_poly.log 2>&1
done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aux --sustained-seconds 0.05 > gpurun_out/r2_plain_short_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aux --sustained-seconds 0.05 > gpurun_out/r2_ncu_launches.log 2>&1; echo "launchlist rc=$?" | tee -a gpurun_out/r2_j_status.txt
tail -4 gpurun_out/r2_j_pytest_single_process.log; cat gpurun_out/r2_j_smoke.log | tail -2
awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18,$19}' gpurun_out/r2_probe_clock.log gpurun_out/r2_probe_poly.log
