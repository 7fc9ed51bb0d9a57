#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python tools/probe_topk.py 1280000 2048 256 20 > gpurun_out/r2_probe_topk.log 2>&1; echo "topk2048 rc=$?"
python tools/probe_topk.py 1280000 512 256 20 >> gpurun_out/r2_probe_topk.log 2>&1; echo "topk512 rc=$?"
cat gpurun_out/r2_probe_topk.log
