#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0, '.'); sys.argv=['x']
import tools.bench_configs as bc
bc.cfg5_from_features()
PY
