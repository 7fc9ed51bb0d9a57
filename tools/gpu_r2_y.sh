#!/bin/bash
# round-2 GPU call Y (1 GPU): influence emit with 4 epilogue sets vs 2 (NW_B200_EMIT_SETS), aux tests
timeout 600 python -m pytest tests/test_gpu_aux.py tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -2
for s in 4 2 4 2; do echo "== NW_B200_EMIT_SETS=$s"; NW_B200_EMIT_SETS=$s timeout 300 python - <<'PY' 2>&1 | grep -v "^$" | cut -c1-250
import sys; sys.path.insert(0, '.'); sys.argv=['x']
import tools.bench_configs as bc
bc.cfg5_from_features()
PY
done
