#!/bin/bash
# round-2 final GPU call (1 GPU): what the driver runs at round end (GPU suite, smoke, both bench arms) on the final
# code, plus the ncu launch list of the bench command
bash tools/gpu_r2_am.sh
bash tools/gpu_r2_k.sh 2>&1 | tail -14
