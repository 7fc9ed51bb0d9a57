#!/bin/bash
# round-2 GPU call AD (1 GPU): ncu --set full of the tensor-core backward's four K1 launches
mkdir -p gpurun_out
python tools/ncu_tensor_backward.py > gpurun_out/r2_ad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nw_forward_kernel -s 1 -c 4 \
    -o gpurun_out/r2_prof_tensor_backward -f python tools/ncu_tensor_backward.py > gpurun_out/r2_ad_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_ad_plain.log; ls -la gpurun_out/*.ncu-rep
