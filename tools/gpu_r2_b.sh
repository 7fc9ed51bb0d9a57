#!/bin/bash
# round-2 GPU call B (2 GPUs): all gpu tests incl. the 2-GPU ones, bench N=2 with the exactness check, bench N=1
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
nvidia-smi -L > gpurun_out/r2_b_smi.txt 2>&1
bash tools/run_gpu_tests.sh > gpurun_out/r2_b_tests_full.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2_b_status.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?" | tee -a gpurun_out/r2_b_status.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 20 --warmup 3 --exchange nccl --no-alt > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err; echo "bench2nccl rc=$?" | tee -a gpurun_out/r2_b_status.txt
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2_b_status.txt
grep -E "passed|failed|===" gpurun_out/r2_b_tests_full.log | tail -24
tail -c 400 gpurun_out/r2_bench_n2.err
