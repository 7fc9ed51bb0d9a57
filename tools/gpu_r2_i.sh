#!/bin/bash
# round-2 GPU call I (multi-GPU box): the 2-GPU tests, then bench at the GPU counts given as arguments (default 8 4 2)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_i_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_serving.py -q -m gpu --timeout 300 > gpurun_out/r2_i_serving_tests.log 2>&1; echo "serving tests rc=$?" | tee gpurun_out/r2_i_status.txt
port=29600
counts="$@"; [ -z "$counts" ] && counts="8 4 2"
for n in $counts; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "bench$n rc=$?" | tee -a gpurun_out/r2_i_status.txt
done
tail -5 gpurun_out/r2_i_serving_tests.log
for n in $counts; do grep "^{" gpurun_out/r2_bench_n$n.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['n_gpus'], 'value %.0f e2e %.0f sustained %.0f' % (d['value'], d['e2e']['value'], d['sustained']['value']), {k: v for k, v in d['check'].items() if 'max_abs_logp' in k or k == 'passed'}, d.get('alt', {}).get('query_sharded_replicated_bank', {}).get('value'))
"; tail -3 gpurun_out/r2_bench_n$n.err; done
