#!/bin/bash
# 8-GPU bench exactly as the driver launches it (torchrun, NCCL)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err
grep "^{" gpurun_out/r2q_bench_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], 'value', round(d['value'],1), 'per_gpu', round(d['per_gpu'],1), 'ms', round(d['ms_per_step'],2), 'comm_ms_per_step', d['comm_ms_per_step'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['clocks'])"
tail -2 gpurun_out/r2q_bench_n8.err
