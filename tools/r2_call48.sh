#!/bin/bash
# 2-GPU bench exactly as the driver launches it (torchrun, NCCL), both arms
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
python tools/bench_print.py gpurun_out/r2n_bench_n2.json | head -3; python - <<'P'
import json; d=json.load(open("gpurun_out/r2n_bench_n2.json")); print("n_gpus", d["n_gpus"], "value", d["value"], "per_gpu", d["per_gpu"], "comm_ms_per_step", d["comm_ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"])
P
tail -3 gpurun_out/r2n_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
