#!/bin/bash
# ncu captures for profiles/: (1) launch list with device times, (2) full-set captures of the attention and GEMM kernels.
# Usage (under gpurun): bash tools/ncu_capture.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# one step = 12 blocks x 7 + 5 encoder launches + ~75 decoder launches; skip the 3 warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 36 -c 3 -o gpurun_out/prof_attn_${TAG} -f $BENCH > gpurun_out/ncu_attn_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 153 -c 5 -o gpurun_out/prof_gemm_${TAG} -f $BENCH > gpurun_out/ncu_gemm_${TAG}.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/ncu_attn_${TAG}.log
