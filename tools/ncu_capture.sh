#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU):  bash tools/ncu_capture.sh <tag>
#  (1) launch list of one bench step with per-launch device time (cold-cache, serialised: compare shares)
#  (2) --set full captures of one launch of each hot kernel (source-level, -lineinfo)
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# a step = 1 (im2col) + 1 + 12*7 + 3 encoder launches + ~58 decoder launches = ~147; skip the 3 warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 441 -c 160 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
cap() {  # tag, demangled regex, skip
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/prof_$1_${TAG} -f $BENCH > gpurun_out/ncu_$1_${TAG}.log 2>&1
}
cap gemm_lin2 "gemm_kernel<.int.256, .int.1, .int.2" 37      # 3 warm-up steps x 12 launches, then block 0's lin2 (EPI_RES_F32 pair): proj is launch 36
cap gemm_lin1 "gemm_kernel<.int.256, .int.10, .int.2" 36
cap gemm_qkv "gemm_kernel<.int.256, .int.2, .int.2" 36
cap attn_global "attention_kernel<.int.64>" 12
cap attn_windowed "attention_kernel<.int.14>" 24
cap layernorm "layernorm_kernel" 72
ls -la gpurun_out/ | tail -12
