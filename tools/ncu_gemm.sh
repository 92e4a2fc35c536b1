#!/bin/bash
# bash tools/ncu_gemm.sh <tag> <proj|lin2|qkv|lin1> ...   (under gpurun): ncu --set full of one launch of each named encoder GEMM
TAG=$1; shift
mkdir -p gpurun_out
for W in "$@"; do
  timeout 240 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:gemm_kernel" -s 2 -c 1 \
    -o gpurun_out/prof_gemm_${W}_${TAG} -f python tools/time_gemm_one.py $W > gpurun_out/ncu_gemm_${W}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_gemm_${W}_${TAG}.log
done
