#!/bin/bash
# bash tools/ncu_proj.sh <tag>   (under gpurun): ncu --set full of one launch of the projection GEMM (fp32 residual epilogue)
TAG=$1
mkdir -p gpurun_out
timeout 240 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:gemm_kernel" -s 2 -c 1 \
  -o gpurun_out/prof_gemm_proj_${TAG} -f python tools/time_gemm_proj.py > gpurun_out/ncu_proj_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_proj_${TAG}.log
