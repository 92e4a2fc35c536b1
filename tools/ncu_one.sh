#!/bin/bash
# bash tools/ncu_one.sh <tag> <demangled-kernel-regex> <skip> <count>   (under gpurun)
TAG=$1; RX=$2; SKIP=$3; CNT=$4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RX" -s $SKIP -c $CNT -o gpurun_out/prof_${TAG} -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
