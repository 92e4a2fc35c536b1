#!/bin/bash
# bash tools/ncu_attn.sh <tag> [grid=64] [extra ncu args]  (under gpurun): capture one launch of the fused attention op.
# Default sections are hardware-counter only (no SASS patching); pass "--set full --import-source on" for the full set.
TAG=$1; G=${2:-64}; shift; shift
EXTRA=${@:---section SpeedOfLight --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis}
mkdir -p gpurun_out
timeout 200 ncu $EXTRA --clock-control none --kernel-name-base demangled -k "regex:attention_kernel<.int.$G" -s 2 -c 1 \
  -o gpurun_out/prof_attn${G}_${TAG} -f python tools/time_attention.py > gpurun_out/ncu_attn${G}_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_attn${G}_${TAG}.log
