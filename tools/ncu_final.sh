#!/bin/bash
# Final-build ncu evidence (run under gpurun, 1 GPU):  bash tools/ncu_final.sh <tag>
#  (1) launch list of one bench step (per-launch device time; cold-cache, serialised: compare shares)
#  (2) hardware-counter sections + PC sampling of one launch of both attention kernels (no SASS patching: the
#      --set full replay of the 168-register global kernel does not launch once patched)
TAG=${1:-r1g}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 441 -c 160 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
SECT="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis"
for g in 64 14; do
  RX="attention_kernel<.int.64"; [ $g = 14 ] && RX="windowed_attention_kernel"
  timeout 300 ncu $SECT --metrics group:smsp__pcsamp_warp_stall_reasons,group:smsp__pcsamp_warp_stall_reasons_not_issued,dram__bytes_read.sum,dram__bytes_write.sum --import-source on \
    --clock-control none --kernel-name-base demangled -k "regex:$RX" -s 2 -c 1 -o gpurun_out/prof_attn${g}_${TAG} -f \
    python tools/time_attention.py > gpurun_out/ncu_attn${g}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_attn${g}_${TAG}.log
done
ls -la gpurun_out/ | grep ${TAG}
