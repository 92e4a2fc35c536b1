#!/bin/bash
# Round-2 call 11: absolute-time schedule of one mid-grid CTA of the global attention (softmax warps of both query tiles + MMA warp).
mkdir -p gpurun_out
for v in timeline tl_nola; do
  echo "=== $v"
  ULLSAM_B200_LIB=variants/lib_$v.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_sched_$v.log 2>&1
  grep "^  at" gpurun_out/r2_sched_$v.log | sort -k2 -n | head -60
  grep "per tile\|mma warp per" gpurun_out/r2_sched_$v.log
done
