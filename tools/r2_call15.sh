#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_variants.sh ko4 ko20 ko28 ko31 2>&1 | head -4
for v in tl_ko4; do
  echo "=== $v"
  ULLSAM_B200_LIB=variants/lib_$v.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_sched_$v.log 2>&1
  grep "^  at" gpurun_out/r2_sched_$v.log | sort -k2 -n | head -24
  grep "per tile\|mma warp . per\|cta (5" gpurun_out/r2_sched_$v.log
done
