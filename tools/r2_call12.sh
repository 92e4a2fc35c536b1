#!/bin/bash
# Round-2 call 12: deeper K/V rings (5 / 4 stages, extra stages in the dead rel_w scratch): parity, A/B timing, schedule trace.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -s 2>&1 | grep -E "global|determin|passed|failed|Error|error|timeout" | head -20
bash tools/ab_variants.sh nola default nola2
for v in timeline; do
  echo "=== $v"
  ULLSAM_B200_LIB=variants/lib_$v.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_sched_$v.log 2>&1
  grep "^  at" gpurun_out/r2_sched_$v.log | sort -k2 -n | head -40
  grep "per tile\|mma warp per\|cta (5" gpurun_out/r2_sched_$v.log
done
