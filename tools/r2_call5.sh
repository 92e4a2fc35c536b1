#!/bin/bash
# Round-2 call 5: per-tile wait trace of the global attention (timeline builds), with and without query tile 1 working.
mkdir -p gpurun_out
echo "=== timeline, both query tiles"
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_tl_both.log 2>&1; sort gpurun_out/r2_tl_both.log | head -80
echo "=== timeline, query tile 1 idle (solo)"
ULLSAM_B200_LIB=variants/lib_tl_solo.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_tl_solo.log 2>&1; sort gpurun_out/r2_tl_solo.log | head -80
