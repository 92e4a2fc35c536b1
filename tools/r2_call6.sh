#!/bin/bash
# Round-2 call 6: CTA-level timeline (prologue / loop / epilogue) of a first-wave and a mid-grid CTA of the global attention.
mkdir -p gpurun_out
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_tl_cta.log 2>&1; sort gpurun_out/r2_tl_cta.log | head -90
timeout 100 python tools/time_attention.py 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
