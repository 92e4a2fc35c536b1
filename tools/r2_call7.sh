#!/bin/bash
# Round-2 call 7: branch-free prologue of the global attention: parity, timing, CTA timeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -s 2>&1 | grep -E "global|determin|passed|failed|Error|error|timeout" | head -20
timeout 100 python tools/time_attention.py 2>&1 | tail -2
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_tl_cta.log 2>&1; sort gpurun_out/r2_tl_cta.log | grep -v "tile" | head
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel|IoU|passed|failed|Error|error" | head -20
