#!/bin/bash
# Round-2 call 8: cross-tile software pipelining of the global attention's softmax: parity, timing, variants, timeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -s 2>&1 | grep -E "global|determin|passed|failed|Error|error|timeout" | head -20
echo "--- default"; timeout 100 python tools/time_attention.py 2>&1 | tail -2
for v in p32 poly1 poly2; do echo "--- $v"; ULLSAM_B200_LIB=variants/lib_$v.so timeout 100 python tools/time_attention.py 2>&1 | tail -2 | head -1; done
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 1 > gpurun_out/r2_tl_cta.log 2>&1; sort gpurun_out/r2_tl_cta.log | grep -v "tile  [0-9]\|tile [0-9]" | head; grep "qt 0 tile" gpurun_out/r2_tl_cta.log | sort | head -12
