#!/bin/bash
# Round-2 call 4: (a) can ONE softmax warp per sub-partition saturate the MUFU?  (timeline builds, query tile 1 idle in the
# second one)  (b) where the config 5 fixture's mask error comes from.
mkdir -p gpurun_out
echo "=== timeline, both query tiles"
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 2>&1 | tail -4
echo "=== timeline, query tile 1 idle (solo)"
ULLSAM_B200_LIB=variants/lib_tl_solo.so timeout 120 python tools/time_attention_once.py 2>&1 | tail -4
echo "=== config 5 error decomposition"
timeout 600 python tools/exp_config5_error.py 2>&1 | tail -40
