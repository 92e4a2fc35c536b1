#!/bin/bash
timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "windowed or (deterministic and 14)" 2>&1 | tail -2
bash tools/ab_variants.sh prev default
ULLSAM_B200_LIB=variants/lib_wintl.so python tools/time_attention_win_once.py 2>&1 | tail -2
