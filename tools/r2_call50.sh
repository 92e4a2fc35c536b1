#!/bin/bash
# global attention, two threads per row with the exchange-free fast mode (variants/lib_att2.so) vs the default build
L=variants/lib_att2.so
ULLSAM_B200_LIB=$L timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu 2>&1 | tail -3
for r in 1 2; do
  echo -n "default: "; timeout 100 python tools/time_attention.py 2>&1 | grep "grid 64"
  echo -n "att2:    "; ULLSAM_B200_LIB=$L timeout 100 python tools/time_attention.py 2>&1 | grep "grid 64"
done
ULLSAM_B200_LIB=$L timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel-L2|passed|failed|rror" | head
