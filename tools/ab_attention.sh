#!/bin/bash
# time the attention op with every variants/lib_*.so (and the default build)
echo "default:"; timeout 120 python tools/time_attention.py
for f in variants/lib_*.so; do echo "$f:"; ULLSAM_B200_LIB=$PWD/$f timeout 120 python tools/time_attention.py; done
