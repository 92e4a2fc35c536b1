#!/bin/bash
# A/B/A timing of attention variants inside ONE gpurun call (box-to-box clocks differ by a few %): usage ab_variants.sh v1 v2 ...
# each name is variants/lib_<name>.so ("default" = the in-tree library); two rounds, interleaved.
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then L=""; else L="variants/lib_$v.so"; fi
    printf "%-10s " "$v"; ULLSAM_B200_LIB=$L timeout 100 python tools/time_attention.py 2>&1 | tail -2 | tr '\n' ' '; echo
  done
done
