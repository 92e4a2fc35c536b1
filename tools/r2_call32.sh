#!/bin/bash
for round in 1 2; do for v in default gp1 gp5 gpA gp7 gpF; do
  if [ "$v" = default ]; then L=""; else L="variants/lib_$v.so"; fi
  printf "%-8s " $v; ULLSAM_B200_LIB=$L timeout 100 python tools/time_gemm.py 2>&1 | grep "N=3072 K=768 bf16 + GELU"
done; done
