#!/bin/bash
# last A/B of the round: residual L2 prefetch for short-K GEMMs on top of the register-level addend prefetch
cd /root/repo
for lib in ullsam_b200/lib/libullsam_b200.so variants/lib_pf2k.so; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 100 python tools/time_gemm.py 2>&1 | grep "residual"
done
ULLSAM_B200_LIB=/root/repo/variants/lib_pf2k.so timeout 200 python -m pytest tests/test_ops_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -1
