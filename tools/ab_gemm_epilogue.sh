#!/bin/bash
# A/B of GEMM epilogue variants (under gpurun): variants/lib_<tag>.so built with tools/build_variant.sh vs the in-tree library
cd /root/repo
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
for lib in variants/lib_base.so variants/lib_relaxed.so ullsam_b200/lib/libullsam_b200.so variants/lib_widepipe.so; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 300 python tools/time_gemm.py 2>&1 | grep "residual\|qkv\|3072 K=768 bf16 + GELU"
done
bash tools/bench_modes.sh "ULLSAM_B200_LIB=/root/repo/variants/lib_base.so" "ULLSAM_B200_LIB=/root/repo/variants/lib_relaxed.so" "ULLSAM_B200_LIB=/root/repo/ullsam_b200/lib/libullsam_b200.so" "ULLSAM_B200_LIB=/root/repo/variants/lib_widepipe.so"
