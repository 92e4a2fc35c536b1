#!/bin/bash
# A/B of GEMM epilogue variants (under gpurun): every variants/lib_*.so (tools/build_variant.sh) vs the in-tree library
cd /root/repo
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
LIBS="$(ls variants/lib_*.so 2>/dev/null) ullsam_b200/lib/libullsam_b200.so"
for lib in $LIBS; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 300 python tools/time_gemm.py 2>&1 | grep "residual\|qkv\|3072 K=768 bf16 + GELU"
done
ARGS=(); for lib in $LIBS; do ARGS+=("ULLSAM_B200_LIB=/root/repo/$lib"); done
bash tools/bench_modes.sh "${ARGS[@]}"
