cd /root/repo
for lib in $(ls variants/lib_*.so 2>/dev/null) ullsam_b200/lib/libullsam_b200.so; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 300 python tools/time_gemm.py 2>&1 | grep "residual\|768 K=3072 bf16 "
done
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu 2>&1 | tail -2
ARGS=(); for lib in $(ls variants/lib_*.so 2>/dev/null) ullsam_b200/lib/libullsam_b200.so; do ARGS+=("ULLSAM_B200_LIB=/root/repo/$lib"); done; bash tools/bench_modes.sh "${ARGS[@]}"
