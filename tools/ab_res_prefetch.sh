cd /root/repo
for lib in variants/lib_nopf.so ullsam_b200/lib/libullsam_b200.so variants/lib_pf2.so; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 300 python tools/time_gemm.py 2>&1 | grep "residual\|768 K=3072 bf16 "
done
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu 2>&1 | tail -2
bash tools/bench_modes.sh "ULLSAM_B200_LIB=/root/repo/variants/lib_nopf.so" "ULLSAM_B200_LIB=/root/repo/ullsam_b200/lib/libullsam_b200.so" "ULLSAM_B200_LIB=/root/repo/variants/lib_pf2.so"
