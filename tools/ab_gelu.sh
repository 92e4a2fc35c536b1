cd /root/repo
for lib in ullsam_b200/lib/libullsam_b200.so variants/lib_gelu0x1.so variants/lib_gelu0x5.so variants/lib_gelu0xF.so; do
  echo "=== $lib"; ULLSAM_B200_LIB=/root/repo/$lib timeout 300 python tools/time_gemm.py 2>&1 | grep "3072 K=768 bf16"
done
ULLSAM_B200_LIB=/root/repo/variants/lib_gelu0x5.so timeout 300 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu 2>&1 | tail -1
