TAG=r1f
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
cap() { ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/prof_$1_${TAG} -f $BENCH > gpurun_out/ncu_$1_${TAG}.log 2>&1; }
cap gemm_lin2 "gemm_kernel<.int.256, .int.1, .int.2" 73
cap gemm_lin1 "gemm_kernel<.int.256, .int.10, .int.2" 36
cap gemm_qkv "gemm_kernel<.int.256, .int.2, .int.2" 36
ls -la gpurun_out | grep prof_gemm
