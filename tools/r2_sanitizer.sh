#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / cluster protocols (SURVEY §5 "race detection"): memcheck on the
# op-level attention, GEMM and decoder tests (small shapes), racecheck on the cluster kernel of the decoder's token side.
mkdir -p gpurun_out
export ULLSAM_B200_LIB=variants/lib_sanitizer.so
S=/usr/local/cuda/bin/compute-sanitizer
echo "=== memcheck: attention (global G=1, windowed G=50) + determinism skipped"
timeout 1500 $S --tool memcheck --print-limit 10 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "test_global_attention or test_windowed_attention or layernorm or im2col" > gpurun_out/r2_memcheck_attention.log 2>&1
tail -4 gpurun_out/r2_memcheck_attention.log
echo "=== memcheck: GEMM epilogues (small shapes)"
timeout 1500 $S --tool memcheck --print-limit 10 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu -k "not 16384 and not 131072" > gpurun_out/r2_memcheck_gemm.log 2>&1
tail -4 gpurun_out/r2_memcheck_gemm.log
echo "=== memcheck: decoder modules (fused token kernels, cluster of 4)"
timeout 1500 $S --tool memcheck --print-limit 10 python -m pytest tests/test_modules_gpu.py -x -q -m gpu > gpurun_out/r2_memcheck_modules.log 2>&1
tail -4 gpurun_out/r2_memcheck_modules.log
echo "=== racecheck: decoder modules"
timeout 1500 $S --tool racecheck --print-limit 10 python -m pytest tests/test_modules_gpu.py -x -q -m gpu -k "transformer or decoder" > gpurun_out/r2_racecheck_modules.log 2>&1
tail -6 gpurun_out/r2_racecheck_modules.log
echo "=== initcheck: attention (global)"
timeout 900 $S --tool initcheck --print-limit 10 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "test_global_attention and 0.25" > gpurun_out/r2_initcheck_attention.log 2>&1
tail -4 gpurun_out/r2_initcheck_attention.log
