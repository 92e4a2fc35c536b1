#!/bin/bash
# compute-sanitizer follow-up (time-boxed): the kernels the first run never reached because the global attention kernel
# stalled under memcheck - windowed attention, the decoder (fused cluster-of-4 token kernels, image-side GEMM epilogues).
mkdir -p gpurun_out
export ULLSAM_B200_LIB=variants/lib_sanitizer.so
S=/usr/local/cuda/bin/compute-sanitizer
echo "=== memcheck: decoder alone (32 prompts, 1 pass)"
timeout 400 $S --tool memcheck --print-limit 10 python tools/time_decoder.py 1 > gpurun_out/r2_memcheck_decoder.log 2>&1; tail -3 gpurun_out/r2_memcheck_decoder.log
echo "=== racecheck: decoder alone"
timeout 500 $S --tool racecheck --racecheck-report all --print-limit 10 python tools/time_decoder.py 1 > gpurun_out/r2_racecheck_decoder.log 2>&1; tail -3 gpurun_out/r2_racecheck_decoder.log
echo "=== memcheck: windowed attention op tests"
timeout 400 $S --tool memcheck --print-limit 10 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "windowed" > gpurun_out/r2_memcheck_windowed.log 2>&1; tail -4 gpurun_out/r2_memcheck_windowed.log
echo "=== memcheck: folded-LayerNorm GEMM epilogues"
timeout 300 $S --tool memcheck --print-limit 10 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu -k "fold or statistics" > gpurun_out/r2_memcheck_gemm_fold.log 2>&1; tail -4 gpurun_out/r2_memcheck_gemm_fold.log
echo "=== memcheck: global attention op test, G=1 (stalled in the first run: the P hand-over lap, fixed since)"
timeout 600 $S --tool memcheck --print-limit 10 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "test_global_attention and 0.25" > gpurun_out/r2_memcheck_global2.log 2>&1; tail -4 gpurun_out/r2_memcheck_global2.log
