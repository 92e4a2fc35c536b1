#!/bin/bash
# racecheck on the two attention kernels (op tests, smallest shapes), every report printed for classification
mkdir -p gpurun_out
export ULLSAM_B200_LIB=variants/lib_sanitizer.so
S=/usr/local/cuda/bin/compute-sanitizer
timeout 500 $S --tool racecheck --racecheck-report all --print-limit 3000 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "test_global_attention and 0.25" 2>&1 | grep -v "Host Frame\|Saved host\|^========= $" > gpurun_out/r2_racecheck_global.log; tail -3 gpurun_out/r2_racecheck_global.log
timeout 500 $S --tool racecheck --racecheck-report all --print-limit 3000 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "windowed" 2>&1 | grep -v "Host Frame\|Saved host\|^========= $" > gpurun_out/r2_racecheck_windowed.log; tail -3 gpurun_out/r2_racecheck_windowed.log
