#!/bin/bash
# lin1's GELU epilogue now that the GELU is 9 instructions: 16 epilogue warps (default) vs 8 rolled vs 8 with pipelined TMEM loads
for r in 1 2; do
echo "ew16 (default)  "; python tools/time_gemm_fold.py 2>&1 | grep "lin1 gelu  "
echo "ew8 rolled      "; ULLSAM_GEMM_EW16=0 python tools/time_gemm_fold.py 2>&1 | grep "lin1 gelu  "
echo "ew8 pipelined   "; ULLSAM_GEMM_EW16=0 ULLSAM_B200_LIB=variants/lib_gelu_pipe.so python tools/time_gemm_fold.py 2>&1 | grep "lin1 gelu  "
done
python tools/time_gemm.py 2>&1 | grep "N=3072"
