#!/bin/bash
# LayerNorm fold: op-level tests (api_ops rebuilt), then the per-GEMM cost of the fold with its parts switched off one by one
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu -s 2>&1 | grep -E "fold|passed|failed|Error|error|assert" | head -30
for v in default fdbg1 fdbg2 fdbg3 nopf sdbg1 sdbg2 sdbg3 default; do
  if [ "$v" = default ]; then L=""; else L="variants/lib_$v.so"; fi
  echo "=== $v"; ULLSAM_B200_LIB=$L timeout 200 python tools/time_gemm_fold.py 2>&1 | tail -9
done
