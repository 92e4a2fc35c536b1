#!/bin/bash
# proj residual add as TMA reduce-add stores (EPI_RES_RED): op test, model parity, then same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu -k "reduce_add or residual" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel-L2|passed|failed|rror" | head -8
for v in 1 0 1 0; do
  ULLSAM_PROJ_RED=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_red$v.json 2> gpurun_out/r2s_bench_red$v.err
  echo "=== ULLSAM_PROJ_RED=$v"; python tools/bench_print.py gpurun_out/r2s_bench_red$v.json 2>/dev/null | grep -E "value|proj|layernorm"; tail -1 gpurun_out/r2s_bench_red$v.err
done
