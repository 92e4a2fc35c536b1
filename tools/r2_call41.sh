#!/bin/bash
# LayerNorm fold, same-box A/B of the masks: 0 = LayerNorm kernels, 1 = norm1 folded (default), 3 = both folded
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | head -12
for f in 1 0 3 1 0; do
  ULLSAM_LN_FOLD=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2j_bench_fold$f.json 2> gpurun_out/r2j_bench_fold$f.err
  echo "=== ULLSAM_LN_FOLD=$f"; python tools/bench_print.py gpurun_out/r2j_bench_fold$f.json 2>/dev/null | head -12; tail -2 gpurun_out/r2j_bench_fold$f.err
done
