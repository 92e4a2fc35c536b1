#!/bin/bash
# LayerNorm folded into the GEMMs: op-level tests, model parity, then the bench with the fold on (default) and off.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu -s 2>&1 | grep -E "fold|stats|row sums|passed|failed|Error|error|assert" | head -40
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_mllm_config5_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel-L2|IoU|transformer:|passed|failed|Error|error|assert" | head -30
for f in 1 0; do
  ULLSAM_LN_FOLD=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i_bench_fold$f.json 2> gpurun_out/r2i_bench_fold$f.err
  echo "=== ULLSAM_LN_FOLD=$f"; python tools/bench_print.py gpurun_out/r2i_bench_fold$f.json 2>/dev/null | head -12; tail -3 gpurun_out/r2i_bench_fold$f.err
done
ULLSAM_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i_bench_fold1_pdl0.json 2> gpurun_out/r2i_bench_fold1_pdl0.err
echo "=== fold on, PDL off"; python tools/bench_print.py gpurun_out/r2i_bench_fold1_pdl0.json 2>/dev/null | head -3
