#!/bin/bash
# logistic-form GELU in lin1's epilogue + coalesced key-norm pre-pass: tests, A/B against the A&S erf build, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gemm_gpu.py tests/test_ops_attention_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "^gelu|rel-L2|passed|failed|Error|error|assert" | head -20
for v in default gelu_as default gelu_as; do
  if [ "$v" = default ]; then L=""; else L="variants/lib_$v.so"; fi
  echo "=== $v"; ULLSAM_B200_LIB=$L timeout 200 python tools/time_gemm_fold.py 2>&1 | grep "lin1"
done
python tools/time_attention.py 2>&1 | tail -2
for v in default gelu_as default; do
  if [ "$v" = default ]; then L=""; else L="variants/lib_$v.so"; fi
  ULLSAM_B200_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m_bench_$v.json 2> gpurun_out/r2m_bench_$v.err
  echo "=== bench $v"; python tools/bench_print.py gpurun_out/r2m_bench_$v.json 2>/dev/null | head -12 | grep -E "value|lin1|attn_global"
done
