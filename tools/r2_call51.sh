#!/bin/bash
# key-norm maximum on the QKV epilogue (no pre-pass launch) + CTA-pair patch GEMM: full GPU suite, smoke, then same-box A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in 1 0 1 0; do
  ULLSAM_KMAX_EPI=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench_kmax$v.json 2> gpurun_out/r2o_bench_kmax$v.err
  echo "=== ULLSAM_KMAX_EPI=$v"; python tools/bench_print.py gpurun_out/r2o_bench_kmax$v.json 2>/dev/null | grep -E "value|patch|qkv|attn_global"; tail -1 gpurun_out/r2o_bench_kmax$v.err
done
