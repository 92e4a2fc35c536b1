#!/bin/bash
# After the hand-over fix: attention tests (determinism included), the model tests, attention timing, the default bench,
# then the configuration that had hung (PDL on, LayerNorm kernels) three times over.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_attention_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/time_attention.py 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python tools/bench_print.py gpurun_out/r2k_bench.json 2>/dev/null | head -12; tail -2 gpurun_out/r2k_bench.err
for i in 1 2 3; do
  ULLSAM_PDL=1 ULLSAM_LN_FOLD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_pdl_$i.json 2> gpurun_out/r2k_bench_pdl_$i.err
  echo "=== PDL on, LN kernels, run $i"; python tools/bench_print.py gpurun_out/r2k_bench_pdl_$i.json 2>/dev/null | head -1; grep -c "mbarrier timeout" gpurun_out/r2k_bench_pdl_$i.json; tail -1 gpurun_out/r2k_bench_pdl_$i.err
done
