#!/bin/bash
# Round-2 first GPU call: experimental windowed kernel on hardware, baseline bench of this pod, ViT-L, sanitizer.
mkdir -p gpurun_out
echo "=== win4 (ULLSAM_WIN_ATT=3) parity, child process"
ULLSAM_WIN_ATT=3 timeout 300 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "windowed or (deterministic and 14)" 2>&1 | tail -15
echo "=== timing v2"
timeout 120 python tools/time_attention.py 2>&1 | tail -3
echo "=== timing v3"
ULLSAM_WIN_ATT=3 timeout 120 python tools/time_attention.py 2>&1 | tail -3
echo "=== bench vit_b"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_vit_b.json 2> gpurun_out/r2a_bench_vit_b.err
tail -2 gpurun_out/r2a_bench_vit_b.err; cut -c1-600 gpurun_out/r2a_bench_vit_b.json
echo "=== bench vit_l"
timeout 900 python bench.py --model vit_l --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_vit_l.json 2> gpurun_out/r2a_bench_vit_l.err
tail -2 gpurun_out/r2a_bench_vit_l.err; cut -c1-600 gpurun_out/r2a_bench_vit_l.json
echo "=== compute-sanitizer memcheck (op-level attention + gemm tests, bounded)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "windowed_attention or global_attention and not moving" > gpurun_out/r2a_memcheck_attention.log 2>&1
tail -8 gpurun_out/r2a_memcheck_attention.log
