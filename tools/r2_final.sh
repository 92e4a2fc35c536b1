#!/bin/bash
# Final-build record of round 2: launch list (ncu), the three recorded workloads, the reference arm.
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 300 --csv --log-file gpurun_out/launches_r2c_vit_b.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_r2c_vit_b.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench_vit_b.json 2> gpurun_out/r2p_bench_vit_b.err
python tools/bench_print.py gpurun_out/r2p_bench_vit_b.json > gpurun_out/r2p_print.txt 2>&1; cat gpurun_out/r2p_print.txt
timeout 600 python bench.py --model vit_l --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench_vit_l.json 2> gpurun_out/r2p_bench_vit_l.err
python tools/bench_print.py gpurun_out/r2p_bench_vit_l.json > gpurun_out/r2p_print_l.txt 2>&1; head -12 gpurun_out/r2p_print_l.txt
timeout 600 python bench.py --workload llm_dense --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench_llm_dense.json 2> gpurun_out/r2p_bench_llm_dense.err
python tools/bench_print.py gpurun_out/r2p_bench_llm_dense.json 2>&1 | head -2
