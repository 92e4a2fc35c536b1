#!/bin/bash
# PDL validation: full GPU suite with programmatic dependent launch on (default), then the bench with it on and off.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_pdl1.json 2> gpurun_out/r2h_bench_pdl1.err
python tools/bench_print.py gpurun_out/r2h_bench_pdl1.json 2>/dev/null | head -40
ULLSAM_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_pdl0.json 2> gpurun_out/r2h_bench_pdl0.err
python tools/bench_print.py gpurun_out/r2h_bench_pdl0.json 2>/dev/null | head -40
