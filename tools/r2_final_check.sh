#!/bin/bash
# last validation of the committed build, the way the driver runs it: GPU suite, smoke, both bench arms; then the bench twice more (stability)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2r_bench_$i.json 2> gpurun_out/r2r_bench_$i.err
  python tools/bench_print.py gpurun_out/r2r_bench_$i.json 2>/dev/null | head -1; grep -c "mbarrier timeout" gpurun_out/r2r_bench_$i.json; tail -1 gpurun_out/r2r_bench_$i.err
done
