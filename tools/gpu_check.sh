#!/bin/bash
# Standard GPU iteration: selected tests (each file in its own process group) then a short bench.
# Usage (under gpurun): bash tools/gpu_check.sh [tag] [test files...]
TAG=${1:-iter}; shift
FILES=${@:-tests/test_ops_attention_gpu.py tests/test_model_gpu.py}
mkdir -p gpurun_out
for f in $FILES; do
  timeout 600 python -m pytest $f -m gpu -q -rA -x 2>&1 | grep -E "err|IoU|PASS|FAIL|Error|error|assert|passed|failed|timeout" | head -40
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.1f img/s  ms/step %.2f  e2e %s  step frac(sustained) %.3f  gemm TF/s %.0f" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"],1), d["roofline"]["step"]["frac_of_sustained"], d["roofline"]["achieved"]))
for k, v in d["roofline"]["kernels"].items():
    if k == "decoder_parts":
        for kk, vv in v.items(): print("      %-22s %7.3f ms/step  %5.1f launches" % (kk, vv["ms_per_step"], vv["launches_per_step"]))
        continue
    print("  %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
print("clocks", d["clocks"])
PY
