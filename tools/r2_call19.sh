#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel-L2|IoU|passed|failed|Error|error|assert" | head -30
echo "=== bench"
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_vit_b.json 2> gpurun_out/r2e_bench_vit_b.err; tail -2 gpurun_out/r2e_bench_vit_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench_vit_b.json"))
print("value %.1f img/s  ms/step %.2f  e2e %s  step frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"], 1), d["roofline"]["frac"], d["gpu_launches"]))
print("   clocks", d["clocks"])
for k, v in d["roofline"]["kernels"].items():
    if k == "decoder_parts": continue
    print("   %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
PY
