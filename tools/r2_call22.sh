#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_modules_gpu.py tests/test_mllm_config5_gpu.py tests/test_prompt_encoder_gpu.py -x -q -m gpu -s 2>&1 | grep -E "IoU|iou|queries|keys|passed|failed|Error|error|assert" | head -40
echo "=== bench (fused decoder tokens)"
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2f_bench_vit_b.json 2> gpurun_out/r2f_bench_vit_b.err; tail -2 gpurun_out/r2f_bench_vit_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench_vit_b.json"))
print("value %.1f img/s  ms/step %.2f  e2e %s  step frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"], 1), d["roofline"]["frac"], d["gpu_launches"]))
for k, v in d["roofline"]["kernels"].items():
    if k == "decoder_parts":
        for kk, vv in v.items(): print("      %-22s %7.3f ms  %5.1f launches" % (kk, vv["ms_per_step"], vv["launches_per_step"]))
        continue
    print("   %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
PY
