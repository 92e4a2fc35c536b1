#!/bin/bash
# Round-2 third GPU call: the 384-row global attention (attention_g3.cu) on hardware.
mkdir -p gpurun_out
echo "=== global attention tests, new kernel (default)"
timeout 600 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -s 2>&1 | grep -E "global|windowed|determin|passed|failed|Error|error|timeout" | head -30
echo "=== timing: new (default) vs old (ULLSAM_GLOBAL_ATT=2)"
timeout 120 python tools/time_attention.py 2>&1 | tail -2
ULLSAM_GLOBAL_ATT=2 timeout 120 python tools/time_attention.py 2>&1 | tail -2
echo "=== model parity + config 5"
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_mllm_config5_gpu.py -x -q -m gpu -s 2>&1 | grep -E "rel|IoU|iou|passed|failed|Error|error" | head -30
echo "=== bench"
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2c_bench_vit_b.json 2> gpurun_out/r2c_bench_vit_b.err
tail -3 gpurun_out/r2c_bench_vit_b.err
python - <<'PY'
import json
for f in ("r2c_bench_vit_b",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.1f img/s  ms/step %.2f  e2e %s  step frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"], 1), d["roofline"]["frac"], d["gpu_launches"]))
    print("   clocks", d["clocks"])
    for k, v in d["roofline"]["kernels"].items():
        if k == "decoder_parts": continue
        print("   %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
PY
echo "=== rest of the GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -8
