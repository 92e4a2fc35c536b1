#!/bin/bash
# Round-2 second GPU call: full GPU suite on the refreshed tree, then the new bench lines.
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15
echo "=== config 5 / prompt encoder prints"
timeout 600 python -m pytest tests/test_mllm_config5_gpu.py tests/test_prompt_encoder_gpu.py -m gpu -q -s 2>&1 | grep -E "config 5|multi|single|InternVLSAMModel|mask-prompt|passed|failed" | head -20
echo "=== bench vit_b (default line, all extras)"
timeout 900 python bench.py > gpurun_out/r2b_bench_vit_b.json 2> gpurun_out/r2b_bench_vit_b.err
tail -3 gpurun_out/r2b_bench_vit_b.err
echo "=== bench vit_b llm_dense"
timeout 600 python bench.py --workload llm_dense --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_vit_b_llm_dense.json 2> gpurun_out/r2b_bench_vit_b_llm_dense.err
tail -3 gpurun_out/r2b_bench_vit_b_llm_dense.err
echo "=== bench vit_l"
timeout 900 python bench.py --model vit_l --steps 5 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_vit_l.json 2> gpurun_out/r2b_bench_vit_l.err
tail -3 gpurun_out/r2b_bench_vit_l.err
python - <<'PY'
import json
for f in ("r2b_bench_vit_b", "r2b_bench_vit_b_llm_dense", "r2b_bench_vit_l"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.1f img/s  ms/step %.2f  e2e %s  c_abi %s  step frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"], 1), d["e2e"] and d["e2e"].get("c_abi_host", {}).get("value"), d["roofline"]["frac"], d["gpu_launches"]))
    print("   latency", d.get("latency_b1"), "\n   eager", d.get("gpu_eager_bf16"), "\n   cpu", d.get("cpu_baseline"), "\n   clocks", d["clocks"])
    for k, v in d["roofline"]["kernels"].items():
        if k == "decoder_parts": continue
        print("   %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
PY
echo "=== attention timeline (variant build)"
ULLSAM_B200_LIB=variants/lib_timeline.so timeout 120 python tools/time_attention_once.py 2>&1 | tail -8
