#!/bin/bash
# bash tools/bench_modes.sh "<ENV=val ...>" ...   one short bench per environment setting, breakdown printed
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "=== $envs"
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_mode$i.json 2> gpurun_out/bench_mode$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_mode$i.json"))
k = d["roofline"]["kernels"]
print("value %.1f img/s  ms/step %.2f | " % (d["value"], d["ms_per_step"]) + "  ".join("%s %.2f" % (n.replace("gemm_","").replace("mlp_",""), v["ms_per_step"]) for n, v in k.items() if n != "decoder_parts"))
PY
done
