"""Pretty-print a bench.py JSON line: python tools/bench_print.py gpurun_out/x.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f img/s  ms/step %.2f  e2e %s  step frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"], 1), d["roofline"]["frac"], d["gpu_launches"]))
print("   clocks", d["clocks"])
for k, v in d["roofline"]["kernels"].items():
    if k == "decoder_parts":
        for kk, vv in v.items(): print("      %-22s %7.3f ms  %5.1f launches" % (kk, vv["ms_per_step"], vv["launches_per_step"]))
        continue
    print("   %-16s %7.3f ms/step  %5.1f launches  %s TF/s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"] and round(v["tflops"])))
for k in ("latency_b1", "gpu_eager_bf16", "cpu_baseline"):
    if d.get(k): print("  ", k, d[k])
