#!/bin/bash
# ViT-B launch list again (the first 350 launches of the process are the weight upload: skip them), then the other
# recorded workloads: ViT-L (config 4) and the LLM-conditioned dense prompt (config 3), and the default line once more.
mkdir -p gpurun_out
ULLSAM_PDL=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 300 --csv --log-file gpurun_out/launches_r2b_vit_b.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_r2b_vit_b.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench_vit_b.json 2> gpurun_out/r2l_bench_vit_b.err
python tools/bench_print.py gpurun_out/r2l_bench_vit_b.json | head -30
timeout 600 python bench.py --model vit_l --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_vit_l.json 2> gpurun_out/r2l_bench_vit_l.err
python tools/bench_print.py gpurun_out/r2l_bench_vit_l.json | head -14
timeout 600 python bench.py --workload llm_dense --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_llm_dense.json 2> gpurun_out/r2l_bench_llm_dense.err
python tools/bench_print.py gpurun_out/r2l_bench_llm_dense.json | head -3
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err; cat gpurun_out/r2l_bench_reference.json | cut -c1-600
# racecheck of the decoder again with every hazard printed, to classify them (the first run showed 10 of 1185, all on the
# reserved shared-memory words below 0x400 that tcgen05.alloc / barrier.cluster use)
ULLSAM_B200_LIB=variants/lib_sanitizer.so timeout 500 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report all --print-limit 3000 python tools/time_decoder.py 1 2>&1 | grep -v "Host Frame\|Saved host\|^========= $" > gpurun_out/r2_racecheck_decoder_all.log; tail -1 gpurun_out/r2_racecheck_decoder_all.log
