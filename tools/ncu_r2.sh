#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU):  bash tools/ncu_r2.sh
#  (1) launch list of one ViT-B bench step and one ViT-L step (per-launch device time; cold-cache, serialised: compare shares)
#  (2) hardware-counter sections of one launch of: both attention kernels, the decoder's fused token kernel, the ViT-L qkv
#      GEMM (N = 3072, K = 1024) — no SASS patching (the --set full replay of the 168-register kernels does not launch once
#      patched, see profiles/README.md); dram bytes added explicitly.
TAG=${1:-r2a}
mkdir -p gpurun_out
export ULLSAM_PDL=0   # (per-launch times: programmatic dependent launch would let consecutive grids overlap)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_vit_b.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_${TAG}_vit_b.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}_vit_l.csv \
  python bench.py --model vit_l --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_${TAG}_vit_l.log 2>&1
SECT="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis"
MET="--metrics dram__bytes_read.sum,dram__bytes_write.sum"
cap() {  # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 ncu $SECT $MET --clock-control none --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 -o gpurun_out/prof_${name}_${TAG} -f "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_${name}_${TAG}.log
}
cap attn64 "attention_kernel<.int.64" 2 python tools/time_attention.py
cap attn14 "windowed_attention_kernel" 2 python tools/time_attention.py
cap dectok "dec_token_layer_kernel" 3 python tools/time_decoder.py 3
cap decfin "dec_token_final_kernel" 1 python tools/time_decoder.py 3
cap qkv_vitl "gemm_kernel<.int.256, .int.2," 6 python tools/time_gemm_vitl.py
# LayerNorm-folded GEMMs (ViT-B, batch 32): the consumer (qkv, epilogue 14) and the producer with K = 3072 (lin2, epilogue 12; the
# first two matches are the K = 768 launches of the proj shape, so skip 12 = 10 proj + 2 lin2 warm-ups)
cap qkv_fold "gemm_kernel<.int.256, .int.14," 4 python tools/time_gemm_fold.py
cap lin2_stats "gemm_kernel<.int.256, .int.12," 12 python tools/time_gemm_fold.py
ls -la gpurun_out/ | grep ${TAG}
