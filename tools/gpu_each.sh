#!/bin/bash
# Run every GPU test function of the given files in its own process (a trapping kernel poisons the
# CUDA context of its process only) and collect the output under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/gpu_each.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for f in "$@"; do
  ids=$(python -m pytest "$f" -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
  for id in $ids; do
    echo "=== $id" >> $LOG
    timeout 300 python -m pytest "$id" -m gpu -q -rA -x 2>&1 | grep -vE "^$" | tail -40 >> $LOG
  done
done
tail -150 $LOG
