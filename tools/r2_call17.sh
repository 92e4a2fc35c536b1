#!/bin/bash
mkdir -p gpurun_out
for v in mA mB mC mD mE; do echo "--- parity $v"; ULLSAM_B200_LIB=variants/lib_$v.so timeout 300 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -k "global or (deterministic and 64)" 2>&1 | tail -1; done
bash tools/ab_variants.sh nola mA mB mC mD mE
