#!/bin/bash
timeout 900 python -m pytest tests/test_ops_gemm_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_modules_gpu.py tests/test_mllm_config5_gpu.py -x -q -m gpu -s 2>&1 | grep -E "Ns=|transformer:|all 12|passed|failed|Error|error|assert" | head -12
echo "prev:"; ULLSAM_B200_LIB=variants/lib_prev.so python tools/time_decoder.py 5 | tail -1
echo "new:"; python tools/time_decoder.py 5 | tail -1
echo "prev:"; ULLSAM_B200_LIB=variants/lib_prev.so python tools/time_decoder.py 5 | tail -1
echo "new:"; python tools/time_decoder.py 5 | tail -1
