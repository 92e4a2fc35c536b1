#!/bin/bash
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_modules_gpu.py tests/test_mllm_config5_gpu.py tests/test_prompt_encoder_gpu.py -x -q -m gpu -s 2>&1 | grep -E "Ns=|transformer:|all 12|passed|failed|Error|error|assert" | head -20
python tools/time_decoder.py 5 | tail -1; ULLSAM_DEC_FUSED=0 python tools/time_decoder.py 5 | tail -1
