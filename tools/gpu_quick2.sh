#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_decoder_parity.py -m gpu -q --timeout 500 -x 2>&1 | tail -5
timeout 600 python tools/ablate.py 2>&1 | grep -E "full step|flags"
T=48 timeout 600 python tools/decode_timeline.py 2>&1 | tail -22
