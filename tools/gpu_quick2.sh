#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_kernels.py -m gpu -q --timeout 500 -x -k "cta_pair or fused" 2>&1 | tail -3
timeout 600 python tools/ablate.py 2>&1 | grep -E "full step|epi_|without mlp"
T=48 timeout 600 python tools/decode_timeline.py 2>&1 | tail -3
