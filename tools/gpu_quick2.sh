#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 500 -x 2>&1 | tail -4
timeout 600 python tools/ablate.py 2>&1 | grep -E "full step|without mlp_c_fc"
T=48 timeout 600 python tools/decode_timeline.py 2>&1 | tail -3
