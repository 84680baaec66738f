#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=12 run python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "cta_pair"
TAILN=30 run python tools/ablate.py
bash tools/gpu_r2_m.sh
