#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=25 run python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "cta_pair"
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 300 -x -k "cta_pair"
TAILN=30 run python tools/ablate.py
