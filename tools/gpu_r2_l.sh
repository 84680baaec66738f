#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 800 -x -k "layernorm_head"
TAILN=40 run python tools/ablate.py variants
