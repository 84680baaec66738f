#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 500 -x -k "padded_row or fused_attention or layernorm_head"
TAILN=60 run python tools/ablate.py variants
