#!/bin/bash
# round 2, call A: bring-up of the fused attention kernel / LayerNorm heads, each risky variant in its own process
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 500 -x -k "padded_row or fused_attention or teacher_forced"
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 500 -x -k "layernorm_head"
TAILN=30 run python -m pytest tests -m gpu -q --timeout 500
TAILN=40 run python tools/ablate.py variants
