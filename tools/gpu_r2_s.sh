#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=15 run python -m pytest tests/test_gpu_path.py tests/test_gpu_decoder_parity.py -m gpu -q --timeout 800 -x -k "fused_attention or teacher_forced or padded_row or out_of_memory or cta_pair"
TAILN=30 run python tools/ablate.py
