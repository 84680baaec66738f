#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/decode_timeline.py 2>&1 | tail -24
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -x -q -k "cta_pair or split_k or gemm or bit_identical or golden" 2>&1 | tail -4
timeout 600 python tools/ablate.py 2>&1 | tail -12
