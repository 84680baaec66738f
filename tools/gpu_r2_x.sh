#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -4
timeout 600 python tools/ablate.py 2>&1 | tail -14
