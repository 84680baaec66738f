#!/bin/bash
mkdir -p gpurun_out
for t in 8 16 32 64; do echo "=== T=$t"; T=$t timeout 600 python tools/decode_timeline.py 2>&1 | grep -A3 "^layer 12" ; done
