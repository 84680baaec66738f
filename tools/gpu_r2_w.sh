#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/decode_timeline.py 2>&1 | tail -40
