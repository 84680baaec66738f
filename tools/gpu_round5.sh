#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_diag.py timeline > gpurun_out/timeline5.log 2>&1
echo "timeline exit $?"; grep -v interleaved=1 gpurun_out/timeline5.log | tail -12
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -x > gpurun_out/test_gpu5.log 2>&1
echo "tests exit $?"; tail -8 gpurun_out/test_gpu5.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
echo "bench exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench5.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step']);print(d['roofline']);[print(k,v) for k,v in list(d['kernel_breakdown'].items())[:14]]"
