#!/bin/bash
# quick iteration: gpu tests + bench without the CPU baseline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -x > gpurun_out/test_quick.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/test_quick.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'launches', d['gpu_launches']);print(d['roofline']);[print(k,v) for k,v in list(d['kernel_breakdown'].items())[:12]]"
tail -3 gpurun_out/bench_quick.err
