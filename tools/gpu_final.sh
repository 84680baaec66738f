#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/test_final.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/test_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches']);print(d['roofline']);print(d['cpu_baseline']);print(d['clocks'])"
