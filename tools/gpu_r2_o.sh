#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=40 run python -m pytest tests -m gpu -q --timeout 800
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_o.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'launches', d['gpu_launches']);print(d['roofline']);print(d['whole_path'])"
