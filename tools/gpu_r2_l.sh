#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=15 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 800 -x -k "two_halves or padded_row or fused_attention"
TAILN=15 run python -m pytest tests/test_gpu_decoder_parity.py -m gpu -q --timeout 800 -x -k "928 or real_cache"
TAILN=40 run python tools/ablate.py
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_l.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'launches', d['gpu_launches']);print(d['roofline'])"
