#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
python - <<'PY'
import numpy as np, torch
from rgrg_b200 import Engine
e = Engine(0)
for cs in (2, 4, 8, 16):
    print("max co-resident clusters of %d (216 KB smem GEMM CTA):" % cs, int(e.debug_read("max_clusters_%d" % cs, (1,), np.int32)[0]))
PY
TAILN=40 run python -m pytest tests -m gpu -q --timeout 800 -x
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_e.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'launches', d['gpu_launches']);print(d['roofline']);[print(k,v) for k,v in list(d['kernel_breakdown'].items())[:14]]"
tail -3 gpurun_out/bench_e.err
