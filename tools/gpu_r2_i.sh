#!/bin/bash
# BASELINE.json configs[2..4] at their per-GPU size on one B200 (the 8-GPU runs shard images; per-GPU work is identical)
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=12 run python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 800 -k "beam"
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'value %.1f e2e %.1f ms/step %.1f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
print(' whole_path', d['whole_path']); print(' roofline', d['roofline'])
for k,v in list(d['kernel_breakdown'].items())[:12]: print('  ',k,v)
" $1; }
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 32 --max-length 128 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 exit $?"; show gpurun_out/bench_cfg3.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 16 --max-length 128 --num-beams 4 --early-stopping > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"; show gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 64 --image-size 1024 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?"; show gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
