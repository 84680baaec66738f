#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=8 run python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "cta_pair"
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.1f e2e %.1f ms/step %.1f launches %d frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['whole_path']['frac']))
" $1; }
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 16 --max-length 128 --num-beams 4 --early-stopping > gpurun_out/bench_u_cfg4.json 2> gpurun_out/bench_u_cfg4.err; echo "cfg4 exit $?"; show gpurun_out/bench_u_cfg4.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 64 --image-size 1024 > gpurun_out/bench_u_cfg5.json 2> gpurun_out/bench_u_cfg5.err; echo "cfg5 exit $?"; show gpurun_out/bench_u_cfg5.json
timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch 16 --max-length 300 --num-beams 4 --early-stopping > gpurun_out/bench_u_script.json 2> gpurun_out/bench_u_script.err; echo "script exit $?"; show gpurun_out/bench_u_script.json
TAILN=20 run python tools/ablate_only.py
