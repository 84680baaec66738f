#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=12 run python -m pytest tests -m gpu -q --timeout 800
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.1f e2e %.1f ms/step %.1f launches %d frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['whole_path']['frac']))
print('  roofline', {k: d['roofline'][k] for k in ('kernel','achieved','frac','traffic','avg_ms')} if d['roofline'] else None, 'cpu', d.get('cpu_baseline',{}).get('value'))
" $1; }
timeout 900 python bench.py > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo "bench exit $?"; show gpurun_out/bench_t.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 32 --max-length 128 > gpurun_out/bench_t_cfg3.json 2> gpurun_out/bench_t_cfg3.err; echo "cfg3 exit $?"; show gpurun_out/bench_t_cfg3.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 16 --max-length 128 --num-beams 4 --early-stopping > gpurun_out/bench_t_cfg4.json 2> gpurun_out/bench_t_cfg4.err; echo "cfg4 exit $?"; show gpurun_out/bench_t_cfg4.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 64 --image-size 1024 > gpurun_out/bench_t_cfg5.json 2> gpurun_out/bench_t_cfg5.err; echo "cfg5 exit $?"; show gpurun_out/bench_t_cfg5.json
timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch 16 --max-length 300 --num-beams 4 --early-stopping > gpurun_out/bench_t_script.json 2> gpurun_out/bench_t_script.err; echo "script exit $?"; show gpurun_out/bench_t_script.json
