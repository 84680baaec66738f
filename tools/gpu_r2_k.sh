#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -${TAILN:-8}; echo "exit ${PIPESTATUS[0]}"; }
TAILN=25 run python -m pytest tests/test_gpu_path.py tests/test_gpu_kernels.py -m gpu -q --timeout 800 -k "beam or roi_align or end_to_end or masks or layernorm_head"
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'value %.1f e2e %.1f ms/step %.1f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
for k,v in list(d['kernel_breakdown'].items())[:10]: print('  ',k,v)
" $1; }
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 16 --max-length 128 --num-beams 4 --early-stopping > gpurun_out/bench_cfg4b.json 2> gpurun_out/bench_cfg4b.err; echo "cfg4 exit $?"; show gpurun_out/bench_cfg4b.json
timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch 16 --max-length 300 --num-beams 4 --early-stopping > gpurun_out/bench_script_defaults.json 2> gpurun_out/bench_script_defaults.err; echo "T300 exit $?"; show gpurun_out/bench_script_defaults.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench exit $?"; show gpurun_out/bench_k.json
TAILN=40 run python tools/ablate.py variants
