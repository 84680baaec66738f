#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -5
for o in 1 0; do
RGRG_OPTS=epi_tma=$o timeout 600 python bench.py --no-cpu-baseline --steps 4 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); kb=d['kernel_breakdown']
print('epi_tma=$o value %.2f ms/step %.2f' % (d['value'], d['ms_per_step']), {k: kb[k]['ms'] for k in ('conv1x1','conv3x3','fc6','fc7','mlp_c_fc') if k in kb})"
done
