#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_y_bench.json 2> gpurun_out/r2_y_bench.err; tail -c 600 gpurun_out/r2_y_bench.json | head -c 600; echo
for w in 1 2; do
  RGRG_OPTS=gemm_2cta_waves=$w timeout 900 python bench.py --no-cpu-baseline --batch 16 --num-beams 4 --early-stopping --max-length 128 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 waves=$w', d['value'], d['ms_per_step'])"
done
RGRG_OPTS=gemm_2cta_waves=2 timeout 900 python bench.py --no-cpu-baseline --batch 64 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b64 waves=2', d['value'], d['ms_per_step'])"
timeout 900 python bench.py --no-cpu-baseline --batch 64 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b64 waves=1', d['value'], d['ms_per_step'])"
