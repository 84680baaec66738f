#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_q.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'launches', d['gpu_launches']);print(d['roofline']);print(d['whole_path']);print(d['cpu_baseline']);print(d['clocks'])"
timeout 1500 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_q_ref.json 2> gpurun_out/bench_q_ref.err; echo "ref exit $?"; tail -3 gpurun_out/bench_q_ref.err; cut -c1-600 gpurun_out/bench_q_ref.json
timeout 600 python tools/gemm_slope.py 2>&1 | tail -8
