#!/bin/bash
# last check of the round: GPU suite, conv timings with the TMA-store epilogue, memcheck over the whole path (smoke)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_last.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1]); kb=d['kernel_breakdown']
print('value %.2f e2e %.2f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: kb[k]['ms'] for k in ('conv1x1','conv3x3','fc6','rpn_conv') if k in kb}, d['clocks'])"
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/memcheck_smoke3.log 2>&1
echo "memcheck smoke exit $?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|Misaligned|Error" gpurun_out/memcheck_smoke3.log | head -6
