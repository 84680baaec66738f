#!/bin/bash
# last run of the round on the final commit: GPU suite, smoke, the bench line, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/test_final.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/test_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value %.2f e2e %.2f ms/step %.1f launches %d whole-path frac %.3f roofline frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['whole_path']['frac'], d['roofline']['frac'], d['clocks']))
print(d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1; echo "launch list exit $?"
