#!/bin/bash
# end-of-round run on one B200: GPU suite, smoke, the bench line, the reference arm, BASELINE configs[2..4] at their per-GPU size,
# the ncu launch list of the bench command and one --set full capture of the fused attention kernel (CSV exported on the box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/test_final.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/test_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.2f e2e %.2f ms/step %.1f launches %d whole-path frac %.3f roofline frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d.get('gpu_launches', 0), d.get('whole_path', {}).get('frac', 0), d.get('roofline', {}).get('frac', 0), d.get('clocks')))
" $1; }
show gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference arm exit $?"; tail -c 400 gpurun_out/bench_final_reference.json; echo
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --max-length 128 > gpurun_out/bench_final_cfg3.json 2> gpurun_out/bench_final_cfg3.err; echo "cfg3 exit $?"; show gpurun_out/bench_final_cfg3.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 16 --max-length 128 --num-beams 4 --early-stopping > gpurun_out/bench_final_cfg4.json 2> gpurun_out/bench_final_cfg4.err; echo "cfg4 exit $?"; show gpurun_out/bench_final_cfg4.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch 64 --image-size 1024 > gpurun_out/bench_final_cfg5.json 2> gpurun_out/bench_final_cfg5.err; echo "cfg5 exit $?"; show gpurun_out/bench_final_cfg5.json
timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch 16 --max-length 300 --num-beams 4 --early-stopping > gpurun_out/bench_final_script.json 2> gpurun_out/bench_final_script.err; echo "script defaults exit $?"; show gpurun_out/bench_final_script.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none -k regex:attn_fused -s 1368 -c 1 -o /tmp/r2_attn_fused_final -f python tools/attn_probe.py > gpurun_out/ncu_attn_fused_final.log 2>&1; echo "attn capture exit $?"
ncu -i /tmp/r2_attn_fused_final.ncu-rep --page raw --csv > gpurun_out/r2_attn_fused_final.raw.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail -4
