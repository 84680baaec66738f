#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "bench exit $?"; tail -2 gpurun_out/bench7.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err
echo "ref exit $?"; cat gpurun_out/bench7_ref.json | head -c 1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench7_n2.json 2> gpurun_out/bench7_n2.err
echo "n2 exit $?"; python -c "
import json
for f in ('gpurun_out/bench7.json','gpurun_out/bench7_n2.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, 'value',d['value'],'e2e',d['e2e']['value'],'n',d['n_gpus'],'ms',d['ms_per_step'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
"; tail -5 gpurun_out/bench7_n2.err
