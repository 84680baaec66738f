#!/bin/bash
# 2 GPUs: engine-side gather vs torch gather, then the bench under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'n',d['n_gpus'])"
grep -i "gather" gpurun_out/bench_n2.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --gather torch > gpurun_out/bench_n2_torch.json 2> gpurun_out/bench_n2_torch.err; echo "bench n2 torch exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n2_torch.json'));print('torch gather: value',d['value'],'ms/step',d['ms_per_step'])"
