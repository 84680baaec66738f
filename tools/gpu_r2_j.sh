#!/bin/bash
# ncu --set full: first invocation of every kernel of one generate() at batch 32 (detector + decode step 0..2)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-id :::1 -o gpurun_out/r2_all_kernels -f python tools/ncu_probe.py > gpurun_out/ncu_j1.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_j1.log
# launch list of the bench command (first 1500 launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_j2.log 2>&1
echo "ncu list exit $?"; tail -2 gpurun_out/ncu_j2.log
ls -la gpurun_out/r2_all_kernels.ncu-rep gpurun_out/r2_launches.csv
