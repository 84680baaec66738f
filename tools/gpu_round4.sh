#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py timeline > gpurun_out/timeline.log 2>&1
echo "timeline exit $?"; cat gpurun_out/timeline.log | tail -20
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 1 -o gpurun_out/prof_v2_cattn python tools/gpu_diag.py timeline > gpurun_out/ncu_v2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_v2.log
