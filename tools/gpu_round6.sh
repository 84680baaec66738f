#!/bin/bash
mkdir -p gpurun_out
# 4th case of the timeline step = M=928 N=1024 K=4096 bn=64 without LN interleave: launches (50 warm + 50) per case, 2 cases per shape
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 650 -c 1 -o gpurun_out/prof_v2_mproj python tools/gpu_diag.py timeline > gpurun_out/ncu6.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu6.log
