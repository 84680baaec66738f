#!/bin/bash
# ncu --set full of the fused attention kernel at cache length ~58 (step 56 of 63), plus the decode GEMMs
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 1344 -c 2 -o gpurun_out/r2_attn_fused -f python tools/attn_probe.py > gpurun_out/ncu_f1.log 2>&1
echo "ncu1 exit $?"; tail -3 gpurun_out/ncu_f1.log
ncu --set full --clock-control none -k regex:gemm_tc -s 3000 -c 4 -o gpurun_out/r2_decode_gemms -f python tools/attn_probe.py > gpurun_out/ncu_f2.log 2>&1
echo "ncu2 exit $?"; tail -3 gpurun_out/ncu_f2.log
ls -la gpurun_out/*.ncu-rep
