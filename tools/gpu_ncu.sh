#!/bin/bash
# ncu --set full of the CTA-pair projections (TMA-store epilogue) and of the fused attention kernel at cache
# length ~58; reports are converted to CSV on the box (gpurun brings back at most 64 MiB)
mkdir -p gpurun_out

ncu --set full --clock-control none -k regex:gemm_2cta -s 3000 -c 3 -o /tmp/r2_gemm_2cta_tma -f python tools/attn_probe.py > gpurun_out/ncu_r2_gemm_2cta_tma.log 2>&1; echo "2cta exit $?"
ncu -i /tmp/r2_gemm_2cta_tma.ncu-rep --page raw --csv > gpurun_out/r2_gemm_2cta_tma.raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:attn_fused -s 1368 -c 1 -o /tmp/r2_attn_fused_v2 -f python tools/attn_probe.py > gpurun_out/ncu_r2_attn_fused_v2.log 2>&1; echo "attn exit $?"
ncu -i /tmp/r2_attn_fused_v2.ncu-rep --page raw --csv > gpurun_out/r2_attn_fused_v2.raw.csv 2>/dev/null
ls -la gpurun_out/*.raw.csv | tail -3
T=48 timeout 600 python tools/decode_timeline.py 2>&1 | tail -22
