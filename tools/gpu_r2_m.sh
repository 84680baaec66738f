#!/bin/bash
# ncu --set full of every GEMM launch of the detector + the first decode layer, of lm_head and of RoIAlign; the reports are
# converted to CSV on the box (gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
cap() { # name, ncu args...
  name=$1; shift
  ncu --set full --clock-control none "$@" -o /tmp/$name -f python tools/ncu_probe.py > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep gpurun_out/$name.raw.csv
}
cap r2_detector_gemms -k regex:gemm_tc -c 64
cap r2_lm_head -k regex:gemm_tc -s 132 -c 1
cap r2_roi_align_sep -k regex:roi_align -c 1
ncu --set full --clock-control none -k regex:gemm_2cta -s 3000 -c 3 -o /tmp/r2_gemm_2cta -f python tools/attn_probe.py > gpurun_out/ncu_r2_gemm_2cta.log 2>&1; echo "2cta exit $?"
ncu -i /tmp/r2_gemm_2cta.ncu-rep --page raw --csv > gpurun_out/r2_gemm_2cta.raw.csv 2>/dev/null
