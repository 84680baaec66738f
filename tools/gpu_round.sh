#!/bin/bash
# One GPU-box session: bring-up diagnostics, then the test suites.  Every step under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for step in simt tc_small tc_k tc_shapes conv gemm_perf; do
  echo "=== diag $step" >> gpurun_out/diag.log
  timeout 240 python tools/gpu_diag.py $step >> gpurun_out/diag.log 2>&1
  echo "exit $?" >> gpurun_out/diag.log
done
tail -60 gpurun_out/diag.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/test_kernels.log 2>&1
echo "kernels exit $?"; tail -15 gpurun_out/test_kernels.log
timeout 1200 python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 900 > gpurun_out/test_path.log 2>&1
echo "path exit $?"; tail -30 gpurun_out/test_path.log
