#!/bin/bash
mkdir -p gpurun_out
for step in tc_small tc_shapes conv gemm_perf; do
  echo "=== diag $step" >> gpurun_out/diag3.log
  timeout 150 python tools/gpu_diag.py $step >> gpurun_out/diag3.log 2>&1
  echo "exit $?" >> gpurun_out/diag3.log
done
grep -E "MISMATCH|exit|perf|Error|error|timed out" gpurun_out/diag3.log | head -60
timeout 700 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/test_gpu3.log 2>&1
echo "tests exit $?"; tail -25 gpurun_out/test_gpu3.log
timeout 600 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err
echo "bench exit $?"; cat gpurun_out/bench3.json | head -c 6000; tail -12 gpurun_out/bench3.err
