#!/bin/bash
# compute-sanitizer passes over the small-shape tests (memcheck: OOB / misaligned accesses; racecheck: shared-memory hazards)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x \
  -k "rpn or roi or beam or conv3x3 or (gemm and 128-128-64) or (gemm and 300-200) or (gemm and 37-800) or activations" > gpurun_out/memcheck_kernels.log 2>&1
echo "memcheck kernels exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/memcheck_kernels.log | head -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|Misaligned|Error" gpurun_out/memcheck_smoke.log | head -12
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "rpn_filter_bit_exact or roi_tail or beam_bookkeeping" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | head -12
