#!/bin/bash
# compute-sanitizer passes over the small-shape tests (memcheck: OOB / misaligned accesses; racecheck: shared-memory hazards)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_preprocess.py -m gpu -q -x \
  -k "rpn or roi or beam or conv3x3 or (gemm and 128-128-64) or (gemm and 300-200) or (gemm and 37-800) or activations or (cta_pair and 100-256) or (cta_pair and 300-512) or (cta_pair and 129-1024) or (split_k and 300-512) or (preprocess and 1500-1000) or (preprocess and 1024-1024)" > gpurun_out/memcheck_kernels.log 2>&1
echo "memcheck kernels exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/memcheck_kernels.log | head -12
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|Misaligned|Error" gpurun_out/memcheck_smoke.log | head -12
# the decode step at a size that exercises the fused attention kernel with several M tiles, the CTA-pair projections and beam search
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "
import torch, numpy as np
from rgrg_b200 import Engine, synth
e = Engine(0); e.load_state_dict(synth.make_partial_state_dict(0, ('detector', 'heads', 'lm')))
f = torch.randn(300, 1024, generator=torch.Generator().manual_seed(1)).cuda()
a = e.lm_generate(f, 20); print('greedy', a.shape)
b = e.lm_generate(f[:40], 10, num_beams=4, early_stopping=True); print('beam', b.shape)
" > gpurun_out/memcheck_decode.log 2>&1
echo "memcheck decode exit $?"; grep -E "ERROR SUMMARY|greedy|beam|Invalid|Misaligned|Error" gpurun_out/memcheck_decode.log | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "rpn_filter_bit_exact or roi_tail or beam_bookkeeping or roi_align_separable" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | head -12
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -c "
import torch
from rgrg_b200 import Engine, synth
e = Engine(0); e.load_state_dict(synth.make_partial_state_dict(0, ('detector', 'heads', 'lm')))
f = torch.randn(150, 1024, generator=torch.Generator().manual_seed(1)).cuda()
e.set_option('cuda_graph', 0)
print('greedy', e.lm_generate(f, 6).shape)
" > gpurun_out/racecheck_decode.log 2>&1
echo "racecheck decode exit $?"; grep -E "RACECHECK SUMMARY|greedy|hazard|Error" gpurun_out/racecheck_decode.log | head -12
