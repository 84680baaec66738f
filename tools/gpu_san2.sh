#!/bin/bash
# memcheck passes for the kernels changed late in round 2 (TMA-store epilogue, slot-interleaved cache, attention options), then the GPU suite
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x \
  -k "(cta_pair and 100-256) or (cta_pair and 300-512) or (cta_pair and 129-1024) or (split_k and 300-512)" > gpurun_out/memcheck_kernels2.log 2>&1
echo "memcheck kernels exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/memcheck_kernels2.log | head -12
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "
import torch, numpy as np
from rgrg_b200 import Engine, synth
e = Engine(0); e.load_state_dict(synth.make_partial_state_dict(0, ('detector', 'heads', 'lm')))
f = torch.randn(300, 1024, generator=torch.Generator().manual_seed(1)).cuda()
a = e.lm_generate(f, 20); print('greedy', a.shape)
for k in ('attn_mc', 'attn_early'):
    e.set_option(k, 1); b = e.lm_generate(f, 20); e.set_option(k, 0); print(k, bool((a == b).all()))
e.set_option('epi_tma', 0); b = e.lm_generate(f, 20); e.set_option('epi_tma', 1); print('epi_tma=0', bool((a == b).all()))
b = e.lm_generate(f[:40], 10, num_beams=4, early_stopping=True); print('beam', b.shape)
" > gpurun_out/memcheck_decode2.log 2>&1
echo "memcheck decode exit $?"; grep -E "ERROR SUMMARY|greedy|beam|attn_|epi_tma|Invalid|Misaligned|Error" gpurun_out/memcheck_decode2.log | head -12
timeout 900 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -3
