#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/test_gpu8.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/test_gpu8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
# decode-loop kernels, current versions: full ncu capture of one c_fc GEMM, one split-K c_proj and one attention launch
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attention_kernel" -s 4500 -c 4 -o gpurun_out/prof_decode python tools/attn_probe.py > gpurun_out/ncu8.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu8.log
# launch list of the bench command, first 1500 launches (weight repack + detector + first decode steps)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches8.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu8_list.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches8.csv
