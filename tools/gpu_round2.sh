#!/bin/bash
# tests + bench + ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_gpu.log 2>&1
echo "tests exit $?"; tail -15 gpurun_out/test_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3000 -c 3 -o gpurun_out/prof_gemm \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/
