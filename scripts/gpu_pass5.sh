#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json | cut -c1-600
tail -5 gpurun_out/bench_err.txt
QP_MODELS=l2,dm,cx timeout 300 python scripts/quick_perf.py 1000000 8192 2>&1 | tee gpurun_out/quick_perf.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recheck_kernel -s 2 -c 2 -o gpurun_out/recheck_c2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_recheck.txt 2>&1
tail -3 gpurun_out/ncu_recheck.txt
ls -la gpurun_out
