#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
rm -f gpurun_out/quick_perf.txt
QP_MODELS=rot,rot1k timeout 300 python scripts/quick_perf.py 1000000 2048 2>&1 | sed "s/^/refine: /" | tee -a gpurun_out/quick_perf.txt
QP_TC=0 QP_MODELS=rot,rot1k timeout 300 python scripts/quick_perf.py 1000000 2048 2>&1 | sed "s/^/exact : /" | tee -a gpurun_out/quick_perf.txt
QP_MODELS=l2,dm,cx timeout 300 python scripts/quick_perf.py 1000000 8192 2>&1 | tee -a gpurun_out/quick_perf.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json | cut -c1-600
timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tee gpurun_out/bench_c3.json | cut -c1-600
tail -5 gpurun_out/bench_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|recheck|pack|row_norms|prep|true_scores|filter|finalize|stats|fill" --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
ls -la gpurun_out
