#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
rm -f gpurun_out/quick_perf.txt
for cfg in "32 1 0" "32 0 0" "64 0 0" "32 0 32" "32 0 8"; do
  set -- $cfg
  KGE_TC_BK=$1 KGE_TC_RESIDENT=$2 KGE_TC_GROUP=$3 QP_MODELS=l2,dm,cx timeout 300 python scripts/quick_perf.py 1000000 8192 2>&1 | sed "s/^/bk$1 res$2 grp$3: /" | tee -a gpurun_out/quick_perf.txt
done
KGE_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_trace.txt | tee gpurun_out/bench_c2_trace.json | cut -c1-400
grep -c trace gpurun_out/bench_trace.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json | cut -c1-600
KGE_TC_RESIDENT=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2_res0.json | cut -c1-600
tail -5 gpurun_out/bench_err.txt
ls -la gpurun_out
