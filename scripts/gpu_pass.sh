#!/bin/bash
# GPU pass: parity tests, step anatomy, geometry comparison, bench (c2), training bench (c5),
# ncu launch list + one full capture of the dominant kernel.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
# geometry comparison, one side launches at 8192 queries x 1M rows
for cfg in "32 1" "32 0" "64 0"; do
  set -- $cfg
  KGE_TC_BK=$1 KGE_TC_RESIDENT=$2 QP_MODELS=l2,dm,cx timeout 300 python scripts/quick_perf.py 1000000 8192 2>&1 | sed "s/^/bk$1 res$2: /" | tee -a gpurun_out/quick_perf.txt
done
timeout 600 python scripts/step_anatomy.py c2 2 > gpurun_out/anatomy_stdout.txt 2>&1; tail -3 gpurun_out/anatomy_stdout.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json | cut -c1-600
tail -5 gpurun_out/bench_err.txt
timeout 600 python scripts/train_bench.py 2>gpurun_out/train_err.txt | cut -c1-900
tail -5 gpurun_out/train_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|recheck|pack|rows|prep|true_scores|filter|finalize|stats|fill" --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -o gpurun_out/tc_scan_c2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.txt 2>&1
tail -3 gpurun_out/ncu_full.txt
ls -la gpurun_out
