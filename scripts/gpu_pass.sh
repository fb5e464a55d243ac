#!/bin/bash
# One GPU pass (run with: gpurun --timeout 1500 -- 'bash scripts/gpu_pass.sh'): parity tests, smoke,
# bench (c2), per-model probes, training bench (c5), ncu launch list of a step and full captures of
# the three kernels the roofline claims rest on.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json | cut -c1-600
tail -5 gpurun_out/bench_err.txt
rm -f gpurun_out/train_bench.jsonl
timeout 600 python scripts/train_bench.py --cpu-batch 0 2>gpurun_out/train_err.txt | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:margin_step_fast_kernel -s 8 -c 2 -o gpurun_out/train_fast -f python scripts/train_bench.py --batch 32768 --cpu-batch 0 --reps 2 > gpurun_out/ncu_train.txt 2>&1
QP_MODELS=rot1k timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/rot_scan -f python scripts/quick_perf.py 1000000 2048 > gpurun_out/ncu_rot.txt 2>&1
tail -2 gpurun_out/ncu_train.txt gpurun_out/ncu_rot.txt
ls -la gpurun_out
