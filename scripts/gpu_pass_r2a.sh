#!/bin/bash
# Round-2 GPU pass A (1 GPU): tensor-core numerics probes, parity tests, new bench lines (c2 with extras, c3).
# gpurun --timeout 1500 -- 'bash scripts/gpu_pass_r2a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400 > gpurun_out/lscpu.txt
free -g | head -2 >> gpurun_out/lscpu.txt
timeout 280 python scripts/tc_numerics_probe.py > gpurun_out/tc_numerics_probe.txt 2>&1; tail -3 gpurun_out/tc_numerics_probe.txt | cut -c1-300
timeout 200 python scripts/tc_accum_probe.py > gpurun_out/tc_accum_probe.txt 2>&1; tail -2 gpurun_out/tc_accum_probe.txt | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_c2_err.txt > gpurun_out/bench_c2.json; echo "bench c2 rc=$?"; cut -c1-1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2_err.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --workload c3 --steps 3 --warmup 2 --no-extras 2>gpurun_out/bench_c3_err.txt > gpurun_out/bench_c3.json; echo "bench c3 rc=$?"; cut -c1-1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3_err.txt
ls -la gpurun_out | head -40
