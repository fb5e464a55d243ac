#!/bin/bash
# Round-2 GPU pass D (1 GPU): full parity suite on the final kernels, bench c2 (default line), c3, c1, the
# 8-GPU per-rank workload (2558 test triples) on one GPU.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_d.txt
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_c2_d_err.txt > gpurun_out/bench_c2_d.json; echo "c2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_d.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],r['near_tie_fraction'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d['cpu_baseline']['parity_on_sample'])"
tail -3 gpurun_out/bench_c2_d_err.txt
KGE_TRACE=1 timeout 300 python bench.py --n-test 2558 --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench_c2_q2558_d_err.txt > gpurun_out/bench_c2_q2558_d.json; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_q2558_d.json'));r=d['roofline'];print('q2558',d['value'],d['ms_per_step'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['e2e']['ms_per_step'])"
grep trace gpurun_out/bench_c2_q2558_d_err.txt | tail -8
timeout 600 python bench.py --workload c3 --steps 3 --warmup 2 --no-extras --cpu-sample 64 2>gpurun_out/bench_c3_d_err.txt > gpurun_out/bench_c3_d.json; echo "c3 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c3_d.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],r['near_tie_fraction'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d['cpu_baseline'])"
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 --no-extras 2>gpurun_out/bench_c1_d_err.txt > gpurun_out/bench_c1_d.json; echo "c1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c1_d.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_full']['ranks_equal'],d['cpu_baseline'])"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
