#!/bin/bash
# Round-2 GPU pass G (1 GPU): final build -- full parity suite, the default bench line, RotatE d = 1000 against the
# (fast-modulus) CPU oracle at 500k entities and, time permitting, at C4's full 5M-entity table.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_g.txt
timeout 500 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_c2_g_err.txt > gpurun_out/bench_c2_g.json; echo "c2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_g.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d['cpu_baseline']['parity_on_sample'], d['api_reference_kg']['first_call']['value'], d['api_reference_kg']['steady_state']['value'])"
tail -2 gpurun_out/bench_c2_g_err.txt
timeout 400 python bench.py --workload c4s --n-test 512 --cpu-sample 4 --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench_c4s_err.txt > gpurun_out/bench_c4s.json; echo "c4s rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c4s.json'));print(d['value'],d['ms_per_step'],d['parity_full']['ranks_equal'],d['cpu_baseline'])"
tail -2 gpurun_out/bench_c4s_err.txt
timeout 420 python bench.py --workload c4 --n-test 256 --cpu-sample 2 --steps 1 --warmup 1 --no-extras 2>gpurun_out/bench_c4_n1_err.txt > gpurun_out/bench_c4_n1_q256.json; echo "c4 n1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c4_n1_q256.json'));print(d['value'],d['ms_per_step'],d['parity_full']['ranks_equal'],d['cpu_baseline'],d['ranks_first8'])"
tail -4 gpurun_out/bench_c4_n1_err.txt
