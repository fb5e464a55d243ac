#!/bin/bash
# Round-2 GPU pass H (1 GPU): C3 with a 256-triple sample through the unmodified reference on the host cores.
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c3 --steps 5 --warmup 2 --no-extras --cpu-sample 256 2>gpurun_out/bench_c3_s256_err.txt > gpurun_out/bench_c3_s256.json; echo "c3 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c3_s256.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d['cpu_baseline'])"
tail -2 gpurun_out/bench_c3_s256_err.txt
