#!/bin/bash
# pass N (1 GPU, last minutes of the round): the default bench line on the final build, CPU baseline leg off
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c2_n_err.txt > gpurun_out/bench_c2_n.json; echo "c2 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d.get('gpu_launches'))"
tail -2 gpurun_out/bench_c2_n_err.txt
