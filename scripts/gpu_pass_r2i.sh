#!/bin/bash
# last sanity pass of the round (1 GPU): the parity suite and the default bench line on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_i.txt
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_c2_i_err.txt > gpurun_out/bench_c2_i.json; echo "c2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_i.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],d['e2e']['value'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'],d['cpu_baseline']['parity_on_sample']['ranks_equal'], d['api_reference_kg']['first_call']['value'], d['api_reference_kg']['steady_state']['value'], d['filter_index_build_s'], d['c5_training_step']['by_batch']['32768']['roofline_fwd_bwd']['frac'])"
tail -2 gpurun_out/bench_c2_i_err.txt
