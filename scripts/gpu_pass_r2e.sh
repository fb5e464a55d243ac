#!/bin/bash
# Round-2 GPU pass E (1 GPU): C4's table (RotatE d=1000, 5M entities, 40 GB + 40 GB packed) on ONE GPU with the
# first 256 test triples: GPU ranks against the CPU oracle (table scored in 8 entity-range pieces) on a 4-triple
# sample -- the same first test triples as the 8-GPU run (compare ranks_first8); top-k inference probe at |E| = 1M.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_inference_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_e.txt
timeout 400 python scripts/topk_perf.py 1000000 4096 10 2>&1 | tail -4
timeout 1500 python bench.py --workload c4 --n-test 256 --cpu-sample 4 --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench_c4_n1_err.txt > gpurun_out/bench_c4_n1_q256.json; echo "c4 n1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c4_n1_q256.json'));print(d['value'],d['ms_per_step'],d['parity_full']['ranks_equal'],d['cpu_baseline'],d['ranks_first8'])"
tail -3 gpurun_out/bench_c4_n1_err.txt
