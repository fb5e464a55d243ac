#!/bin/bash
# 2 x B200: the multi-device guard test (model on cuda:1 while cuda:0 is current), the sharding GPU paths, and the
# default bench line at N = 2 (both decompositions).   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_pass_n2.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lp_gpu.py -m gpu -q -k "non_current_device or golden" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_n2.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c2_n2_err.txt > gpurun_out/bench_c2_n2.json
echo "c2 n2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n2.json'))
print('queries', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_full']['ranks_equal'], d['ranks_sha256_16'])
e=d.get('entities',{}); print('entities', e.get('value'), e.get('ms_per_step'), e.get('e2e',{}).get('value'), e.get('parity_full',{}).get('ranks_equal'), e.get('ranks_sha256_16'))"
grep -v "NCCL INFO" gpurun_out/bench_c2_n2_err.txt | tail -5
grep "NCCL INFO" gpurun_out/bench_c2_n2_err.txt | grep -i -E "nranks|NVLS|Connected all|Init COMPLETE" | head -6
