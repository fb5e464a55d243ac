#!/bin/bash
# final 8-GPU check of the default bench line (what the driver's scaling run executes at N = 8)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench_c2_n8b_err.txt > gpurun_out/bench_c2_n8b.json
echo "c2 n8 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n8b.json'))
r=d['roofline']; print('queries', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_full']['ranks_equal'], d['ranks_sha256_16'], r['ms_per_launch'], r['recheck_ms_per_launch'], d['cpu_baseline'])
e=d.get('entities',{}); print('entities', e.get('value'), e.get('ms_per_step'), e.get('e2e',{}).get('value'), e.get('parity_full',{}).get('ranks_equal'), e.get('ranks_sha256_16'))"
grep -v "NCCL INFO" gpurun_out/bench_c2_n8b_err.txt | tail -4
grep "NCCL INFO" gpurun_out/bench_c2_n8b_err.txt | grep -i -E "nranks 8|NVLS multicast" | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c2_n4_err.txt > gpurun_out/bench_c2_n4.json
echo "c2 n4 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n4.json'))
print('n4 queries', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_full']['ranks_equal'], d['ranks_sha256_16']); e=d.get('entities',{}); print('n4 entities', e.get('value'), e.get('ms_per_step'))"
