#!/bin/bash
# C4 (BASELINE.json configs[3]): RotatE d=1000, |E| = 5M, entity table range-partitioned over 8 B200 (each rank
# generates and holds only its 625k rows), query rows exchanged by all-reduce, ONE all-reduce of the rank counters;
# then the default workload (c2) at 8 GPUs in both decompositions, as the driver's scaling run does.
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_pass_c4.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -9
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/bench_c4_n8_err.txt > gpurun_out/bench_c4_n8.json
echo "c4 rc=$?"; cut -c1-900 gpurun_out/bench_c4_n8.json; grep -v "NCCL INFO" gpurun_out/bench_c4_n8_err.txt | tail -5; grep -c "NCCL INFO" gpurun_out/bench_c4_n8_err.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c2_n8_err.txt > gpurun_out/bench_c2_n8.json
echo "c2 n8 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n8.json'))
print('queries', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_full']['ranks_equal'], d['ranks_sha256_16'])
e=d.get('entities',{}); print('entities', e.get('value'), e.get('ms_per_step'), e.get('e2e',{}).get('value'), e.get('parity_full',{}).get('ranks_equal'), e.get('ranks_sha256_16'))"
grep -v "NCCL INFO" gpurun_out/bench_c2_n8_err.txt | tail -5
grep "NCCL INFO" gpurun_out/bench_c2_n8_err.txt | grep -i -E "nranks|NVLS|Connected all" | head -8
