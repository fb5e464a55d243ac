#!/bin/bash
# C4 (BASELINE.json configs[3]): RotatE d=1000, |E| = 5M, entity table range-partitioned over 8 B200,
# one NCCL all-reduce on the rank counters.  Run with: gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_pass_c4.sh'
# (~6 s per evaluation expected; the CPU baseline is skipped: it would need the 40 GB table on the host).
set -x
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c4_err.txt | tee gpurun_out/bench_c4_n8.json | cut -c1-700
tail -5 gpurun_out/bench_c4_err.txt
