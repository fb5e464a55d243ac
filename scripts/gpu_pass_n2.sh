#!/bin/bash
# two-GPU pass: bench through torchrun in both decompositions
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench_n2_err.txt | tee gpurun_out/bench_c2_n2.json | cut -c1-700
tail -5 gpurun_out/bench_n2_err.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --split entities --no-cpu-baseline 2>gpurun_out/bench_n2e_err.txt | tee gpurun_out/bench_c2_n2_entities.json | cut -c1-700
tail -5 gpurun_out/bench_n2e_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-300
ls -la gpurun_out
