#!/bin/bash
# pass M (1 GPU): tensor-core path for Analogy (K = 3 x plane width) -- whole parity suite, then throughput
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_m.txt
timeout 120 python scripts/analogy_perf.py 2>&1 | tail -3 | tee gpurun_out/analogy_perf_tc.txt
