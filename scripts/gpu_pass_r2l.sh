#!/bin/bash
# pass L (1 GPU, short): Analogy throughput through the public evaluator + the parity suite of the final tree
mkdir -p gpurun_out
timeout 200 python scripts/analogy_perf.py 2>&1 | tail -4 | tee gpurun_out/analogy_perf.txt
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_l.txt
