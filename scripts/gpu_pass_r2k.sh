#!/bin/bash
# pass K (1 GPU): Analogy on the GPU (three-plane element), then the whole parity suite on the new build
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_analogy_gpu.py -m gpu -q 2>&1 | tail -45 | tee gpurun_out/pytest_analogy_k.txt
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_k.txt
