#!/bin/bash
# Round-2 GPU pass B (1 GPU): all parity tests (top-k epilogue, RESCAL relation path, cached image, fp16 split),
# fp16 numerics probe, bench c2 / c3 under the fp16 operand split.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_b.txt
timeout 280 python scripts/tc_numerics_probe.py --fp16 > gpurun_out/tc_numerics_probe_fp16.txt 2>&1; tail -2 gpurun_out/tc_numerics_probe_fp16.txt | cut -c1-300
for fmt in 1 0; do
  KGE_TC_FP16=$fmt timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/bench_c2_fp16_${fmt}_err.txt > gpurun_out/bench_c2_fp16_${fmt}.json; echo "bench c2 fp16=$fmt rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_fp16_${fmt}.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['ms_per_launch'],r['recheck_ms_per_launch'],r['near_tie_fraction'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'])"
  tail -3 gpurun_out/bench_c2_fp16_${fmt}_err.txt
done
KGE_TC_FP16=1 timeout 400 python bench.py --workload c3 --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench_c3_fp16_1_err.txt > gpurun_out/bench_c3_fp16_1.json; echo "bench c3 fp16 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c3_fp16_1.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['ms_per_launch'],r['recheck_ms_per_launch'],r['near_tie_fraction'],d['parity_full']['ranks_equal'],d['ranks_sha256_16'])"
tail -3 gpurun_out/bench_c3_fp16_1_err.txt
ls -la gpurun_out | head -30
