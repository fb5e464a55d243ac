#!/bin/bash
# Round-2 GPU pass F (1 GPU): ncu captures that back the C3 and C4 lines -- the RotatE approximate-sqrt scan at the
# shape of ONE C4 shard (625,000 rows, d = 1000), the tensor-core scan and the recheck at C3 (ComplEx d = 400).
mkdir -p gpurun_out
# RotatE d = 1000 against the CPU oracle at a table the oracle can still score in one piece (500k entities)
timeout 420 python bench.py --workload c4s --n-test 512 --cpu-sample 4 --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench_c4s_err.txt > gpurun_out/bench_c4s.json; echo "c4s rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c4s.json'));print(d['value'],d['ms_per_step'],d['parity_full']['ranks_equal'],d['cpu_baseline'])"
tail -3 gpurun_out/bench_c4s_err.txt
QP_MODELS=rot1k timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/rot_scan_c4_shard -f python scripts/quick_perf.py 625000 2048 > gpurun_out/ncu_rot_c4.txt 2>&1
tail -3 gpurun_out/ncu_rot_c4.txt | cut -c1-300
timeout 700 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -o gpurun_out/tc_scan_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_tc_c3.txt 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:recheck_kernel -s 2 -c 1 -o gpurun_out/recheck_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_rc_c3.txt 2>&1
tail -2 gpurun_out/ncu_tc_c3.txt gpurun_out/ncu_rc_c3.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu_c3.txt 2>&1
ls -la gpurun_out | grep -E "ncu-rep|launches_c3"
