#!/bin/bash
# Round-2 GPU pass C (1 GPU): training tests with the ring kernel, c5 bench (ring vs register form), c2 with a
# 256-triple reference sample, ncu launch list + full captures of the three kernels the roofline claims rest on.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_relpred_gpu.py tests/test_tripletclf_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_c.txt
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 2>gpurun_out/bench_c5_err.txt > gpurun_out/bench_c5_ring.json; echo "c5 ring rc=$?"
KGE_TRAIN_RING=0 timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_c5_err.txt > gpurun_out/bench_c5_regs.json; echo "c5 regs rc=$?"
python - <<'PY'
import json
for f in ("ring","regs"):
    d=json.load(open("gpurun_out/bench_c5_%s.json"%f))
    for b,v in d["by_batch"].items():
        print(f, b, "fwd %.3f ms (%.2f)  fwd+bwd %.3f ms (%.2f)  e2e %.0f/s" % (v["fwd_ms"], v["roofline_fwd"]["frac"], v["fwd_bwd_ms"], v["roofline_fwd_bwd"]["frac"], v["e2e"]["value"]))
    print(f, "parity", d["parity"], "cpu", d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_c5_err.txt
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 256 2>gpurun_out/bench_c2_s256_err.txt > gpurun_out/bench_c2_s256.json; echo "c2 s256 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_s256.json'));print(d['value'],d['e2e']['value'],d['cpu_baseline'],d['api_reference_kg']['first_call'],d['api_reference_kg']['steady_state'])"
tail -3 gpurun_out/bench_c2_s256_err.txt
# the per-rank workload of an 8-GPU query-sharded run (2558 test triples) on one GPU: where does the step go?
KGE_TRACE=1 timeout 300 python bench.py --n-test 2558 --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench_c2_q2558_err.txt > gpurun_out/bench_c2_q2558.json; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_q2558.json'));r=d['roofline'];print('q2558',d['value'],d['ms_per_step'],r['ms_per_launch'],r['recheck_ms_per_launch'],d['e2e']['ms_per_step'])"
grep trace gpurun_out/bench_c2_q2558_err.txt | tail -14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2_q2558.csv python bench.py --n-test 2558 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu_q2558.txt 2>&1
# ncu: launch list of two c2 steps, then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -o gpurun_out/tc_scan_c2_fp16 -f python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_tc.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recheck_kernel -s 2 -c 1 -o gpurun_out/recheck_c2_fp16 -f python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_rc.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:margin_step_ring_kernel -s 6 -c 2 -o gpurun_out/train_ring -f python bench.py --workload c5 --steps 3 --no-cpu-baseline > gpurun_out/ncu_train.txt 2>&1
tail -2 gpurun_out/ncu_tc.txt gpurun_out/ncu_rc.txt gpurun_out/ncu_train.txt
ls -la gpurun_out | head -40
