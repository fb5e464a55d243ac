#!/bin/bash
# pass J (1 GPU, short): backward training kernel held to 96 registers (5 CTAs per SM) against the compiler's 128
mkdir -p gpurun_out
KGE_TRAIN_BWD_BLOCKS=5 timeout 200 python -m pytest tests/test_train_gpu.py tests/test_toruse_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_train_j.txt
timeout 150 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c5_j_err.txt > gpurun_out/bench_c5_j128.json; echo "c5 128 rc=$?"
KGE_TRAIN_BWD_BLOCKS=5 timeout 150 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_c5_j_err.txt > gpurun_out/bench_c5_j96.json; echo "c5 96 rc=$?"
python - <<'PY'
import json
for f in ("j128", "j96"):
    d = json.load(open("gpurun_out/bench_c5_%s.json" % f))
    for b, v in d["by_batch"].items():
        print(f, b, "fwd %.3f ms (%.2f)  fwd+bwd %.3f ms (%.2f)  e2e %.0f/s" % (v["fwd_ms"], v["roofline_fwd"]["frac"], v["fwd_bwd_ms"], v["roofline_fwd_bwd"]["frac"], v["e2e"]["value"]))
    print(f, "parity", d["parity"])
PY
tail -2 gpurun_out/bench_c5_j_err.txt
