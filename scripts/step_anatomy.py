"""Where does one link-prediction step spend its time?  Runs the device-resident step of a
bench workload under torch.profiler (CUPTI), then prints per-kernel totals, the GPU busy time
of a step and every idle gap > 50 us with the kernels on either side.

    python scripts/step_anatomy.py [workload] [steps]      ->  gpurun_out/anatomy_<workload>.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torchkge_b200 import synthetic as S  # noqa: E402
from torchkge_b200.engine import CudaEngine, ModelSpec, rank_link_prediction  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    wl = S.WORKLOADS[name]
    dev = torch.device("cuda:0")
    code = S.MODEL_CODE[(wl["model"], wl["diss"])]
    tabs = S.make_tables(code, wl["dim"], wl["n_ent"], wl["n_rel"], 0, wl["n_ent"], 0, dev)
    graph = S.make_graph(wl["n_ent"], wl["n_rel"], wl["n_facts"], wl["n_test"], 0, dev)
    csr_t, csr_h = S.make_filters(graph, wl["n_ent"], wl["n_rel"])
    spec = ModelSpec(code, wl["dim"], wl["n_ent"], wl["n_rel"], tabs["ent0"], tabs["ent1"],
                     tabs["rel0"], tabs["rel1"])
    eng = CudaEngine()
    h, t, r = graph["test_h"], graph["test_t"], graph["test_r"]

    def step():
        return rank_link_prediction(spec, h, t, r, csr_t, csr_h, engine=eng)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    trace = os.path.join(ROOT, "gpurun_out", "anatomy_%s_trace.json" % name)
    prof.export_chrome_trace(trace)
    ev = json.load(open(trace))["traceEvents"]
    ks = sorted(((e["ts"], e["dur"], e["name"]) for e in ev
                 if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e))
    os.remove(trace)
    if not ks:
        print("no GPU activity captured (CUPTI unavailable?)")
        return
    per = {}
    for ts, dur, nm in ks:
        short = nm.split("(")[0][-60:]
        a = per.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += dur
    span = ks[-1][0] + ks[-1][1] - ks[0][0]
    busy = sum(d for _, d, _ in ks)
    gaps = []
    end = ks[0][0] + ks[0][1]
    prev = ks[0][2]
    for ts, dur, nm in ks[1:]:
        if ts - end > 50:
            gaps.append({"us": round(ts - end, 1), "after": prev.split("(")[0][-50:], "before": nm.split("(")[0][-50:]})
        if ts + dur > end:
            end, prev = ts + dur, nm
    out = {"workload": name, "steps": steps, "span_ms_per_step": span / 1000 / steps,
           "busy_ms_per_step": busy / 1000 / steps, "idle_ms_per_step": (span - busy) / 1000 / steps,
           "kernels_ms_per_step": {k: [v[0] // steps, round(v[1] / 1000 / steps, 3)]
                                   for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])},
           "gaps_over_50us": sorted(gaps, key=lambda g: -g["us"])[:25]}
    path = os.path.join(ROOT, "gpurun_out", "anatomy_%s.json" % name)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1)[:6000])


if __name__ == "__main__":
    main()
