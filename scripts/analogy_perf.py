"""Analogy (three-plane element, exact scalar scan) through the public evaluator: triples/s of a full
filtered link-prediction evaluation at the FB15k-237 shape and at 1M entities, with a small oracle
sample at size.   python scripts/analogy_perf.py  ->  gpurun_out/analogy_perf.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchkge_b200 as tk  # noqa: E402
from oracle import kge_oracle as oracle  # noqa: E402  (checker only)


def run(n_ent, n_rel, emb_dim, n_test, sample):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    h = torch.randint(0, n_ent, (n_test,), generator=g)
    t = torch.randint(0, n_ent, (n_test,), generator=g)
    r = torch.randint(0, n_rel, (n_test,), generator=g)
    dh, dt = oracle.build_filter_dicts(h, t, r)
    kg = tk.KnowledgeGraph(h, t, r, n_ent, n_rel, dict_of_heads=dh, dict_of_tails=dt)
    torch.manual_seed(0)
    model = tk.AnalogyModel(emb_dim, n_ent, n_rel).to(dev)
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=256, verbose=False)          # warm-up (filter CSR built and cached on kg)
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev.evaluate(b_size=256, verbose=False)       # ranks land on the host: end to end
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    P = {k: v.detach().cpu() for k, v in (("sc_ent", model.sc_ent_emb.weight), ("re_ent", model.re_ent_emb.weight),
                                          ("im_ent", model.im_ent_emb.weight), ("sc_rel", model.sc_rel_emb.weight),
                                          ("re_rel", model.re_rel_emb.weight), ("im_rel", model.im_rel_emb.weight))}
    t0 = time.perf_counter()
    ref = oracle.link_prediction("analogy", P, h[:sample], t[:sample], r[:sample], dh, dt, b_size=4)
    cpu_s = time.perf_counter() - t0
    names = ["rank_true_heads", "rank_true_tails", "filt_rank_true_heads", "filt_rank_true_tails"]
    equal = all(torch.equal(getattr(ev, nm)[:sample], want) for nm, want in zip(names, ref))
    best = min(times)
    return {"workload": "Analogy emb_dim=%d (planes of %d) |E|=%d |R|=%d, %d test triples, full filtered LP"
                        % (emb_dim, emb_dim // 2, n_ent, n_rel, n_test),
            "triples_per_s_e2e": n_test / best, "ms_per_evaluate": [round(1e3 * x, 2) for x in times],
            "oracle_sample": {"n": sample, "ranks_equal": bool(equal), "cpu_triples_per_s": sample / cpu_s,
                              "threads": torch.get_num_threads()}}


def main():
    out = [run(14541, 237, 200, 20466, 256), run(1_000_000, 1000, 200, 20466, 8)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "analogy_perf.json"), "w") as f:
        json.dump(out, f, indent=1)
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
