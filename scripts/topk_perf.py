"""Top-k inference probe: EntityInference through the public API at |E| = 1M (no (b, |E|) score matrix is
allocated: peak memory is reported), against the dense-scores-plus-torch.topk form it replaced.

    python scripts/topk_perf.py [n_ent] [n_queries] [k]   -> one line per model, also gpurun_out/topk_perf.txt
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchkge_b200 as tk  # noqa: E402
from torchkge_b200 import _lib  # noqa: E402
from torchkge_b200.engine import ModelSpec, default_engine  # noqa: E402


def main():
    n_ent = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda:0")
    lines = []
    for name, cls, d, kw in (("TransE-L2", tk.TransEModel, 200, {"dissimilarity_type": "L2"}),
                             ("DistMult", tk.DistMultModel, 200, {}), ("ComplEx", tk.ComplExModel, 200, {})):
        torch.manual_seed(0)
        model = cls(d, n_ent, 1000, **kw).to(dev)
        g = torch.Generator().manual_seed(1)
        ents = torch.randint(0, n_ent, (n_q,), generator=g)
        rels = torch.randint(0, 1000, (n_q,), generator=g)
        inf = tk.EntityInference(model, ents, rels, top_k=k, missing="tails")
        inf.evaluate(b_size=256, verbose=False)          # warm-up
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        t0 = time.perf_counter()
        inf.evaluate(b_size=256, verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        peak = torch.cuda.max_memory_allocated() - base
        # the form it replaced: dense scores (kge_score_all) + torch.topk, 256 queries at a time (1 GiB of scores)
        eng = default_engine()
        spec = ModelSpec.from_model(model)
        packed = eng.pack(spec)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        same = True
        for lo in range(0, n_q, 256):
            e = ents[lo:lo + 256].to(dev)
            rows = eng.gather_rows(spec, e)
            s = eng.score_all(spec, packed, _lib.SIDE_TAIL, rows, rows, rels[lo:lo + 256].to(dev))
            v, i = torch.topk(s, k, dim=1)
            same &= bool(torch.equal(v.cpu(), inf.scores[lo:lo + 256]))
        torch.cuda.synchronize()
        dt_dense = time.perf_counter() - t0
        line = ("%-10s d=%d |E|=%d queries=%d k=%d: scan-epilogue top-k %.1f ms (%.0f queries/s, peak extra memory %.0f MB; "
                "a (queries, |E|) fp32 matrix would be %.0f MB) | dense scores + torch.topk %.1f ms | top-k scores equal: %s"
                % (name, d, n_ent, n_q, k, dt * 1e3, n_q / dt, peak / 1e6, 4.0 * n_q * n_ent / 1e6, dt_dense * 1e3, same))
        print(line, flush=True)
        lines.append(line)
        del model, inf, packed
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/topk_perf.txt", "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
