"""C5 (BASELINE.json configs[4]): DistMult dim=200 training step -- Bernoulli negative sampling
(n_neg=256) fused with scoring + margin loss -- timed on one B200, forward and forward+backward,
against the HBM roofline (algorithmic bytes per positive = (n_neg+3)*4d
forward, ~3x that forward+backward), next to the unfused three-call path and to the reference
algorithm on the host cores (oracle port: gather, F.normalize, product, sum, MarginRankingLoss,
autograd backward).

    python scripts/train_bench.py [--n-ent 1000000] [--batch 4096 32768] [--model DistMult]
        -> one JSON line per batch size, also appended to gpurun_out/train_bench.jsonl
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchkge_b200 as tk  # noqa: E402
from torchkge_b200 import _lib, synthetic as S  # noqa: E402
from torchkge_b200.training import fused_margin_step  # noqa: E402


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="DistMult")
    ap.add_argument("--diss", default=None)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--n-ent", type=int, default=1000000)
    ap.add_argument("--n-rel", type=int, default=1000)
    ap.add_argument("--n-neg", type=int, default=256)
    ap.add_argument("--batch", type=int, nargs="+", default=[4096, 32768])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu-batch", type=int, default=4096, help="0 disables the CPU baseline")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    code = S.MODEL_CODE[(args.model, args.diss)]
    kind = S.ORACLE_KIND[code]
    cls = getattr(tk, args.model + "Model")
    torch.manual_seed(0)
    model = cls(args.dim, args.n_ent, args.n_rel) if args.diss is None else cls(
        args.dim, args.n_ent, args.n_rel, dissimilarity_type=args.diss)
    model = model.to(dev)
    planes = 2 if code in (_lib.COMPLEX, _lib.ROTATE) else 1
    row = 4 * args.dim * planes
    peak, src = peaks()
    g = torch.Generator(device=dev).manual_seed(1)
    probs = torch.rand(args.n_rel, generator=g, device=dev) * 0.8 + 0.1
    out_path = os.path.join(ROOT, "gpurun_out", "train_bench.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    for b in args.batch:
        h = torch.randint(0, args.n_ent, (b,), generator=g, device=dev)
        t = torch.randint(0, args.n_ent, (b,), generator=g, device=dev)
        r = torch.randint(0, args.n_rel, (b,), generator=g, device=dev)
        params = [p for p in model.parameters()]
        calls = [0]

        def fwd():
            calls[0] += 1
            with torch.no_grad():
                return fused_margin_step(model, h, t, r, 1.0, n_neg=args.n_neg, bern_probs=probs, seed=7,
                                         offset=calls[0])

        def fwd_bwd():
            calls[0] += 1
            for p in params:
                p.grad = None
            loss = fused_margin_step(model, h, t, r, 1.0, n_neg=args.n_neg, bern_probs=probs, seed=7,
                                     offset=calls[0])
            loss.backward()
            return loss

        sampler_like = dict(seed=7)

        def unfused():
            # the reference's three calls on the GPU kernels: corrupt_batch -> model(...) -> MarginLoss
            calls[0] += 1
            for p in params:
                p.grad = None
            nh = torch.empty(b * args.n_neg, dtype=torch.int64, device=dev)
            nt = torch.empty_like(nh)
            _lib.check(_lib.load().kge_corrupt_batch(h.data_ptr(), t.data_ptr(), r.data_ptr(), b, args.n_neg,
                                                     probs.data_ptr(), args.n_ent, sampler_like["seed"],
                                                     calls[0], nh.data_ptr(), nt.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "corrupt")
            pos, neg = model(h, t, r, nh, nt)
            loss = tk.MarginLoss(1.0)(pos, neg)
            loss.backward()
            return loss

        ms_f = timed(fwd, args.reps)
        ms_fb = timed(fwd_bwd, args.reps)
        # backward alone includes zero-filling the dense gradient tables (what autograd hands to
        # the optimizer); report the table-zeroing floor beside it
        grad_bytes = sum(p.numel() * 4 for p in params)
        try:
            ms_u = timed(unfused, max(2, args.reps // 3), warm=1)
        except Exception as e:  # memory at very large batches
            ms_u = None
            print("unfused path failed: %r" % (e,), file=sys.stderr)
        bytes_f = b * (args.n_neg + 3) * row
        bytes_fb = 3 * bytes_f
        rec = {
            "workload": "c5: %s%s dim=%d |E|=%d |R|=%d, B=%d, n_neg=%d, margin 1.0, Bernoulli corruption" % (
                args.model, "-" + args.diss if args.diss else "", args.dim, args.n_ent, args.n_rel, b, args.n_neg),
            "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "unfused_fwd_bwd_ms": ms_u,
            "positives_per_s_fwd": b / ms_f * 1e3, "positives_per_s_fwd_bwd": b / ms_fb * 1e3,
            "negatives_per_s_fwd_bwd": b * args.n_neg / ms_fb * 1e3,
            "roofline_fwd": {"bound": "hbm", "achieved": bytes_f / ms_f / 1e6, "peak": peak, "unit": "GB/s",
                             "frac": bytes_f / ms_f / 1e6 / peak, "alg_bytes": bytes_f},
            "roofline_fwd_bwd": {"bound": "hbm", "achieved": bytes_fb / ms_fb / 1e6, "peak": peak,
                                 "unit": "GB/s", "frac": bytes_fb / ms_fb / 1e6 / peak, "alg_bytes": bytes_fb,
                                 "dense_grad_zero_fill_bytes": grad_bytes},
            "peak_source": src,
        }
        if args.cpu_batch and b == args.cpu_batch:
            from oracle import kge_oracle as oracle
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in _oracle_params(kind, model).items()}
            hc, tc_, rc = h.cpu(), t.cpu(), r.cpu()
            nh, nt = oracle.corrupt_batch(hc, tc_, rc, probs.cpu(), args.n_ent, args.n_neg)
            t0 = time.perf_counter()
            pos, neg = oracle.forward_pos_neg(kind, P, hc, tc_, rc, nh, nt)
            loss = oracle.margin_loss(pos, neg, 1.0)
            t1 = time.perf_counter()
            loss.backward()
            t2 = time.perf_counter()
            rec["cpu_baseline"] = {"kind": "port", "cores": cores, "fwd_s": t1 - t0, "bwd_s": t2 - t1,
                                   "positives_per_s_fwd_bwd": b / (t2 - t0),
                                   "sample": "one step at B=%d (torch %s CPU)" % (b, torch.__version__)}
        line = json.dumps(rec)
        print(line, flush=True)
        with open(out_path, "a") as f:
            f.write(line + "\n")


def _oracle_params(kind, model):
    if kind in ("transe_l1", "transe_l2", "distmult"):
        return {"ent": model.ent_emb.weight, "rel": model.rel_emb.weight}
    if kind == "rescal":
        return {"ent": model.ent_emb.weight, "rel_mat": model.rel_mat.weight}
    return {"re_ent": model.re_ent_emb.weight, "im_ent": model.im_ent_emb.weight,
            "re_rel": model.re_rel_emb.weight, "im_rel": model.im_rel_emb.weight}


if __name__ == "__main__":
    main()
