"""Generates tests/golden/*.npz by running the UNMODIFIED reference (torchkge v0.17.7 at
/root/reference) on seeded inputs.  Run in the authoring container only:

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Each file holds the inputs (weights, triples, filter dictionaries flattened to arrays) and
the reference's outputs for the hot path: the four rank vectors of
LinkPredictionEvaluator.evaluate, dense inference scores for a few queries, per-triple
scoring_function values, Model.forward + MarginLoss with fixed negatives, and the Bernoulli
probabilities.  The oracle (oracle/kge_oracle.py) and the CUDA path are both tested against
these files, on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, "/root/reference")
import torchkge  # noqa: E402
from torchkge.data_structures import KnowledgeGraph  # noqa: E402
from torchkge.evaluation import LinkPredictionEvaluator  # noqa: E402
from torchkge.models import ComplExModel, DistMultModel, RESCALModel, TransEModel  # noqa: E402
from torchkge.sampling import BernoulliNegativeSampler  # noqa: E402
from torchkge.utils import MarginLoss  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

TOY = [[0, 1, 0], [0, 2, 0], [0, 3, 0], [0, 4, 0], [1, 2, 1], [1, 3, 2], [2, 4, 0], [3, 4, 4],
       [5, 4, 0]]  # reference tests/test_evaluation.py:14-15 (columns from, to, rel)


def dicts_to_arrays(d):
    keys = sorted(d.keys())
    k = np.array(keys, dtype=np.int64).reshape(-1, 2)
    offs = np.zeros(len(keys) + 1, dtype=np.int64)
    vals = []
    for i, key in enumerate(keys):
        v = sorted(d[key])
        vals.extend(v)
        offs[i + 1] = len(vals)
    return k, offs, np.array(vals, dtype=np.int64)


def synthetic_df(n_ent, n_rel, n_facts, seed):
    g = torch.Generator().manual_seed(seed)
    w_e = 1.0 / torch.arange(1, n_ent + 1, dtype=torch.float64) ** 0.8
    w_r = 1.0 / torch.arange(1, n_rel + 1, dtype=torch.float64)
    h = torch.multinomial(w_e, n_facts, replacement=True, generator=g)
    t = torch.multinomial(w_e, n_facts, replacement=True, generator=g)
    r = torch.multinomial(w_r, n_facts, replacement=True, generator=g)
    trip = torch.unique(torch.stack([h, t, r], 1), dim=0)
    trip = trip[torch.randperm(trip.shape[0], generator=g)]
    # make sure every entity / relation id appears so that n_ent / n_rel are as asked
    extra = torch.stack([torch.arange(n_ent), (torch.arange(n_ent) + 1) % n_ent,
                         torch.arange(n_ent) % n_rel], 1)
    trip = torch.unique(torch.cat([trip, extra]), dim=0)
    trip = trip[torch.randperm(trip.shape[0], generator=g)]
    return pd.DataFrame(trip.numpy(), columns=["from", "to", "rel"])


def build(kind, d, n_ent, n_rel):
    if kind == "transe_l1":
        return TransEModel(d, n_ent, n_rel, "L1")
    if kind == "transe_l2":
        return TransEModel(d, n_ent, n_rel, "L2")
    if kind == "distmult":
        return DistMultModel(d, n_ent, n_rel)
    if kind == "rescal":
        return RESCALModel(d, n_ent, n_rel)
    if kind == "complex":
        return ComplExModel(d, n_ent, n_rel)
    raise ValueError(kind)


def run_case(name, kind, d, df, n_test, seed, b_size):
    kg_full = KnowledgeGraph(df=df)
    n_ent, n_rel = kg_full.n_ent, kg_full.n_rel
    test = KnowledgeGraph(kg={"heads": kg_full.head_idx[:n_test], "tails": kg_full.tail_idx[:n_test],
                              "relations": kg_full.relations[:n_test]},
                          ent2ix=kg_full.ent2ix, rel2ix=kg_full.rel2ix,
                          dict_of_heads=kg_full.dict_of_heads, dict_of_tails=kg_full.dict_of_tails,
                          dict_of_rels=kg_full.dict_of_rels)
    torch.manual_seed(seed)
    model = build(kind, d, n_ent, n_rel)
    # perturb away from the normalised init so scoring_function's on-the-fly normalisation
    # is exercised
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.0 + 0.5 * torch.rand(p.shape[0], 1))
    out = {"kind": kind, "dim": d, "n_ent": n_ent, "n_rel": n_rel, "b_size": b_size,
           "torch_version": torch.__version__, "reference_version": torchkge.__version__}
    for k, v in model.state_dict().items():
        out["w:" + k] = v.numpy().copy()
    out["heads"], out["tails"], out["rels"] = (test.head_idx.numpy(), test.tail_idx.numpy(),
                                               test.relations.numpy())
    out["all_heads"], out["all_tails"], out["all_rels"] = (kg_full.head_idx.numpy(),
                                                           kg_full.tail_idx.numpy(),
                                                           kg_full.relations.numpy())
    for nm, dd in (("dh", kg_full.dict_of_heads), ("dt", kg_full.dict_of_tails)):
        k, o, v = dicts_to_arrays(dd)
        out[nm + "_keys"], out[nm + "_offs"], out[nm + "_vals"] = k, o, v

    ev = LinkPredictionEvaluator(model, test)
    ev.evaluate(b_size=b_size, verbose=False)
    out["rank_true_heads"] = ev.rank_true_heads.numpy()
    out["rank_true_tails"] = ev.rank_true_tails.numpy()
    out["filt_rank_true_heads"] = ev.filt_rank_true_heads.numpy()
    out["filt_rank_true_tails"] = ev.filt_rank_true_tails.numpy()
    out["metrics"] = np.array([*ev.mean_rank(), *ev.hit_at_k(10), *ev.mrr()], dtype=np.float64)

    nq = min(8, n_test)
    h, t, r = test.head_idx[:nq], test.tail_idx[:nq], test.relations[:nq]
    with torch.no_grad():
        he, te, re_, cands = model.inference_prepare_candidates(h, t, r, entities=True)
        out["scores_tail"] = model.inference_scoring_function(he, cands, re_).numpy()
        out["scores_head"] = model.inference_scoring_function(cands, te, re_).numpy()
        out["triple_scores"] = model.scoring_function(test.head_idx, test.tail_idx,
                                                      test.relations).numpy()

    # training-side golden values with fixed negatives
    sampler = BernoulliNegativeSampler(kg_full, n_neg=3)
    out["bern_probs"] = sampler.bern_probs.numpy()
    torch.manual_seed(seed + 1)
    nh, nt = sampler.corrupt_batch(test.head_idx, test.tail_idx, test.relations)
    out["neg_heads"], out["neg_tails"] = nh.numpy(), nt.numpy()
    pos, neg = model(test.head_idx, test.tail_idx, test.relations, nh, nt)
    loss = MarginLoss(0.5)(pos, neg)
    model.zero_grad()
    loss.backward()
    out["fwd_pos"], out["fwd_neg"] = pos.detach().numpy(), neg.detach().numpy()
    out["loss_margin_0p5"] = np.array(loss.item(), dtype=np.float64)
    for k, p in model.named_parameters():
        out["g:" + k] = p.grad.numpy().copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, kind, "n_ent", n_ent, "n_rel", n_rel, "test", n_test, "->", os.path.getsize(path), "bytes")


def main():
    toy = pd.DataFrame(TOY, columns=["from", "to", "rel"])
    for kind, d in (("transe_l1", 100), ("transe_l2", 100), ("distmult", 100), ("rescal", 20),
                    ("complex", 100)):
        run_case("toy_" + kind, kind, d, toy, 9, seed=7, b_size=9)
    df = synthetic_df(300, 9, 2500, seed=11)
    for kind, d in (("transe_l1", 50), ("transe_l2", 50), ("distmult", 64), ("rescal", 24),
                    ("complex", 36)):
        run_case("syn_" + kind, kind, d, df, 160, seed=13, b_size=40)


if __name__ == "__main__":
    main()
