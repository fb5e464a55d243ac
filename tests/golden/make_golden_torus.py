"""TorusE golden vectors from the UNMODIFIED reference (torchkge v0.17.7 at /root/reference).
Run in the authoring container only:

    PYTHONPATH=/root/reference python tests/golden/make_golden_torus.py

Writes torus_<l1|l2>.npz: the weights of a perturbed TorusEModel (translation.py:655-767) as they
are after normalize_parameters (fractional parts), the facts and filter dictionaries of the
synthetic graph of make_golden.py, the four rank vectors of the reference's
LinkPredictionEvaluator, dense inference scores of a few queries and per-triple scoring_function
values.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from torchkge.data_structures import KnowledgeGraph  # noqa: E402
from torchkge.evaluation import LinkPredictionEvaluator  # noqa: E402
from torchkge.models import TorusEModel  # noqa: E402

from make_golden import dicts_to_arrays, synthetic_df  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    df = synthetic_df(300, 9, 2500, seed=11)
    kg_full = KnowledgeGraph(df=df)
    n_ent, n_rel, n_test = kg_full.n_ent, kg_full.n_rel, 160
    test = KnowledgeGraph(kg={"heads": kg_full.head_idx[:n_test], "tails": kg_full.tail_idx[:n_test],
                              "relations": kg_full.relations[:n_test]},
                          ent2ix=kg_full.ent2ix, rel2ix=kg_full.rel2ix,
                          dict_of_heads=kg_full.dict_of_heads, dict_of_tails=kg_full.dict_of_tails,
                          dict_of_rels=kg_full.dict_of_rels)
    for tag, diss, d in (("l1", "torus_L1", 50), ("l2", "torus_L2", 36)):
        torch.manual_seed(17)
        model = TorusEModel(d, n_ent, n_rel, diss)
        with torch.no_grad():   # spread the values over the torus (Xavier init is tiny)
            model.ent_emb.weight.mul_(40.0)
            model.rel_emb.weight.mul_(40.0)
        model.normalize_parameters()
        out = {"kind": "toruse_" + tag, "dim": d, "n_ent": n_ent, "n_rel": n_rel, "b_size": 40}
        for k, v in model.state_dict().items():
            out["w:" + k] = v.numpy().copy()
        out["heads"], out["tails"], out["rels"] = (test.head_idx.numpy(), test.tail_idx.numpy(),
                                                   test.relations.numpy())
        for nm, dd in (("dh", kg_full.dict_of_heads), ("dt", kg_full.dict_of_tails)):
            k, o, v = dicts_to_arrays(dd)
            out[nm + "_keys"], out[nm + "_offs"], out[nm + "_vals"] = k, o, v
        ev = LinkPredictionEvaluator(model, test)
        ev.evaluate(b_size=40, verbose=False)
        out["rank_true_heads"] = ev.rank_true_heads.numpy()
        out["rank_true_tails"] = ev.rank_true_tails.numpy()
        out["filt_rank_true_heads"] = ev.filt_rank_true_heads.numpy()
        out["filt_rank_true_tails"] = ev.filt_rank_true_tails.numpy()
        out["metrics"] = np.array([*ev.mean_rank(), *ev.hit_at_k(10), *ev.mrr()], dtype=np.float64)
        h, t, r = test.head_idx[:8], test.tail_idx[:8], test.relations[:8]
        with torch.no_grad():
            he, te, re_, cands = model.inference_prepare_candidates(h, t, r, entities=True)
            out["scores_tail"] = model.inference_scoring_function(he, cands, re_).numpy()
            out["scores_head"] = model.inference_scoring_function(cands, te, re_).numpy()
            out["triple_scores"] = model.scoring_function(test.head_idx, test.tail_idx, test.relations).numpy()
        path = os.path.join(OUT, "torus_%s.npz" % tag)
        np.savez_compressed(path, **out)
        print(tag, diss, "->", os.path.getsize(path), "bytes; MR", out["metrics"][:2])


if __name__ == "__main__":
    main()
