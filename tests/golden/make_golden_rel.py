"""Relation-prediction golden vectors from the UNMODIFIED reference (torchkge v0.17.7 at
/root/reference).  Run in the authoring container only:

    PYTHONPATH=/root/reference python tests/golden/make_golden_rel.py

For every existing fixture <case>.npz (weights + facts already pinned there) this writes
rel_<case>.npz with the outputs of the reference's RelationPredictionEvaluator
(evaluation.py:16-204) -- raw and filtered ranks of the true relation, directed and undirected --
its dict_of_rels flattened to arrays, and the dense relation scores of a few (h, t) pairs
(inference_scoring_function, relation case).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from torchkge.data_structures import KnowledgeGraph  # noqa: E402
from torchkge.evaluation import RelationPredictionEvaluator  # noqa: E402
from torchkge.models import ComplExModel, DistMultModel, RESCALModel, TransEModel  # noqa: E402

from make_golden import dicts_to_arrays  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = ["toy_transe_l1", "toy_transe_l2", "toy_distmult", "toy_rescal", "toy_complex",
         "syn_transe_l1", "syn_transe_l2", "syn_distmult", "syn_rescal", "syn_complex"]


def build(kind, d, n_ent, n_rel):
    if kind == "transe_l1":
        return TransEModel(d, n_ent, n_rel, "L1")
    if kind == "transe_l2":
        return TransEModel(d, n_ent, n_rel, "L2")
    if kind == "distmult":
        return DistMultModel(d, n_ent, n_rel)
    if kind == "rescal":
        return RESCALModel(d, n_ent, n_rel)
    return ComplExModel(d, n_ent, n_rel)


def main():
    for case in CASES:
        z = np.load(os.path.join(OUT, case + ".npz"), allow_pickle=False)
        kind, d, n_ent, n_rel = str(z["kind"]), int(z["dim"]), int(z["n_ent"]), int(z["n_rel"])
        model = build(kind, d, n_ent, n_rel)
        model.load_state_dict({k[2:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("w:")})
        ah, at, ar = (torch.from_numpy(z[k].copy()).long() for k in ("all_heads", "all_tails", "all_rels"))
        ent2ix = {i: i for i in range(n_ent)}
        rel2ix = {i: i for i in range(n_rel)}
        full = KnowledgeGraph(kg={"heads": ah, "tails": at, "relations": ar}, ent2ix=ent2ix, rel2ix=rel2ix)
        th, tt, tr = (torch.from_numpy(z[k].copy()).long() for k in ("heads", "tails", "rels"))
        test = KnowledgeGraph(kg={"heads": th, "tails": tt, "relations": tr}, ent2ix=ent2ix, rel2ix=rel2ix,
                              dict_of_heads=full.dict_of_heads, dict_of_tails=full.dict_of_tails,
                              dict_of_rels=full.dict_of_rels)
        out = {"kind": kind, "b_size": int(z["b_size"])}
        k, o, v = dicts_to_arrays(full.dict_of_rels)
        out["dr_keys"], out["dr_offs"], out["dr_vals"] = k, o, v
        for directed in (True, False):
            ev = RelationPredictionEvaluator(model, test, directed=directed)
            ev.evaluate(b_size=int(z["b_size"]), verbose=False)
            tag = "dir" if directed else "undir"
            out["rank_true_rels_" + tag] = ev.rank_true_rels.numpy()
            out["filt_rank_true_rels_" + tag] = ev.filt_rank_true_rels.numpy()
            out["metrics_" + tag] = np.array([*ev.mean_rank(), *ev.hit_at_k(3), *ev.mrr()], dtype=np.float64)
        nq = min(8, th.shape[0])
        with torch.no_grad():
            he, te, _, cands = model.inference_prepare_candidates(th[:nq], tt[:nq], tr[:nq], entities=False)
            out["scores_rel"] = model.inference_scoring_function(he, te, cands).numpy()
        path = os.path.join(OUT, "rel_" + case + ".npz")
        np.savez_compressed(path, **out)
        print(case, "->", os.path.getsize(path), "bytes", "MR", out["metrics_dir"][:2])


if __name__ == "__main__":
    main()
